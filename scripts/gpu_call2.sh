set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_mutation_scale_gpu.py -x -q -m gpu -s 2>&1 | tail -30 > gpurun_out/r2_t1.log
python -m pytest tests/test_ga_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_t3.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 2 > gpurun_out/r2_b_splitk.json 2> gpurun_out/r2_b_splitk.err
GAD_L1_KSPLIT=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 2 > gpurun_out/r2_b_nosplitk.json 2> gpurun_out/r2_b_nosplitk.err
tail -n 5 gpurun_out/r2_t1.log gpurun_out/r2_t3.log
