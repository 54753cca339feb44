set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_mutation_scale_gpu.py -x -q -m gpu -s 2>&1 | tail -30 > gpurun_out/r2_t1.log
python -m pytest tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_t2.log
python -m pytest tests -x -q -m gpu --deselect tests/test_mutation_scale_gpu.py --deselect tests/test_dropin_gpu.py 2>&1 | tail -15 > gpurun_out/r2_t3.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 1 > gpurun_out/r2_b_dedup.json 2> gpurun_out/r2_b_dedup.err
GAD_MUTATE_DEDUP=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 1 > gpurun_out/r2_b_nodedup.json 2> gpurun_out/r2_b_nodedup.err
tail -5 gpurun_out/r2_t1.log gpurun_out/r2_t2.log gpurun_out/r2_t3.log
