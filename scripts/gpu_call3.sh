set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_mutation_scale_gpu.py -x -q -m gpu -s -k "batch or action or duplicate" 2>&1 | tail -12 > gpurun_out/r2_t1.log
python -m pytest tests/test_ga_gpu.py -x -q -m gpu -k "qnet or traces or mutation" 2>&1 | tail -5 > gpurun_out/r2_t3.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 2 > gpurun_out/r2_b_splitk.json 2> gpurun_out/r2_b_splitk.err
tail -n 5 gpurun_out/r2_t1.log gpurun_out/r2_t3.log
