cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_tests_all.log
python bench.py --workload c4 --steps 10 --warmup 3 --generations 3 --generation-warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4.err
cat gpurun_out/r2_tests_all.log; tail -n 3 gpurun_out/r02_bench_c4.err
