set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_fitness_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_t1.log
timeout 600 python -m pytest tests/test_mutation_scale_gpu.py -x -q -m gpu -k "batch or duplicate" 2>&1 | tail -8 > gpurun_out/r2_t2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_tma.json 2> gpurun_out/r2_fit_tma.err
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_c4_tma.json 2> gpurun_out/r2_fit_c4_tma.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 6 --generation-warmup 2 > gpurun_out/r2_b_gen.json 2> gpurun_out/r2_b_gen.err
GAD_BENCH_NO_BURN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fitness_tma_kernel -s 2 -c 2 -f -o gpurun_out/r2_fit_tma_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r2_ncu_fit.log 2>&1
tail -n 6 gpurun_out/r2_t1.log gpurun_out/r2_t2.log
