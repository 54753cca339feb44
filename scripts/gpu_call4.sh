set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_fitness_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_t1.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_tma.json 2> gpurun_out/r2_fit_tma.err
GAD_FITNESS_TMA=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_reg.json 2> gpurun_out/r2_fit_reg.err
timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_c4_tma.json 2> gpurun_out/r2_fit_c4_tma.err
GAD_FITNESS_TMA=0 timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r2_fit_c4_reg.json 2> gpurun_out/r2_fit_c4_reg.err
GAD_BENCH_NO_BURN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_gen.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --generations 2 --generation-warmup 3 > gpurun_out/r2_ncu_gen.log 2>&1
tail -n 8 gpurun_out/r2_t1.log
