# Final round-2 evidence for profiles/: the default bench line (no profiler), then the ncu launch list and the
# --set full capture of the fitness kernel (bare launches and launches that carry the hall-of-fame decision).
cd $GRAFT_REPO_ROOT
( time python bench.py > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err ) 2> gpurun_out/r02_bench_c3_n1.time
tail -n 3 gpurun_out/r02_bench_c3_n1.time
( time GAD_BENCH_NO_BURN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_generation_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --generations 2 --generation-warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1 ) 2> gpurun_out/r02_ncu_launches.time
tail -n 3 gpurun_out/r02_ncu_launches.time
GAD_BENCH_NO_BURN=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fitness_tma_kernel -s 6 -c 6 -f -o gpurun_out/r02_fit_tma_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r02_ncu_fit.log 2>&1
ncu -i gpurun_out/r02_fit_tma_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_fitness_tma_c3.csv 2>/dev/null
rm -f gpurun_out/r02_fit_tma_c3.ncu-rep
ls -la gpurun_out | tail -8
