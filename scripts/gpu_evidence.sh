# Round evidence for profiles/: bench lines (no profiler), then the ncu launch list and --set full captures.
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err
python bench.py --impl reference > gpurun_out/r02_bench_c3_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python bench.py --workload c4 --steps 10 --warmup 3 --generations 3 --generation-warmup 4 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4.err
GAD_BENCH_NO_BURN=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_generation_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --generations 2 --generation-warmup 3 > gpurun_out/r02_ncu_launches.log 2>&1
GAD_BENCH_NO_BURN=1 ncu --set full --clock-control none --import-source on -k regex:fitness_tma_kernel -s 3 -c 3 -f -o gpurun_out/r02_fit_tma_c3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --generations 0 > gpurun_out/r02_ncu_fit.log 2>&1
GAD_BENCH_NO_BURN=1 ncu --set full --clock-control none -k regex:"encode_tc|gemm_split_f16|splitk_reduce" -s 40 -c 12 -f -o gpurun_out/r02_qnet_c3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --generations 1 --generation-warmup 2 > gpurun_out/r02_ncu_qnet.log 2>&1
ncu -i gpurun_out/r02_fit_tma_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_fitness_tma_c3.csv 2>/dev/null
ncu -i gpurun_out/r02_qnet_c3.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_qnet_c3.csv 2>/dev/null
rm -f gpurun_out/r02_qnet_c3.ncu-rep
ls -la gpurun_out | tail -20
