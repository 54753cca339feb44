cd $GRAFT_REPO_ROOT
( time timeout 500 python -m pytest tests/test_fitness_gpu.py -x -q -m gpu 2>&1 | tail -5 ) > gpurun_out/r2g_tests.log 2>&1
cat gpurun_out/r2g_tests.log
for v in "GAD_FT_REVERSE=1" "GAD_FT_REVERSE=0"; do
env $v timeout 200 python bench.py --steps 20 --warmup 5 --generations 0 --no-cpu-baseline > gpurun_out/r2g_bench_pass_$v.json 2> gpurun_out/r2g_bench_pass.err
tail -n 3 gpurun_out/r2g_bench_pass.err; python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_pass_$v.json').read().strip().splitlines()[-1]); print('N1 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only'], d['roofline']['frac'])"
done
( time timeout 400 python bench.py --workload c4 --steps 5 --warmup 3 --generations 0 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err ) 2> gpurun_out/r2g_bench_c4.time
tail -n 3 gpurun_out/r2g_bench_c4.time; tail -n 3 gpurun_out/r2g_bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench_c4.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'], d['fitness_kernel_only'], d['roofline'])"
GAD_BENCH_NO_BURN=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:fitness_tma_kernel -s 4 -c 2 -o gpurun_out/r02_c4_fitness python bench.py --workload c4 --steps 3 --warmup 3 --generations 0 --no-cpu-baseline --no-extras > gpurun_out/r2g_ncu_c4.log 2>&1
tail -n 3 gpurun_out/r2g_ncu_c4.log
ncu -i gpurun_out/r02_c4_fitness.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_fitness_tma_c4.csv 2> /dev/null
ls -la gpurun_out/ | tail -8
