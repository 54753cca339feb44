cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_fitness_gpu.py tests/test_ga_gpu.py -x -q -m gpu -k "fitness_pass or hall_of_fame or traces or larger_population or tma or baseline_shapes" 2>&1 | tail -15 ) > gpurun_out/r2d_tests.log 2>&1
cat gpurun_out/r2d_tests.log
for v in "GAD_FITNESS_FUSED_HOF=1" "GAD_FITNESS_FUSED_HOF=0"; do
env $v timeout 200 python bench.py --steps 20 --warmup 5 --generations 0 --no-cpu-baseline > gpurun_out/r2d_bench_pass_$v.json 2> gpurun_out/r2d_bench_pass.err
tail -n 3 gpurun_out/r2d_bench_pass.err; python -c "
import json; d=json.loads(open('gpurun_out/r2d_bench_pass_$v.json').read().strip().splitlines()[-1]); print('N1 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only'], d['gpu_launches'])"
done
