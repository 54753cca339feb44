cd $GRAFT_REPO_ROOT
( time timeout 500 python -m pytest tests/test_fitness_gpu.py tests/test_ga_gpu.py -x -q -m gpu -k "fitness or hall_of_fame or baseline_shapes or larger_population or tma" 2>&1 | tail -5 ) > gpurun_out/r2i_tests.log 2>&1
cat gpurun_out/r2i_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --generations 0 --no-cpu-baseline > gpurun_out/r2i_bench_pass.json 2> gpurun_out/r2i_bench_pass.err
tail -n 3 gpurun_out/r2i_bench_pass.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench_pass.json').read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], d['fitness_kernel_only'], d['roofline']['frac'])"
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --generations 0 --no-cpu-baseline --no-extras > gpurun_out/r2i_bench_c4.json 2> gpurun_out/r2i_bench_c4.err
tail -n 2 gpurun_out/r2i_bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2i_bench_c4.json').read().strip().splitlines()[-1]); print('c4', d['value'], d['ms_per_step'], d['fitness_kernel_only']['ms_per_step'], d['roofline']['frac'])"
