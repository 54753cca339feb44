cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_ga_gpu.py -x -q -m gpu -k "hall_of_fame or traces or sharded" 2>&1 | tail -15 > gpurun_out/r2b_hof_tests.log
cat gpurun_out/r2b_hof_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sharded_check.py > gpurun_out/r2b_sharded_check_world2.log 2>&1
tail -n 5 gpurun_out/r2b_sharded_check_world2.log
timeout 200 python bench.py --steps 20 --warmup 5 --generations 0 --no-cpu-baseline > gpurun_out/r2b_bench_n1_pass.json 2> gpurun_out/r2b_bench_n1_pass.err
tail -n 3 gpurun_out/r2b_bench_n1_pass.err; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_n1_pass.json').read().strip().splitlines()[-1]); print('N1', d['value'], d['ms_per_step'], d['fitness_pass'], d['fitness_kernel_only'])"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err ) 2> gpurun_out/r2b_bench_n2.time
tail -n 4 gpurun_out/r2b_bench_n2.time; tail -n 6 gpurun_out/r2b_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2b_bench_n2.json') if l.startswith('{')][-1]); print('N2', d['value'], d['ms_per_step'], d['fitness_pass'], d['fitness_kernel_only']); g=d['generation']; print(g['ms_per_generation'], g['phase_ms']); print(d['parity'])"
