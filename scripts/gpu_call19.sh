cd $GRAFT_REPO_ROOT
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sharded_check.py > gpurun_out/r2f_sharded_check_world2.log 2>&1 ) 2> gpurun_out/r2f_sharded_check.time
tail -n 4 gpurun_out/r2f_sharded_check_world2.log; tail -n 3 gpurun_out/r2f_sharded_check.time
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err ) 2> gpurun_out/r2f_bench_n2.time
tail -n 4 gpurun_out/r2f_bench_n2.time; tail -n 6 gpurun_out/r2f_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2f_bench_n2.json') if l.startswith('{')][-1]); print('N2', d['value'], d['ms_per_step'], d['fitness_kernel_only']); g=d['generation']; print(g['ms_per_generation'], g['phase_ms']); print(d['parity']); print(d.get('strong_scaling_c4'))"
