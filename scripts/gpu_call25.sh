cd $GRAFT_REPO_ROOT
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/r2j_bench_n4.json 2> gpurun_out/r2j_bench_n4.err ) 2> gpurun_out/r2j_bench_n4.time
tail -n 4 gpurun_out/r2j_bench_n4.time; tail -n 6 gpurun_out/r2j_bench_n4.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2j_bench_n4.json') if l.startswith('{')][-1]); print('N4', d['value'], d['ms_per_step'], d['fitness_kernel_only']); g=d['generation']; print(g['ms_per_generation'], g['phase_ms']); print(d['parity']); s=d.get('strong_scaling_c4'); print(s and (s['ms_per_generation'], s['phase_ms']))"
