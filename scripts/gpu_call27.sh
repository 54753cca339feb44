cd $GRAFT_REPO_ROOT
for v in "GAD_SHARDED_SCATTER=staged" "GAD_SHARDED_SCATTER=direct"; do
env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --no-cpu-baseline --generations 0 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err
tail -n 3 gpurun_out/r2k_bench_n4.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2k_bench_n4.json') if l.startswith('{')][-1]); print('N4 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only']['ms_per_step'])"
done
