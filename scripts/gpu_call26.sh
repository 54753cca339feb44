cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/sharded_check.py > gpurun_out/r2k_sharded_check_world2.log 2>&1
tail -n 3 gpurun_out/r2k_sharded_check_world2.log
for v in "GAD_SHARDED_SCATTER=staged" "GAD_SHARDED_SCATTER=direct"; do
env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu-baseline --generations 0 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
tail -n 3 gpurun_out/r2k_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2k_bench_n2.json') if l.startswith('{')][-1]); print('N2 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only']['ms_per_step'])"
done
