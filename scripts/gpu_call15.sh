cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_ga_gpu.py -x -q -m gpu -k "hall_of_fame or traces or sharded" 2>&1 | tail -15 > gpurun_out/r2b_hof_tests.log
cat gpurun_out/r2b_hof_tests.log
for v in "" "GAD_HOF_PDL=0" "GAD_HOF_CLUSTER=0" "GAD_HOF_UNFUSED=1"; do
env $v timeout 200 python bench.py --steps 20 --warmup 5 --generations 0 --no-cpu-baseline > gpurun_out/r2b_bench_n1_pass.json 2> gpurun_out/r2b_bench_n1_pass.err
tail -n 3 gpurun_out/r2b_bench_n1_pass.err; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_n1_pass.json').read().strip().splitlines()[-1]); print('N1 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only'])"
done
for v in "" "GAD_HOF_PDL=0" "GAD_HOF_CLUSTER=0"; do
env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --no-cpu-baseline --generations 0 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -n 3 gpurun_out/r2b_bench_n2.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2b_bench_n2.json') if l.startswith('{')][-1]); print('N2 $v', d['value'], d['ms_per_step'], d['fitness_kernel_only'])"
done
