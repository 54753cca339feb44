cd $GRAFT_REPO_ROOT
for v in "GAD_FITNESS_GS=16" "GAD_FITNESS_GS=16 GAD_FT_CONSUMERS=448" "GAD_FITNESS_GS=8 GAD_FT_SMEM_KB=200"; do
env $v GAD_BENCH_NO_BURN=0 timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --generations 0 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err
tail -n 2 gpurun_out/r2h_bench_c4.err; python -c "
import json; d=json.loads(open('gpurun_out/r2h_bench_c4.json').read().strip().splitlines()[-1]); print('c4 $v', d['ms_per_step'], d['fitness_kernel_only']['ms_per_step'], d['roofline']['frac'])"
done
