cd $GRAFT_REPO_ROOT
run() {
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --generations 0 2>gpurun_out/sweep_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$name', 'kernel_us %.2f'%(r['kernel_ms']*1e3), 'frac %.3f'%r['frac'], 'value %.3g'%d['value'])" >> gpurun_out/r2_sweep.txt 2>&1
}
rm -f gpurun_out/r2_sweep.txt
timeout 300 python -m pytest tests/test_fitness_gpu.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r2_sweep.txt
run pdl X=1
run nopdl GAD_FT_PDL=0
run pdl_ctas2 GAD_FT_CTAS=2 GAD_FT_BOX_BYTES=6144
run pdl_ctas4 GAD_FT_CTAS=4 GAD_FT_BOX_BYTES=6144
run pdl_ctas2_12k GAD_FT_CTAS=2
run pdl_box6k GAD_FT_BOX_BYTES=6144
cat gpurun_out/r2_sweep.txt
tail -n 3 gpurun_out/sweep_pdl.err
