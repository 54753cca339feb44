cd $GRAFT_REPO_ROOT
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2c_gpu_tests.log 2>&1
cat gpurun_out/r2c_gpu_tests.log
( time timeout 500 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err ) 2> gpurun_out/r2c_bench_n1.time
tail -n 4 gpurun_out/r2c_bench_n1.time; tail -n 3 gpurun_out/r2c_bench_n1.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c_bench_n1.json') if l.startswith('{')][-1]); print('N1', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']); g=d['generation']; print(g['ms_per_generation'], g['phase_ms']); print(d['parity']['equal'])"
