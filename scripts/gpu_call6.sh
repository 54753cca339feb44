cd $GRAFT_REPO_ROOT
( time python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2> gpurun_out/r2_bench_n1.time
echo rc=$? >> gpurun_out/r2_bench_n1.time
( time python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2> gpurun_out/r2_bench_ref.time
tail -n 4 gpurun_out/r2_bench_n1.time gpurun_out/r2_bench_ref.time; tail -n 5 gpurun_out/r2_bench_n1.err
