cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2_tests_all.log
cat gpurun_out/r2_tests_all.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err ) 2> gpurun_out/r2_bench_n2.time
echo rc=$? >> gpurun_out/r2_bench_n2.time; tail -n 4 gpurun_out/r2_bench_n2.time; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2_bench_n2.err | tail -5
