cd $GRAFT_REPO_ROOT
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err ) 2> gpurun_out/r2_bench_n2.time
echo rc=$? >> gpurun_out/r2_bench_n2.time
tail -n 4 gpurun_out/r2_bench_n2.time; tail -n 12 gpurun_out/r2_bench_n2.err
timeout 600 python -m pytest tests/test_ga_gpu.py -x -q -m gpu -k "sharded or vertical or debug" 2>&1 | tail -5
