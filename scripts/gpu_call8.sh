cd $GRAFT_REPO_ROOT
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2> gpurun_out/r2_bench_n$N.time
echo rc=$? >> gpurun_out/r2_bench_n$N.time
tail -n 4 gpurun_out/r2_bench_n$N.time; tail -n 12 gpurun_out/r2_bench_n$N.err
