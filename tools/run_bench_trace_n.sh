cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -n 3 gpurun_out/r2_bench_n$N.err
CSB_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 1 --warmup 3 --parity-n 65536 > gpurun_out/r2_trace_n$N.json 2> gpurun_out/r2_trace_n$N.err
grep -c "csb rank 0" gpurun_out/r2_trace_n$N.err
