#!/bin/bash
# usage: tools/run_bench_n.sh N [extra bench args]   (torchrun launch as the driver does it)
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"
