"""phase timings of a cold multi-rank Domain::sync (CSB_TRACE=1): torchrun ... tools/trace_multirank.py [n_per_gpu]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
os.environ["CSB_TRACE"] = "1"
import numpy as np
import torch
import torch.distributed as dist

from cstone_b200 import capi

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.Comm.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = capi.Comm.nccl(rank, world, bytes(uid.cpu().numpy().tobytes()))
g = torch.Generator(device=dev)
g.manual_seed(42 + rank)
x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
h = torch.full((n,), 0.5 * float(np.cbrt(3.0 * 100 / (4 * np.pi * n * world))), dtype=torch.float64, device=dev)
bnd = (1, 1, 1) if world > 1 else (0, 0, 0)
bucket = max(64, n // 100) if world > 1 else 64
dom = capi.Domain(rank, world, bucket, 64, 0.5, (0, 1, 0, 1, 0, 1), bnd, key="u64", real="d", device=str(dev),
                  comm=comm)
for it in range(3):
    if rank == 0:
        print(f"==== cold sync {it}", file=sys.stderr, flush=True)
    dom.reset()
    dom.sync(x, y, z, h)
for it in range(2):
    if rank == 0:
        print(f"==== steady sync {it}", file=sys.stderr, flush=True)
    dom.sync()
if world > 1:
    dist.destroy_process_group()
