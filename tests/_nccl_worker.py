"""worker of tests/test_gpu_nccl.py: one rank of a multi-rank Domain over NCCL (one process per GPU); writes what
Domain::sync left behind to <outdir>/rank<r>.npz"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

outdir, inputs = sys.argv[1], sys.argv[2]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(capi.Comm.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
comm = capi.Comm.nccl(rank, world, bytes(uid.cpu().numpy().tobytes()))

d = np.load(inputs)
off = d["offsets"]
sl = slice(int(off[rank]), int(off[rank + 1]))
to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(dev)  # noqa: E731
dom = capi.Domain(rank, world, int(d["bucket"]), int(d["bucket_focus"]), 0.5, tuple(d["lim"]), tuple(d["bnd"]),
                  key="u64", real="d", device=str(dev), comm=comm)
dom.sync(to(d["x"]), to(d["y"]), to(d["z"]), to(d["h"]))
for _ in range(int(d["num_syncs"]) - 1):
    dom.sync()
out = {k: dom.field(k).cpu().numpy() for k in ("keys", "x", "y", "z", "h", "focus_leaves", "layout", "global_leaves")}
rho = dom.field("x") * 2 + dom.field("y")
f = torch.full_like(rho, -7)
f[dom.start_index:dom.end_index] = rho[dom.start_index:dom.end_index]
dom.exchange_halos(f)
out["halo_field_ok"] = np.array(bool(torch.equal(f, rho)))
nb, nc = dom.find_neighbors(64)
out.update(start=dom.start_index, end=dom.end_index, nc=nc.cpu().numpy(), nb=nb.cpu().numpy(),
           bytes_sent=comm.bytes_sent)
np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
dom.close()
dist.barrier()
dist.destroy_process_group()
