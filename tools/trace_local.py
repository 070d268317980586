"""cold multi-rank Domain::sync with the ranks as threads on ONE GPU (local communicator): lets ncu look at the LET
kernels without a multi-GPU box.  usage: trace_local.py [ranks] [particles_per_rank]"""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
import numpy as np
import torch

from cstone_b200 import capi

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16 * 1024 * 1024
dev = torch.device("cuda:0")
world = capi.LocalWorld(P)
hval = 0.5 * float(np.cbrt(3.0 * 100 / (4 * np.pi * n * P)))


def body(r):
    with torch.cuda.stream(torch.cuda.Stream(device=dev)):
        g = torch.Generator(device=dev)
        g.manual_seed(42 + r)
        x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
        h = torch.full((n,), hval, dtype=torch.float64, device=dev)
        comm = world.comm(r)
        dom = capi.Domain(r, P, max(64, n // 100), 64, 0.5, (0, 1, 0, 1, 0, 1), (1, 1, 1), key="u64", real="d",
                          device="cuda:0", comm=comm)
        try:
            for _ in range(2):
                dom.reset()
                dom.sync(x, y, z, h)
            torch.cuda.current_stream().synchronize()
            if r == 0:
                print("focus leaves", dom.num_focus_leaves, "assigned", dom.end_index - dom.start_index, flush=True)
        except BaseException:
            world.abort()
            raise


threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(P)]
for t in threads:
    t.start()
for t in threads:
    t.join(timeout=200)
