"""NVLink store bandwidth of the exchangeParticles pack kernel against a copy-engine transfer (2 GPUs, one process)."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

n = 32 * 1024 * 1024
d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
src = [torch.rand(n, dtype=torch.float64, device=d0) for _ in range(4)]
loc = [torch.empty(n, dtype=torch.float64, device=d0) for _ in range(4)]
rem = [torch.empty(n, dtype=torch.float64, device=d1) for _ in range(4)]
rem[0].copy_(src[0])  # enables peer access both ways
src[0].copy_(rem[0])
torch.cuda.synchronize(d0)
torch.cuda.synchronize(d1)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(d0)
        torch.cuda.synchronize(d1)
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(d0)
        torch.cuda.synchronize(d1)
        best = min(best, time.perf_counter() - t0)
    return best


gb = 4 * n * 8 / 1e9
with torch.cuda.device(d0):
    t = timed(lambda: [r.copy_(s, non_blocking=True) for r, s in zip(rem, src)])
    print(f"copy engine, 4 x {n * 8 >> 20} MiB: {gb / t:7.1f} GB/s")
    lib = capi.lib()
    for name, order in (("identity", torch.arange(n, dtype=torch.int32, device=d0)),
                        ("random", torch.randperm(n, device=d0).to(torch.int32))):
        for where, dst in (("local", loc), ("remote", rem)):
            sp = (C.c_void_p * 4)(*[s.data_ptr() for s in src])
            dp = (C.c_void_p * 4)(*[d.data_ptr() for d in dst])
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            t = timed(lambda: lib.cs_gather4(C.c_void_p(order.data_ptr()), C.c_size_t(n), sp, dp, C.c_int(8), st))
            print(f"cs_gather4 {name:8s} -> {where:6s}: {gb / t:7.1f} GB/s  ({t * 1e3:.2f} ms)")
    assert torch.equal(rem[1].to(d0), src[1][order.to(torch.int64)])
