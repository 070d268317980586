"""times cs_merge_sorted_runs_u64 on sorted runs of 8 Mi (key, index) pairs (the received blocks of a multi-rank sync) and
checks the result against a stable sort.  usage: exp_merge.py [runs]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

dev = torch.device("cuda:0")
runs, per = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 8 * 1024 * 1024
n = runs * per
g = torch.Generator(device=dev)
g.manual_seed(1)
keys0 = torch.randint(0, 1 << 62, (n,), device=dev, generator=g, dtype=torch.int64)
keys0 = keys0.view(runs, per).sort(dim=1).values.reshape(-1).contiguous()
vals0 = torch.arange(n, device=dev, dtype=torch.int32)
want = torch.sort(keys0, stable=True)
offsets = (C.c_size_t * (runs + 1))(*[i * per for i in range(runs + 1)])
kb, vb = torch.empty_like(keys0), torch.empty_like(vals0)
f = capi.lib().cs_merge_sorted_runs_u64
for shape in (0,):
    ms = []
    for _ in range(3):
        k, v = keys0.clone(), vals0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi._check(f(capi._ptr(k), capi._ptr(v), offsets, C.c_int(runs), capi._ptr(kb), capi._ptr(vb), capi._stream()),
                    "merge")
        e1.record()
        torch.cuda.synchronize()
        ms.append(round(e0.elapsed_time(e1), 3))
    ok = bool(torch.equal(k, want.values)) and bool(torch.equal(v.to(torch.int64), want.indices))
    print({"runs": runs, "ms": ms, "identical_to_stable_sort": ok}, flush=True)
