"""experiment: effect of cudaLimitMaxL2FetchGranularity on the random gather and on the sort (not a benchmark)"""
import ctypes as C
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

n = 64 * 1024 * 1024
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)
arrs = [torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(4)]
keys = torch.zeros(n, dtype=torch.uint64, device=dev)
capi.compute_sfc_keys(arrs[0], arrs[1], arrs[2], keys, (0, 1, 0, 1, 0, 1), (0, 0, 0))
unsorted = keys.clone()
order = capi.sequence(0, n, dev)
capi.sort_by_key(keys, order)


def timeit(fn, reps=4):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms[1:])


for gran in (128, 64, 32, 128):
    capi._check(capi.lib().cs_set_l2_fetch_granularity(C.c_int(gran)), "granularity")
    t_g = timeit(lambda: capi.gather4(order, arrs))

    def sort_once():
        k = unsorted.clone()
        o = capi.sequence(0, n, dev)
        capi.sort_by_key(k, o)

    t_s = timeit(sort_once)
    print(f"l2 fetch granularity {gran:4d} B: gather4 {t_g:7.3f} ms   clone+sequence+sort {t_s:7.3f} ms", flush=True)
