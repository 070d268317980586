"""experiment: onesweep kernel variants (threads x keys/thread) at 64Mi u64 keys + u32 values (not a benchmark)"""
import ctypes as C
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(1)
arrs = [torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3)]
keys0 = torch.zeros(n, dtype=torch.uint64, device=dev)
capi.compute_sfc_keys(arrs[0], arrs[1], arrs[2], keys0, (0, 1, 0, 1, 0, 1), (0, 0, 0))
ref_sorted = None
tmp_bytes = capi.lib().cs_sort_by_key_temp_bytes_u64(C.c_size_t(n))
key_buf = torch.empty_like(keys0)
val_buf = torch.empty(n, dtype=torch.uint32, device=dev)
tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
keys = torch.empty_like(keys0)
vals = torch.empty(n, dtype=torch.uint32, device=dev)
seq = capi.sequence(0, n, dev)

for variant, name in [(5, "512x15"), (6, "384x18"), (0, "512x12"), (1, "256x12"), (2, "256x15"), (3, "384x12"), (4, "256x9"), (103, "384x12 NO LOOKBACK (wrong result, timing only)"), (100, "512x12 NO LOOKBACK")]:
    capi.lib().cs_sort_set_variant(C.c_int(variant))
    ms = []
    for _ in range(4):
        keys.copy_(keys0)
        vals.copy_(seq)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi._check(capi.lib().cs_sort_by_key_u64(capi._ptr(keys), capi._ptr(vals), C.c_size_t(n), capi._ptr(key_buf),
                                                  capi._ptr(val_buf), capi._ptr(tmp), C.c_size_t(tmp_bytes),
                                                  capi._stream()), "sort")
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = statistics.median(ms[1:])
    if ref_sorted is None:
        ref_sorted = (keys.clone(), vals.clone())
    ok = bool((keys.view(torch.int64) == ref_sorted[0].view(torch.int64)).all()) and \
        bool((vals.view(torch.int32) == ref_sorted[1].view(torch.int32)).all())
    print(f"variant {variant} {name}: {t:7.3f} ms  {200.0 * n / t / 1e6:7.1f} GB/s  same_result={ok}", flush=True)
