"""run one 64Mi (u64 key, u32 value) sort with a given kernel variant (for ncu)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 64 * 1024 * 1024
dev = torch.device("cuda:0")
keys = torch.randint(0, 2 ** 62, (n,), dtype=torch.int64, device=dev).view(torch.uint64)
vals = capi.sequence(0, n, dev)
capi.lib().cs_sort_set_variant(C.c_int(variant))
capi.sort_by_key(keys, vals)
torch.cuda.synchronize()
print("sorted", bool((keys.view(torch.int64)[1:] >= keys.view(torch.int64)[:-1]).all()))
