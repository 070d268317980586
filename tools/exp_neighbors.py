"""Neighbour-search experiments (not a benchmark): times cs_domain_find_neighbors for every (kernel variant, group
policy) pair on the bench workload and checks that all variants return identical lists and counts.

    python tools/exp_neighbors.py [n] [--config uniform|morton] [--only K,G]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cstone_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("n", nargs="?", type=int, default=64 * 1024 * 1024)
ap.add_argument("--config", default="uniform")
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--pbc", type=int, default=0)
ap.add_argument("--knob3", default="0")
ap.add_argument("--bucket", type=int, default=0, help="focus bucket size (default: the bench's); 32 / 16 with 64 Mi uniform particles reproduce the leaf occupancy of the 2- / 4-rank bench trees")
args = ap.parse_args()

dev = torch.device("cuda:0")
n = args.n
g = torch.Generator(device=dev)
g.manual_seed(42)
if args.config == "uniform":
    x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
    h = torch.full((n,), bench.h_for(n, bench.NG0), dtype=torch.float64, device=dev)
    ngmax, key, real = bench.NGMAX, "u64", "d"
else:
    x, y, z = (torch.rand(n, dtype=torch.float32, device=dev, generator=g).clamp_(max=0.99999994) for _ in range(3))
    h = torch.full((n,), bench.h_for(n, 300), dtype=torch.float32, device=dev)
    ngmax, key, real = 384, "u32", "f"
bucket = args.bucket or bench.BUCKET
dom = capi.Domain(0, 1, bucket, bucket, 0.5, (0, 1, 0, 1, 0, 1), (args.pbc,) * 3, key=key, real=real,
                  device="cuda:0")
dom.sync(x, y, z, h)
del x, y, z, h
ref_nb = ref_nc = None
combos = [(0, 0), (0, 1)]  # (unused, target groups): full groups over sibling runs, leaf aligned
if args.only:
    combos = [tuple(int(v) for v in args.only.split(","))]
nb = torch.zeros(n * ngmax, dtype=torch.uint32, device=dev)
nc = torch.zeros(n, dtype=torch.uint32, device=dev)
for kern, grp in [(k3, g) for k3 in [int(v) for v in args.knob3.split(',')] for _, g in combos]:
    capi.tuning_set(1, grp)
    ms = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dom.find_neighbors(ngmax, nb, nc)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out = {"kernel": kern, "groups": grp, "n": n, "config": args.config, "ms": [round(m, 3) for m in ms],
           "mean_nc": round(float(nc.to(torch.float64).mean()), 3), "bucket": bucket,
           "leaves": dom.num_focus_leaves}
    if ref_nb is None and len(combos) > 1:
        ref_nb, ref_nc = nb.clone(), nc.clone()
    elif ref_nb is not None:
        out["identical_to_first"] = bool(torch.equal(nb, ref_nb)) and bool(torch.equal(nc, ref_nc))
    print(json.dumps(out), flush=True)
    nb.zero_()
