"""One warm-up pass of the hot path (cold Domain::sync + findNeighbors), then one pass inside
cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly one launch of every kernel at the
benchmark size.  Not a benchmark: numbers printed under a profiler are never reported."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cstone_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=64 * 1024 * 1024)
ap.add_argument("--steady", action="store_true", help="profile a steady-state sync instead of a cold one")
args = ap.parse_args()

dev = torch.device("cuda:0")
n = args.n
g = torch.Generator(device=dev)
g.manual_seed(42)
x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
h = torch.full((n,), bench.h_for(n, bench.NG0), dtype=torch.float64, device=dev)
dom = capi.Domain(0, 1, bench.BUCKET, bench.BUCKET, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0), device="cuda:0")
nb = torch.empty(n * bench.NGMAX, dtype=torch.uint32, device=dev)
nc = torch.empty(n, dtype=torch.uint32, device=dev)


def step():
    if args.steady:
        dom.sync()
    else:
        dom.reset()
        dom.sync(x, y, z, h)
    dom.find_neighbors(bench.NGMAX, nb, nc)


dom.sync(x, y, z, h)
step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step at n =", n)
