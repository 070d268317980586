"""One warm-up pass of the hot path, then one pass inside cudaProfilerStart/Stop so that
`ncu --profile-from-start off` captures exactly one launch of every kernel at the benchmark size.
Not a benchmark: numbers printed under a profiler are never reported."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=64 * 1024 * 1024)
args = ap.parse_args()

dev = torch.device("cuda:0")
n = args.n
g = torch.Generator(device=dev)
g.manual_seed(42)
x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
h = torch.full((n,), bench.h_for(n, bench.NG0), dtype=torch.float64, device=dev)
hp = bench.HotPath(n, dev)
hp.step(x, y, z, h)
torch.cuda.synchronize()
torch.cuda.profiler.start()
hp.step(x, y, z, h)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step at n =", n)
