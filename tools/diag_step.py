"""Per-call host wall time and device event time of the bench step (diagnostic, not a benchmark)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cstone_b200 import capi  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024 * 1024
g = torch.Generator(device=dev)
g.manual_seed(42)
x, y, z = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) for _ in range(3))
h = torch.full((n,), bench.h_for(n, bench.NG0), dtype=torch.float64, device=dev)
dom = capi.Domain(0, 1, bench.BUCKET, bench.BUCKET, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0), device="cuda:0")
nb = torch.empty(n * bench.NGMAX, dtype=torch.uint32, device=dev)
nc = torch.empty(n, dtype=torch.uint32, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

for it in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e = [ev() for _ in range(4)]
    e[0].record()
    dom.reset()
    t1 = time.perf_counter()
    e[1].record()
    dom.sync(x, y, z, h)
    t2 = time.perf_counter()
    e[2].record()
    dom.find_neighbors(bench.NGMAX, nb, nc)
    t3 = time.perf_counter()
    e[3].record()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"it{it} host ms: reset {1e3*(t1-t0):.2f} sync {1e3*(t2-t1):.2f} nb-call {1e3*(t3-t2):.2f} "
          f"drain {1e3*(t4-t3):.2f} | dev ms: reset {e[0].elapsed_time(e[1]):.2f} sync {e[1].elapsed_time(e[2]):.2f} "
          f"nb {e[2].elapsed_time(e[3]):.2f} | total host {1e3*(t4-t0):.2f}", flush=True)
