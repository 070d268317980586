"""how many leaves hold a particle outside their box / stick out of an ancestor's box (diagnostic for neighbors.cu)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cstone_b200 import capi  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
real = sys.argv[2] if len(sys.argv) > 2 else "f"
dt = torch.float32 if real == "f" else torch.float64
g = torch.Generator(device=dev)
g.manual_seed(42)
x, y, z = (torch.rand(n, dtype=dt, device=dev, generator=g) for _ in range(3))
h = torch.full((n,), bench.h_for(n, 100), dtype=dt, device=dev)
dom = capi.Domain(0, 1, 64, 64, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0), key="u32" if real == "f" else "u64", real=real,
                  device="cuda:0")
dom.sync(x, y, z, h)
print("box", dom.box)
X = [dom.field(k) for k in ("x", "y", "z")]
layout = dom.field("layout").to(torch.int64)
co = dom.field("child_offsets").to(torch.int64)
i2l = dom.field("internal_to_leaf").to(torch.int64)
par = dom.field("parents").to(torch.int64)
c = dom.field("geo_centers").view(-1, 3)
s = dom.field("geo_sizes").view(-1, 3)
nn = co.numel()
leaf_nodes = torch.nonzero(co == 0).flatten()
leaf_of_node = i2l[leaf_nodes]
node_of_leaf = torch.empty(leaf_of_node.numel(), dtype=torch.int64, device=dev)
node_of_leaf[leaf_of_node] = leaf_nodes
counts = layout[1:] - layout[:-1]
pl = torch.repeat_interleave(torch.arange(counts.numel(), device=dev), counts)
pn = node_of_leaf[pl]
over = torch.zeros(pl.numel(), dtype=dt, device=dev)
for d in range(3):
    over = torch.maximum(over, (X[d][: pl.numel()] - c[pn, d]).abs() - s[pn, d])
print("particles outside their leaf box:", int((over > 0).sum()), "of", pl.numel(), "max", float(over.max()))
a = leaf_nodes.clone()
worst = torch.zeros(leaf_nodes.numel(), dtype=dt, device=dev)
while bool((a != 0).any()):
    a = torch.where(a != 0, par[(a - 1).clamp(min=0) >> 3], a)
    for d in range(3):
        stick = (c[leaf_nodes, d] - c[a, d]).abs() + s[leaf_nodes, d] - s[a, d]
        worst = torch.maximum(worst, stick)
eps = 2.0 ** -23 if real == "f" else 2.0 ** -52
cabs = max(abs(v) for v in dom.box)
print("leaves sticking out of an ancestor: >0:", int((worst > 0).sum()), "> 4 eps cabs:", int((worst > 4 * eps * cabs).sum()),
      "of", leaf_nodes.numel(), "max/eps/cabs", float(worst.max()) / eps / cabs)
