"""debug: GPU halo discovery vs the reference on a Plummer sphere (own range = first quarter of the leaves)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs  # noqa: E402
import bench  # noqa: E402
from cstone_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
x, y, z = _libs.ref_plummer(n, np.float64)
h = np.full(n, 0.01)
lim, bnd = (-1, 1, -1, 1, -1, 1), (0, 0, 0)
dev = torch.device("cuda:0")
dom = capi.Domain(0, 1, 64, 64, 0.5, lim, bnd, key="u64", real="d", device="cuda:0")
dom.sync(*(torch.from_numpy(a).to(dev) for a in (x, y, z, h)))
tree, halos, first, last = bench.halo_discovery_quarter(capi, torch, dom, bnd)
flags = halos().cpu().numpy()
H = lambda t: t.cpu().numpy()  # noqa: E731
nl = dom.num_focus_leaves
cen, siz = H(dom.field("geo_centers")), H(dom.field("geo_sizes"))
layout = H(dom.field("layout"))
leaves = H(dom.field("focus_leaves"))
sx, sy, sz, sh = (H(dom.field(k)) for k in ("x", "y", "z", "h"))
to = dict(prefixes=H(tree.prefixes), childOffsets=H(tree.child_offsets)[:tree.num_nodes], parents=H(tree.parents),
          numNodes=tree.num_nodes)
l2i = H(tree.leaf_to_internal)[tree.num_internal:]
init = cen[l2i].copy()
sc_g, ss_g = capi.compute_bounding_boxes(*(dom.field(k) for k in ("x", "y", "z", "h")), dom.field("layout"), first,
                                         last, 2.0, torch.from_numpy(init).to(dev))
sc_g, ss_g = H(sc_g), H(ss_g)
for name, chk in (("ref", _libs.ref()), ("oracle", _libs.oracle())):
    sc, ss = chk.bounding_boxes("u64d", sx, sy, sz, sh, layout, first, last, 2.0, init)
    print(name, "search boxes equal:", np.array_equal(sc, sc_g), np.array_equal(ss, ss_g))
    fl = chk.find_halos("u64d", to, cen, siz, leaves, sc, ss, dom.box, bnd, first, last)
    diff = np.nonzero(fl != flags)[0]
    print(name, "flags: ref", int(fl.sum()), "gpu", int(flags.sum()), "differing nodes", diff.size, diff[:10])
    if diff.size:
        for d in diff[:5]:
            print("  node", d, "ref", fl[d], "gpu", flags[d], "childOffset", to["childOffsets"][d], "prefix", hex(to["prefixes"][d]),
                  "center", cen[d], "size", siz[d])
print("box", dom.box, "leaves", nl, "first/last", first, last)
