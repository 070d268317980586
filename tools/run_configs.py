"""BASELINE.json configs[2] and configs[4] at full size on one B200, with size-independent checks (not bench lines:
bench.py measures configs[1] / configs[3]).  Prints one JSON line per config.

  configs[2]: 64 Mi Plummer-sphere particles (clustered, deep tree), 64-bit Hilbert, double: Domain::sync + halo discovery
              (halo discovery timed standalone over the first quarter of the leaves as own range, like the reference's
              test/performance/octree.cu:110-143 - one rank has no foreign range of its own)
  configs[4]: 16 Mi uniform particles, 32-bit Morton keys, float, <ng> ~ 300, ngmax 384: standalone tree build +
              findNeighbors, checked against brute force on sampled targets
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
from cstone_b200 import capi  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=3):
    ms = []
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2], out


def plummer(n, gen):
    """test/coord_samples/plummer.hpp:15-78 (radius cut at 100, scaled by 3 pi / 16, centred)"""
    parts = []
    need = n
    while need > 0:
        m = int(need * 1.01) + 1024
        u = torch.rand(m, dtype=torch.float64, device=DEV, generator=gen).clamp_min_(1e-300)
        r = 1.0 / torch.sqrt(u.pow(-2.0 / 3.0) - 1.0)
        r = r[r < 100.0][:need]
        parts.append(r)
        need -= r.numel()
    r = torch.cat(parts)
    zc = (1.0 - 2.0 * torch.rand(n, dtype=torch.float64, device=DEV, generator=gen)) * r
    th = 2 * math.pi * torch.rand(n, dtype=torch.float64, device=DEV, generator=gen)
    rho = torch.sqrt(torch.clamp(r * r - zc * zc, min=0.0))
    conv = 3.0 * math.pi / 16.0
    x, y, z = rho * torch.cos(th) * conv, rho * torch.sin(th) * conv, zc * conv
    return x - x.mean(), y - y.mean(), z - z.mean()


def config_plummer(n, bucket):
    gen = torch.Generator(device=DEV)
    gen.manual_seed(42)
    x, y, z = plummer(n, gen)
    h = torch.full((n,), 0.01, dtype=torch.float64, device=DEV)
    lim, bnd = (-1, 1, -1, 1, -1, 1), (0, 0, 0)   # open box: grows to the particle extent in the first sync
    dom = capi.Domain(0, 1, bucket, bucket, 0.5, lim, bnd, key="u64", real="d", device="cuda:0")

    def cold():
        dom.reset()
        dom.sync(x, y, z, h)

    cold()
    t_cold, _ = timed(cold)
    t_steady, _ = timed(dom.sync)
    keys = dom.field("keys").view(torch.int64)
    counts = dom.field("focus_leaf_counts").to(torch.int64)
    leaves = dom.field("focus_leaves").view(torch.int64)
    level = (63 - torch.log2((leaves[1:] - leaves[:-1]).to(torch.float64))) / 3
    checks = {
        "keys_sorted": bool((keys[1:] >= keys[:-1]).all()),
        "counts_sum_equals_n": int(counts.sum()) == n,
        "leaf_counts_within_bucket_or_max_depth": bool(((counts <= bucket) | (level >= 21)).all()),
        "coordinates_are_a_permutation": abs(float(dom.field("x").sum() - x.sum())) < 1e-6 * n,
    }
    # standalone halo discovery on the synchronised tree, own range = first quarter of the leaves
    nl = dom.num_focus_leaves
    tree = capi.Octree(dom.field("focus_leaves").clone())
    cen, siz = dom.field("geo_centers"), dom.field("geo_sizes")
    layout = dom.field("layout")
    first, last = 0, nl // 4
    sx, sy, sz, sh = (dom.field(k) for k in ("x", "y", "z", "h"))
    init = cen[tree.leaf_to_internal[tree.num_internal:].long()].contiguous()

    def halos():
        sc, ss = capi.compute_bounding_boxes(sx, sy, sz, sh, layout, first, last, 2.0, init)
        return capi.find_halos(tree, cen, siz, sc, ss, dom.box, bnd, first, last)

    t_halo, flags = timed(halos)
    own_nodes = tree.leaf_to_internal[tree.num_internal:][first:last].long()
    checks["no_halo_flag_inside_own_range"] = int(flags[own_nodes].sum()) == 0
    checks["some_halos_found"] = int(flags.sum()) > 0
    return {"config": "configs[2]: Plummer sphere, 64-bit Hilbert, double, Domain::sync + halo discovery", "n": n,
            "bucket": bucket, "focus_leaves": nl, "max_leaf_level": int(level.max()),
            "particles_per_leaf": round(n / nl, 2), "ms_sync_cold": round(t_cold, 3),
            "ms_sync_steady": round(t_steady, 3), "ms_halo_discovery_quarter_range": round(t_halo, 3),
            "halo_nodes": int(flags.sum()), "Mparticles_per_s_cold_sync": round(n / t_cold / 1e3, 1), "checks": checks}


def config_morton(n, ng, ngmax, bucket, samples):
    gen = torch.Generator(device=DEV)
    gen.manual_seed(42)
    x, y, z = (torch.rand(n, dtype=torch.float32, device=DEV, generator=gen).clamp_(max=0.99999994) for _ in range(3))
    hval = 0.5 * (3.0 * ng / (4 * math.pi * n)) ** (1.0 / 3.0)
    h = torch.full((n,), hval, dtype=torch.float32, device=DEV)
    lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
    keys = torch.zeros(n, dtype=torch.uint32, device=DEV)
    state = {}

    def build():
        capi.compute_sfc_keys(x, y, z, keys, lim, bnd, kind=1)
        order = capi.sequence(0, n, DEV)
        capi.sort_by_key(keys, order)
        sx, sy, sz, sh = capi.gather_arrays4(order, [x, y, z, h])
        leaves, counts = capi.compute_octree(keys, bucket)
        tree = capi.Octree(leaves)
        cen, siz = capi.compute_geo_centers(tree.prefixes, torch.float32, lim, bnd, kind=1)
        layout = capi.exclusive_scan(torch.cat([counts, torch.zeros(1, dtype=torch.uint32, device=DEV)]))
        state.update(sx=sx, sy=sy, sz=sz, sh=sh, tree=tree, cen=cen, siz=siz, layout=layout, counts=counts,
                     order=order)

    t_build, _ = timed(build)
    nb = torch.empty(n * ngmax, dtype=torch.uint32, device=DEV)
    nc = torch.empty(n, dtype=torch.uint32, device=DEV)
    s = state

    def search():
        capi.find_neighbors(s["sx"], s["sy"], s["sz"], s["sh"], 0, n, lim, bnd, s["tree"], s["layout"], s["cen"],
                            s["siz"], ngmax, nb, nc)

    t_nb, _ = timed(search)
    # brute force on sampled targets with the reference's expression (x^2 + y^2) + z^2 < 4 h^2 in float, no FMA
    sel = torch.randint(0, n, (samples,), device=DEV, generator=gen)
    ok_counts, ok_lists = True, True
    nbv = nb.view(n, ngmax)
    r2 = torch.tensor(4.0, dtype=torch.float32, device=DEV) * h[0] * h[0]
    for i in sel.tolist():
        dx, dy, dz = s["sx"] - s["sx"][i], s["sy"] - s["sy"][i], s["sz"] - s["sz"][i]
        d2 = (dx * dx + dy * dy) + dz * dz
        idx = torch.nonzero(d2 < r2).flatten()
        idx = idx[idx != i]
        ok_counts &= int(nc[i]) == idx.numel()
        m = min(idx.numel(), ngmax)
        ok_lists &= bool((nbv[i, :m].long() == idx[:m]).all())
    ksort = keys.view(torch.int32)
    checks = {"keys_sorted": bool((ksort[1:] >= ksort[:-1]).all()),
              "counts_sum_equals_n": int(s["counts"].to(torch.int64).sum()) == n,
              "neighbor_counts_equal_brute_force": ok_counts, "neighbor_lists_equal_brute_force": ok_lists}
    mean_nc = float(nc.to(torch.float64).mean())
    return {"config": "configs[4]: uniform, 32-bit Morton, float, neighbour-search stress", "n": n, "ngmax": ngmax,
            "mean_neighbors": round(mean_nc, 2), "leaves": s["tree"].num_leaves, "ms_build": round(t_build, 3),
            "ms_find_neighbors": round(t_nb, 3), "Mparticles_per_s_find_neighbors": round(n / t_nb / 1e3, 1),
            "brute_force_samples": samples, "checks": checks}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-plummer", type=int, default=64 * 1024 * 1024)
    ap.add_argument("--n-morton", type=int, default=16 * 1024 * 1024)
    ap.add_argument("--samples", type=int, default=64)
    args = ap.parse_args()
    print(json.dumps(config_plummer(args.n_plummer, 64)), flush=True)
    torch.cuda.empty_cache()
    print(json.dumps(config_morton(args.n_morton, 300, 384, 64, args.samples)), flush=True)
