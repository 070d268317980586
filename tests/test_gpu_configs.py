"""BASELINE.json configs[2] (Plummer sphere: Domain::sync + halo discovery) and configs[4] (32-bit Morton / float
neighbour-search stress) through tools/run_configs.py at reduced sizes; the full-size runs are recorded in
profiles/r1_configs_2_and_4.jsonl.  The checks are size independent: sorted keys, particle conservation, bucket limit,
no halo flag inside the own range, neighbour lists equal to brute force in the reference's float expression."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_plummer_sync_and_halo_discovery():
    import run_configs

    out = run_configs.config_plummer(1 << 21, 64)
    assert all(out["checks"].values()), out
    assert out["max_leaf_level"] >= 9, "the Plummer tree must be deeper than a uniform one of the same size"


def test_morton_float_neighbor_stress():
    import run_configs

    out = run_configs.config_morton(1 << 20, 300, 384, 64, 32)
    assert all(out["checks"].values()), out
    assert 250 < out["mean_neighbors"] < 330


# ---- the same two workloads against the unmodified reference (oracle/_ref) at 4 Mi particles: every result array is
# compared through position-weighted 64-bit digests (bench.py:wsum / oracle/ref_api.cpp:weightedSum), since 4 Mi x 384
# neighbour lists are not held twice.  The Plummer particles come from the reference's own generator
# (test/coord_samples/plummer.hpp:15-78 through ref_api.cpp).
def _ref_or_skip():
    import _libs
    if _libs.ref_lib() is None:
        pytest.skip("needs oracle/_ref (built where /root/reference exists)")
    return _libs


def test_plummer_4mi_bit_identical_vs_reference():
    import numpy as np
    import torch

    import bench
    from cstone_b200 import capi

    _libs = _ref_or_skip()
    n = 4 * 1024 * 1024
    x, y, z = _libs.ref_plummer(n, np.float64)
    h = np.full(n, 0.01)
    lim, bnd = (-1, 1, -1, 1, -1, 1), (0, 0, 0)
    want = _libs.ref_bench_run("u64d", 1, 64, 64, 0.5, lim, bnd, x, y, z, h, [0, n], halo_quarter=True)[0]["digest"]
    dev = torch.device("cuda:0")
    dom = capi.Domain(0, 1, 64, 64, 0.5, lim, bnd, key="u64", real="d", device="cuda:0")
    dom.sync(*(torch.from_numpy(a).to(dev) for a in (x, y, z, h)))
    got = bench.domain_digest(dom)
    _, halos, _, _ = bench.halo_discovery_quarter(capi, torch, dom, bnd)
    flags = halos()
    got["halo_flags"], got["halo_count"] = bench.wsum(flags), int(flags.sum())
    cmp_ = bench.compare_digests(got, want)
    assert cmp_["identical"], cmp_
    assert {"keys", "x", "leaves", "layout", "prefixes", "centers", "halo_flags"} <= set(cmp_["arrays_compared"])
    assert got["halo_count"] > 0 and dom.num_focus_leaves > n // 64


def test_morton_float_4mi_bit_identical_vs_reference():
    import numpy as np
    import torch

    import bench
    from cstone_b200 import capi

    _libs = _ref_or_skip()
    n, ngmax = 4 * 1024 * 1024, 384
    x, y, z, h = bench.make_particles("morton", n, 7)
    lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
    want = _libs.ref_bench_tree_neighbors("u32f", 1, x, y, z, h, 64, lim, bnd, ngmax)["digest"]
    dev = torch.device("cuda:0")
    dx, dy, dz, dh = (torch.from_numpy(a).to(dev) for a in (x, y, z, h))
    keys = torch.zeros(n, dtype=torch.uint32, device=dev)
    capi.compute_sfc_keys(dx, dy, dz, keys, lim, bnd, kind=1)
    order = capi.sequence(0, n, dev)
    capi.sort_by_key(keys, order)
    sx, sy, sz, sh = capi.gather_arrays4(order, [dx, dy, dz, dh])
    leaves, counts = capi.compute_octree(keys, 64)
    tree = capi.Octree(leaves)
    cen, siz = capi.compute_geo_centers(tree.prefixes, torch.float32, lim, bnd, kind=1)
    layout = capi.exclusive_scan(torch.cat([counts, torch.zeros(1, dtype=torch.uint32, device=dev)]))
    nb, nc = capi.find_neighbors(sx, sy, sz, sh, 0, n, lim, bnd, tree, layout, cen, siz, ngmax)
    got = {"keys": bench.wsum(keys), "x": bench.wsum(sx), "y": bench.wsum(sy), "z": bench.wsum(sz),
           "h": bench.wsum(sh), "leaves": bench.wsum(leaves), "num_leaves": tree.num_leaves,
           "layout": bench.wsum(layout), "num_nodes": tree.num_nodes, "prefixes": bench.wsum(tree.prefixes),
           "child_offsets": bench.wsum(tree.child_offsets[: tree.num_nodes]), "centers": bench.wsum(cen),
           "sizes": bench.wsum(siz), "leaf_counts": bench.wsum(counts)}
    got["nc_sum"], got["lists"] = bench.list_digest(nb.reshape(-1), nc, ngmax)
    cmp_ = bench.compare_digests(got, want)
    assert cmp_["identical"], cmp_
    assert 250 < got["nc_sum"] / n < 330


def test_digests_detect_a_single_changed_entry():
    """the digest comparison is not vacuous: one swapped neighbour / one flipped key bit changes it"""
    import torch

    import bench

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    nb = torch.randint(0, 1 << 20, (1000 * 16,), device=dev, generator=g).to(torch.uint32)
    nc = torch.randint(0, 20, (1000,), device=dev, generator=g).to(torch.uint32)
    a = bench.list_digest(nb, nc, 16)
    nb2 = nb.clone()
    row = int(torch.nonzero(nc.to(torch.int64) >= 2)[0])
    nb2[row * 16], nb2[row * 16 + 1] = nb[row * 16 + 1], nb[row * 16]
    if int(nb[row * 16]) != int(nb[row * 16 + 1]):
        assert bench.list_digest(nb2, nc, 16) != a
    keys = torch.arange(1, 5000, device=dev, dtype=torch.int64).view(torch.uint64)
    k2 = keys.clone().view(torch.int64)
    k2[77] ^= 1
    assert bench.wsum(keys) != bench.wsum(k2.view(torch.uint64))


@pytest.mark.parametrize("bucket_focus,pbc,force_group", [(8, 0, False), (16, 1, False), (64, 0, True)])
def test_group_steered_search_2mi_bit_identical_vs_reference(bucket_focus, pbc, force_group):
    """findNeighbors on trees with small leaves takes the group-steered search (csrc/neighbors.cu: no per-lane walk
    state, contiguous particle ranges staged across leaf boundaries, subtrees taken whole).  Its lists must still be
    the reference's bit for bit, including the particles that lie up to one cell of the key grid outside the box of
    their leaf (about 250 of 4 Mi with 64-bit keys; the reference only finds those from targets whose own walk enters
    the leaf): compared with the unmodified reference through the list digests on 2 Mi uniform particles."""
    import numpy as np
    import torch

    import bench
    from cstone_b200 import capi

    _libs = _ref_or_skip()
    n, ngmax = 2 * 1024 * 1024, 150
    x, y, z, h = bench.make_particles("uniform", n, 5)
    lim, bnd = (0, 1, 0, 1, 0, 1), (pbc, pbc, pbc)
    want = _libs.ref_bench_run("u64d", 1, 64, bucket_focus, 0.5, lim, bnd, x, y, z, h, [0, n], ngmax=ngmax)[0]["digest"]
    dev = torch.device("cuda:0")
    dom = capi.Domain(0, 1, 64, bucket_focus, 0.5, lim, bnd, key="u64", real="d", device="cuda:0")
    dom.sync(*(torch.from_numpy(a).to(dev) for a in (x, y, z, h)))
    got = bench.domain_digest(dom)
    capi.tuning_set(2, 2 if force_group else 0)
    try:
        nb, nc = dom.find_neighbors(ngmax)
        got["nc_sum"], got["lists"] = bench.list_digest(nb.reshape(-1), nc, ngmax)
        # the per-lane search returns the same lists
        capi.tuning_set(2, 1)
        nb2, nc2 = dom.find_neighbors(ngmax)
        assert torch.equal(nc, nc2) and torch.equal(nb, nb2)
    finally:
        capi.tuning_set(2, 0)
    cmp_ = bench.compare_digests(got, want)
    assert cmp_["identical"], cmp_
    assert {"lists", "nc_sum", "layout", "centers"} <= set(cmp_["arrays_compared"])
    assert 80 < got["nc_sum"] / n < 120
    if not force_group:
        assert n / dom.num_focus_leaves < 20, "the test is meant to run on a tree with small leaves"
    dom.close()
