"""GPU parity tests: every stage of the hot path, called through the C ABI (cstone_b200.capi -> libcstone_b200.so),
compared bit-for-bit with the oracle (oracle/liboracle.so) and, when present, with the unmodified reference
(oracle/_ref/libcstone_ref.so) on the same seeded inputs.  Run with `pytest -m gpu` on the B200 box."""
import numpy as np
import pytest
import torch

from _libs import COMBOS, KEYS, MAXLEVEL, key_of, oracle, real_of, ref
from _util import const_h, gaussian_particles, plummer_particles, uniform_particles

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def capi():
    from cstone_b200 import capi as c
    return c


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.cpu().numpy()


def checkers():
    out = [("oracle", oracle())]
    if ref() is not None:
        out.append(("reference", ref()))
    return out


BOXES = [
    ((0, 1, 0, 1, 0, 1), (0, 0, 0)),
    ((0, 1, 0, 1, 0, 1), (1, 1, 1)),
    ((-1.2, 1.3, -0.4, 2.2, -3.0, 1.0), (1, 0, 1)),
]


def particles(T, n, lim, seed=7):
    rng = np.random.default_rng(seed)
    out = []
    for d in range(3):
        lo, hi = lim[2 * d], lim[2 * d + 1]
        a = (lo + (hi - lo) * rng.random(n)).astype(T)
        np.clip(a, T(lo), np.nextafter(T(hi), T(lo)), out=a)
        out.append(a)
    return out


# ------------------------------------------------------------------------------------------------ keys
@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("lim,bnd", BOXES)
def test_sfc_keys(combo, kind, lim, bnd):
    T, K = real_of(combo), KEYS[key_of(combo)]
    for n in (0, 1, 3, 1000, 100003):
        x, y, z = particles(T, n + 1, lim)
        # edge coordinates: box corners
        if n >= 3:
            x[0], y[0], z[0] = T(lim[0]), T(lim[2]), T(lim[4])
            x[1], y[1], z[1] = (np.nextafter(T(lim[1]), T(lim[0])), np.nextafter(T(lim[3]), T(lim[2])),
                                np.nextafter(T(lim[5]), T(lim[4])))
        for off in (0, 1):  # off=1: pointers not 16-byte aligned -> scalar path
            xs, ys, zs = x[off:off + n], y[off:off + n], z[off:off + n]
            pre = np.zeros(n + 1, dtype=K)
            remove = K(1) << K(3 * MAXLEVEL[key_of(combo)])
            pre[::5] = remove
            dk = dev(pre)
            capi().compute_sfc_keys(dev(x)[off:off + n], dev(y)[off:off + n], dev(z)[off:off + n], dk[off:off + n],
                                    lim, bnd, kind=kind, n=n)
            got = host(dk)[off:off + n]
            for name, chk in checkers():
                want = chk.sfc_keys(combo, kind, xs.copy(), ys.copy(), zs.copy(), lim, bnd,
                                    keys=pre[off:off + n].copy())
                assert np.array_equal(got, want), (name, n, off)


@pytest.mark.parametrize("combo", COMBOS)
def test_hilbert_lut_exhaustive_low_levels(combo):
    """all 2^15 cells of a 32^3 grid placed at several scales: exercises every state of the 3-level lookup tables"""
    T = real_of(combo)
    g = np.arange(32)
    gx, gy, gz = [a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")]
    L = MAXLEVEL[key_of(combo)]
    for shift in (0, 3, L - 5):
        scale = float(1 << shift) / float(1 << L)
        x, y, z = ((a + 0.5 if shift else a + 0.0) * scale for a in (gx, gy, gz))
        x, y, z = x.astype(T), y.astype(T), z.astype(T)
        lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
        dk = torch.zeros(x.size, dtype=getattr(torch, "uint" + key_of(combo)[1:]), device=DEV)
        capi().compute_sfc_keys(dev(x), dev(y), dev(z), dk, lim, bnd)
        want = oracle().sfc_keys(combo, 0, x, y, z, lim, bnd)
        assert np.array_equal(host(dk), want)


# ------------------------------------------------------------------------------------------------ sort
@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("n", [0, 1, 2, 31, 1000, 6143, 6144, 6145, 8192, 8193, 50000, 1 << 20])
def test_sort_by_key(kt, n):
    rng = np.random.default_rng(n + 1)
    K = KEYS[kt]
    hi = (1 << (3 * MAXLEVEL[kt])) + 1
    for span in (hi, 50):  # span 50: heavy duplication => checks stability
        keys = rng.integers(0, span, n, dtype=np.uint64).astype(K)
        dk, dv = dev(keys), capi().sequence(0, n, DEV) if n else torch.zeros(0, dtype=torch.uint32, device=DEV)
        capi().sort_by_key(dk, dv)
        ko, vo = keys.copy(), np.arange(n, dtype=np.uint32)
        oracle().sort_by_key(kt, ko, vo)
        assert np.array_equal(host(dk), ko)
        assert np.array_equal(host(dv), vo)
    # keys only, with a non-zero sequence start
    keys = rng.integers(0, hi, n, dtype=np.uint64).astype(K)
    dk = dev(keys)
    capi().sort_by_key(dk, None)
    assert np.array_equal(host(dk), np.sort(keys, kind="stable"))
    if n:
        assert np.array_equal(host(capi().sequence(7, n, DEV)), np.arange(7, 7 + n, dtype=np.uint32))


def test_sort_top_bit_and_extremes():
    """removeKey (2^63) and all-ones keys sort to the end; already-sorted and reversed inputs"""
    n = 20000
    keys = np.arange(n, dtype=np.uint64)[::-1].copy()
    keys[::97] = np.uint64(1) << np.uint64(63)
    keys[5] = np.uint64(0xFFFFFFFFFFFFFFFF)
    dk, dv = dev(keys), capi().sequence(0, n, DEV)
    capi().sort_by_key(dk, dv)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(host(dk), keys[order])
    assert np.array_equal(host(dv), order.astype(np.uint32))


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_gather_and_scan(T):
    rng = np.random.default_rng(3)
    n = 100001
    order = rng.permutation(n).astype(np.uint32)
    arrs = [rng.random(n).astype(T) for _ in range(4)]
    do = dev(order)
    got = capi().gather4(do, [dev(a) for a in arrs])
    for g, a in zip(got, arrs):
        assert np.array_equal(host(g), a[order])
    assert np.array_equal(host(capi().gather(do, dev(arrs[0]))), arrs[0][order])
    # record-packed gatherArrays: partial orderings (n < srcCount), both the direct (small) and the packed path
    for sub in (order, order[: n // 3], order[:7]):
        got = capi().gather_arrays4(dev(sub), [dev(a) for a in arrs])
        for g, a in zip(got, arrs):
            assert np.array_equal(host(g), a[sub])
    small = [a[:1000] for a in arrs]
    so = rng.permutation(1000).astype(np.uint32)
    for g, a in zip(capi().gather_arrays4(dev(so), [dev(a) for a in small]), small):
        assert np.array_equal(host(g), a[so])
    for m in (1, 2, 4095, 4096, 4097, 1 << 21):
        v = rng.integers(0, 5000, m).astype(np.uint32)
        want = np.zeros(m, dtype=np.uint32)
        want[1:] = np.cumsum(v[:-1], dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(host(capi().exclusive_scan(dev(v))), want)


# ------------------------------------------------------------------------------------------------ trees
def sorted_keys(combo, n, dist, seed=11):
    T = real_of(combo)
    lim, bnd = (-1, 1, -1, 1, -1, 1), (0, 0, 0)
    if dist == "gaussian":
        x, y, z = gaussian_particles(n, T, seed)
    elif dist == "plummer":
        x, y, z = plummer_particles(n, T, seed)
        m = max(np.abs(a).max() for a in (x, y, z)) * 1.001
        lim = (-m, m, -m, m, -m, m)
    else:
        x, y, z = uniform_particles(n, T, seed, -1, 1)
    keys = oracle().sfc_keys(combo, 0, x, y, z, lim, bnd)
    order = np.arange(n, dtype=np.uint32)
    oracle().sort_by_key(key_of(combo), keys, order)
    return keys, (x[order], y[order], z[order]), lim, bnd


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("dist", ["uniform", "gaussian", "plummer"])
@pytest.mark.parametrize("bucket", [1, 16, 64])
def test_csarray_and_octree(combo, dist, bucket):
    kt, T = key_of(combo), real_of(combo)
    keys, _, lim, bnd = sorted_keys(combo, 40000, dist)
    dk = dev(keys)
    lo, co = oracle().compute_octree(kt, keys, bucket)
    leaves, counts = capi().compute_octree(dk, bucket)
    assert np.array_equal(host(leaves), lo) and np.array_equal(host(counts), co)

    # stage functions one update step at a time from a coarse tree
    lv = np.array([0, int(lo[-1])], dtype=KEYS[kt])
    ct = np.array([keys.size], dtype=np.uint32)
    for _ in range(3):
        ops_o, conv_o = oracle().rebalance_decision(kt, lv, ct, bucket)
        scan_o = np.zeros(lv.size, dtype=np.int64)
        scan_o[1:] = np.cumsum(ops_o)
        ops, new_n, conv = capi().compute_node_ops(dev(lv), dev(ct), bucket)
        assert np.array_equal(host(ops).astype(np.int64), scan_o) and new_n == scan_o[-1] and conv == conv_o
        new_leaves = capi().rebalance_tree(dev(lv), ops, new_n)
        lv, ct, _ = oracle().update_octree(kt, keys, bucket, lv, ct)
        assert np.array_equal(host(new_leaves), lv)
        assert np.array_equal(host(capi().compute_node_counts(new_leaves, dk)), ct)
    assert np.array_equal(host(capi().compute_node_counts(leaves, dev(keys[::3].copy()), 5)),
                          oracle().compute_node_counts(kt, lo, keys[::3].copy(), 5))
    # empty key set and keys beyond the last leaf
    assert host(capi().compute_node_counts(leaves, dk, n=0)).sum() == 0

    tree = capi().Octree(leaves)
    to = oracle().build_octree(kt, lo)
    nn = to["numNodes"]
    assert np.array_equal(host(tree.prefixes), to["prefixes"])
    assert np.array_equal(host(tree.child_offsets)[:nn], to["childOffsets"])
    assert np.array_equal(host(tree.parents)[:(nn - 1) // 8], to["parents"])
    assert np.array_equal(host(tree.level_range), to["levelRange"])
    assert np.array_equal(host(tree.internal_to_leaf), to["internalToLeaf"])
    assert np.array_equal(host(tree.leaf_to_internal), to["leafToInternal"])

    tdtype = torch.float32 if T == np.float32 else torch.float64
    for kind in (0, 1):
        cen, siz = capi().compute_geo_centers(tree.prefixes, tdtype, lim, bnd, kind=kind)
        cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd, kind=kind)
        assert np.array_equal(host(cen), cen_o) and np.array_equal(host(siz), siz_o)

    # count upsweep (octree_gpu.cu:205-236)
    node_counts = np.zeros(nn, dtype=np.uint32)
    node_counts[to["leafToInternal"][to["numInternal"]:]] = co
    dn = dev(node_counts)
    capi().upsweep_sum(kt, to["levelRange"], tree.child_offsets, dn)
    got = host(dn)
    assert got[0] == keys.size
    internal = to["childOffsets"] > 0
    kids = to["childOffsets"][internal][:, None] + np.arange(8)[None, :]
    assert np.array_equal(got[internal], got[kids].sum(axis=1))


def test_single_leaf_tree():
    leaves = dev(np.array([0, 1 << 63], dtype=np.uint64))
    tree = capi().Octree(leaves)
    assert tree.num_nodes == 1 and host(tree.prefixes)[0] == 1
    assert host(tree.child_offsets)[0] == 0
    assert host(tree.level_range).tolist() == [0] + [1] * 22


# ------------------------------------------------------------------------------------------------ neighbours
@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("pbc", [0, 1])
@pytest.mark.parametrize("dist", ["uniform", "gaussian"])
def test_find_neighbors(combo, pbc, dist):
    kt, T = key_of(combo), real_of(combo)
    n = 20000
    keys, (x, y, z), lim, _ = sorted_keys(combo, n, dist)
    bnd = (pbc, pbc, pbc)
    lo, co = oracle().compute_octree(kt, keys, 16)
    to = oracle().build_octree(kt, lo)
    cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    rng = np.random.default_rng(2)
    h = (const_h(n, 40, T, 8.0) * (0.6 + 0.8 * rng.random(n))).astype(T)

    tree = capi().Octree(dev(lo))
    tdtype = torch.float32 if T == np.float32 else torch.float64
    cen, siz = capi().compute_geo_centers(tree.prefixes, tdtype, lim, bnd)
    dx, dy, dz, dh, dl = dev(x), dev(y), dev(z), dev(h), dev(layout_o)
    for first, last, ngmax in [(0, n, 150), (0, n, 16), (37, n - 45, 64), (100, 101, 8)]:
        nb, nc = capi().find_neighbors(dx, dy, dz, dh, first, last, lim, bnd, tree, dl, cen, siz, ngmax)
        nb, nc = host(nb), host(nc)
        for name, chk in checkers():
            nb_o, nc_o = chk.find_neighbors(combo, x, y, z, h, first, last, lim, bnd, to, lo, layout_o, cen_o, siz_o,
                                            ngmax)
            assert np.array_equal(nc, nc_o), name
            m = np.arange(ngmax)[None, :] < np.minimum(nc_o, ngmax)[:, None]
            assert np.array_equal(nb[m], nb_o[m]), name
    # sorted ascending (H3) and no self entries (H2)
    nb, nc = capi().find_neighbors(dx, dy, dz, dh, 0, n, lim, bnd, tree, dl, cen, siz, 200)
    nb, nc = host(nb).astype(np.int64), host(nc)
    m = np.arange(200)[None, :] < np.minimum(nc, 200)[:, None]
    assert not np.any((nb == np.arange(n)[:, None]) & m)
    d = np.diff(nb, axis=1)
    assert np.all((d > 0) | ~m[:, 1:])


def _neighbors_vs_oracle(combo, x, y, z, h, lim, bnd, bucket, ngmax):
    """sort the particles along the curve, build the tree with the oracle, compare GPU lists with the checkers"""
    kt, T = key_of(combo), real_of(combo)
    n = x.size
    keys = oracle().sfc_keys(combo, 0, x, y, z, lim, bnd)
    order = np.arange(n, dtype=np.uint32)
    oracle().sort_by_key(kt, keys, order)
    x, y, z, h = x[order], y[order], z[order], h[order]
    lo, co = oracle().compute_octree(kt, keys, bucket)
    to = oracle().build_octree(kt, lo)
    cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    tree = capi().Octree(dev(lo))
    tdtype = torch.float32 if T == np.float32 else torch.float64
    cen, siz = capi().compute_geo_centers(tree.prefixes, tdtype, lim, bnd)
    nb, nc = capi().find_neighbors(dev(x), dev(y), dev(z), dev(h), 0, n, lim, bnd, tree, dev(layout_o), cen, siz, ngmax)
    nb, nc = host(nb), host(nc)
    for name, chk in checkers():
        nb_o, nc_o = chk.find_neighbors(combo, x, y, z, h, 0, n, lim, bnd, to, lo, layout_o, cen_o, siz_o, ngmax)
        assert np.array_equal(nc, nc_o), name
        m = np.arange(ngmax)[None, :] < np.minimum(nc_o, ngmax)[:, None]
        assert np.array_equal(nb[m], nb_o[m]), name
    return nc


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("pbc", [0, 1])
def test_find_neighbors_lattice_exact_radius(combo, pbc):
    """adversarial for the certified float pre-filter: lattice particles whose distances are exactly 2h (and exactly
    representable), so d2 == r2 and the strict `<` of findneighbors.hpp:134 must exclude them; plus duplicates"""
    T = real_of(combo)
    g = 24
    c = (np.arange(g, dtype=np.float64) + 0.5) / 32.0
    x, y, z = (a.ravel().astype(T) for a in np.meshgrid(c, c, c, indexing="ij"))
    x, y, z = (np.concatenate([a, a[:50]]) for a in (x, y, z))  # 50 coincident pairs (d2 == 0, kept: H2)
    lim, bnd = (0, 1, 0, 1, 0, 1), (pbc, pbc, pbc)
    for mult in (1.0, 2.0, np.sqrt(2.0), np.sqrt(3.0)):
        h = np.full(x.size, T(0.5 * mult / 32.0), dtype=T)
        nc = _neighbors_vs_oracle(combo, x, y, z, h, lim, bnd, 8, 64)
        assert nc.max() > 0 or mult == 1.0


@pytest.mark.parametrize("combo", COMBOS)
def test_find_neighbors_offset_box_and_clustered(combo):
    """large coordinate magnitudes relative to h (box far from the origin) and a Plummer sphere with per-particle h
    spanning orders of magnitude: the float pre-filter's error bound must hold or defer to the exact expression"""
    T = real_of(combo)
    n = 30000
    rng = np.random.default_rng(5)
    off = 1000.0 if T == np.float32 else 1.0e6
    x, y, z = ((off + rng.random(n)).astype(T) for _ in range(3))
    lim = (off, off + 1, off, off + 1, off, off + 1)
    for a in (x, y, z):
        np.clip(a, T(off), np.nextafter(T(off + 1), T(off)), out=a)
    h = (const_h(n, 50, T) * (0.5 + rng.random(n))).astype(T)
    _neighbors_vs_oracle(combo, x, y, z, h, lim, (0, 0, 0), 16, 100)

    x, y, z = plummer_particles(n, T, 9)
    m = float(max(np.abs(a).max() for a in (x, y, z))) * 1.001
    lim = (-m, m, -m, m, -m, m)
    r = np.sqrt(x.astype(np.float64) ** 2 + y.astype(np.float64) ** 2 + z.astype(np.float64) ** 2)
    h = (0.02 * (0.05 + r) * (0.5 + rng.random(n))).astype(T)
    _neighbors_vs_oracle(combo, x, y, z, h, lim, (0, 0, 0), 8, 128)
    _neighbors_vs_oracle(combo, x, y, z, h, lim, (1, 1, 1), 64, 32)


# ------------------------------------------------------------------------------------------------ halos
@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("pbc", [0, 1])
def test_halo_discovery(combo, pbc):
    kt, T = key_of(combo), real_of(combo)
    n = 30000
    keys, (x, y, z), lim, _ = sorted_keys(combo, n, "gaussian")
    bnd = (pbc, pbc, pbc)
    lo, co = oracle().compute_octree(kt, keys, 8)
    to = oracle().build_octree(kt, lo)
    cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    h = const_h(n, 30, T, 8.0)
    nl = lo.size - 1

    tree = capi().Octree(dev(lo))
    tdtype = torch.float32 if T == np.float32 else torch.float64
    cen, siz = capi().compute_geo_centers(tree.prefixes, tdtype, lim, bnd)
    init = cen_o[to["leafToInternal"][to["numInternal"]:]]
    dx, dy, dz, dh, dl = dev(x), dev(y), dev(z), dev(h), dev(layout_o)
    for first, last in [(0, nl // 4), (nl // 3, 2 * nl // 3), (nl - 5, nl), (0, nl)]:
        sc, ss = capi().compute_bounding_boxes(dx, dy, dz, dh, dl, first, last, 2.0, dev(init))
        sc_o, ss_o = oracle().bounding_boxes(combo, x, y, z, h, layout_o, first, last, 2.0, init)
        assert np.array_equal(host(sc)[first:last], sc_o[first:last])
        assert np.array_equal(host(ss)[first:last], ss_o[first:last])
        flags = capi().find_halos(tree, cen, siz, sc, ss, lim, bnd, first, last)
        for name, chk in checkers():
            want = chk.find_halos(combo, to, cen_o, siz_o, lo, sc_o, ss_o, lim, bnd, first, last)
            assert np.array_equal(host(flags), want), name


# ------------------------------------------------------------------------------------------------ full size
def test_full_size_keys_and_sort_properties():
    """BASELINE config 2 size (64 Mi, u64/double): size-independent properties of keys + sort + gather"""
    n = 64 * 1024 * 1024
    g = torch.Generator(device=DEV)
    g.manual_seed(42)
    x, y, z = (torch.rand(n, dtype=torch.float64, device=DEV, generator=g) for _ in range(3))
    lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
    keys = torch.zeros(n, dtype=torch.uint64, device=DEV)
    capi().compute_sfc_keys(x, y, z, keys, lim, bnd)
    # sample check against the oracle
    idx = torch.randint(0, n, (100000,), device=DEV, generator=g)
    want = oracle().sfc_keys("u64d", 0, host(x[idx]), host(y[idx]), host(z[idx]), lim, bnd)
    assert np.array_equal(host(keys.view(torch.int64)[idx]).view(np.uint64), want)

    unsorted = keys.clone()
    order = capi().sequence(0, n, DEV)
    capi().sort_by_key(keys, order)
    k64 = keys.view(torch.int64)  # keys < 2^63: signed comparison is order preserving
    assert bool((k64[1:] >= k64[:-1]).all()), "keys not sorted"
    o64 = order.view(torch.int32).to(torch.int64)  # n < 2^31
    assert int(o64.sum()) == n * (n - 1) // 2, "values are not a permutation (checksum)"
    assert bool((unsorted.view(torch.int64)[o64] == k64).all()), "values do not follow their keys"
    # stability: among equal keys the original indices ascend
    eq = k64[1:] == k64[:-1]
    assert bool((o64[1:][eq] > o64[:-1][eq]).all())
    xs = capi().gather(order, x)
    assert bool((xs == x[o64]).all())


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("run_sizes", [[5, 0, 7], [4096, 4096], [1, 100000, 3, 0, 77777], [50000] * 8, [12345, 1],
                                       [300000, 200000, 100000]])
def test_merge_sorted_runs_equals_stable_sort(kt, run_sizes):
    """cs_merge_sorted_runs_*: the stable merge of sorted runs is the stable sort of their concatenation (what the
    reference's second sortByKey computes, domain/assignment.hpp:197-201); few distinct keys force ties across runs"""
    import ctypes as C
    from cstone_b200 import capi
    np_t = np.uint32 if kt == "u32" else np.uint64
    torch_t = torch.uint32 if kt == "u32" else torch.uint64
    rng = np.random.default_rng(5)
    runs = [np.sort(rng.integers(0, 1000 if i % 2 else 1 << 28, size=n).astype(np_t)) for i, n in enumerate(run_sizes)]
    keys = np.concatenate(runs) if runs else np.zeros(0, np_t)
    n = keys.size
    vals = rng.permutation(n).astype(np.uint32)
    order = np.argsort(keys, kind="stable")
    offsets = (C.c_size_t * (len(run_sizes) + 1))(*np.concatenate([[0], np.cumsum(run_sizes)]).tolist())
    dk = torch.from_numpy(keys.view(np.int32 if kt == "u32" else np.int64)).to(DEV).view(torch_t)
    dv = torch.from_numpy(vals.view(np.int32)).to(DEV).view(torch.uint32)
    kb, vb = torch.empty_like(dk), torch.empty_like(dv)
    f = getattr(capi.lib(), "cs_merge_sorted_runs_" + kt)
    capi._check(f(capi._ptr(dk), capi._ptr(dv), offsets, C.c_int(len(run_sizes)), capi._ptr(kb), capi._ptr(vb),
                  capi._stream()), "merge")
    torch.cuda.synchronize()
    assert np.array_equal(dk.cpu().view(torch.int32 if kt == "u32" else torch.int64).numpy().view(np_t), keys[order])
    assert np.array_equal(dv.cpu().view(torch.int32).numpy().view(np.uint32), vals[order])


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("first,second,want", [(0, 0, []), (0, 3, []), (0, 4, [3, 4]), (0, 5, [3, 5]), (0, 7, [3, 6]),
                                               (0, 10, [3, 6, 7, 8, 9, 10]), (9, 10, [9, 10])])
def test_extract_marked_elements_reference_vectors(kt, first, second, want):
    """test/unit/domain/layout.cpp:60-102: request keys of the leaves marked as halos within an index range (the device
    version of extractMarkedElements that Halos::exchangeRequests runs per peer)"""
    import ctypes as C
    from cstone_b200 import capi
    np_t = np.uint32 if kt == "u32" else np.uint64
    torch_t = torch.uint32 if kt == "u32" else torch.uint64
    view = np.int32 if kt == "u32" else np.int64
    leaves = torch.from_numpy(np.arange(11, dtype=np_t).view(view)).to(DEV).view(torch_t)
    layout = torch.from_numpy(np.array([0, 0, 0, 0, 1, 2, 3, 3, 4, 4, 5], dtype=np.uint32).view(np.int32)).to(DEV)
    out = torch.zeros(16, dtype=torch.int64 if kt == "u64" else torch.int32, device=DEV)
    f = getattr(capi.lib(), "cs_extract_marked_elements_" + kt)
    f.restype = C.c_long
    n = f(capi._ptr(leaves), capi._ptr(layout), C.c_int(10), C.c_int(first), C.c_int(second), capi._ptr(out),
          C.c_long(16), capi._stream())
    assert n == len(want), capi.lib().cs_last_error()
    assert out[:n].cpu().tolist() == want
    # a buffer that is too small is reported, not overrun
    if want:
        assert f(capi._ptr(leaves), capi._ptr(layout), C.c_int(10), C.c_int(first), C.c_int(second), capi._ptr(out),
                 C.c_long(len(want) - 1), capi._stream()) == -len(want)


@pytest.mark.parametrize("combo", ["u32f", "u64f", "u64d"])
def test_find_halos_flags_21(combo):
    """test/unit/traversal/discovery.cpp:52-103 on the GPU: 16 leaves + 5 internal nodes are flagged from either half of
    the 4x4x4 tree, and the flags equal the oracle's"""
    from cstone_b200 import capi
    from test_oracle_golden import halo_flags_setup

    orc = oracle()
    kt, leaves, tree_o, cen_o, siz_o, sc, ss, lim, bnd = halo_flags_setup(orc, combo)
    T = real_of(combo)
    torch_t = torch.uint32 if kt == "u32" else torch.uint64
    d_leaves = torch.from_numpy(leaves.view(np.int32 if kt == "u32" else np.int64)).to(DEV).view(torch_t)
    tree = capi.Octree(d_leaves)
    cen, siz = capi.compute_geo_centers(tree.prefixes, torch.float32 if T == np.float32 else torch.float64, lim, bnd)
    d_sc, d_ss = torch.from_numpy(sc).to(DEV), torch.from_numpy(ss).to(DEV)
    for first, last in ((0, 32), (32, 64)):
        flags = capi.find_halos(tree, cen, siz, d_sc, d_ss, lim, bnd, first, last).cpu().numpy()
        assert int(flags.sum()) == 21
        want = orc.find_halos(combo, tree_o, cen_o, siz_o, leaves, sc, ss, lim, bnd, first, last)
        assert np.array_equal(flags, want)


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref")
@pytest.mark.parametrize("kind", [0, 1])
def test_sfc_keys_u32_from_double_coordinates(kind):
    """computeSfcKeys<{Hilbert,Morton}Key<unsigned>, double> (sfc/sfc_gpu.cu:46-62) vs the unmodified reference"""
    import ctypes as C

    from _libs import ref_lib
    n = 100003
    rng = np.random.default_rng(11)
    x, y, z = (rng.random(n) * 3.0 - 1.0 for _ in range(3))
    lim, bnd = (-1, 2, -1, 2, -1, 2), (0, 1, 0)
    want = np.zeros(n, dtype=np.uint32)
    lim_a, bnd_a = np.array(lim, dtype=np.float64), np.array(bnd, dtype=np.int32)
    ref_lib().ref_sfc_keys_u32d(C.c_int(kind), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                z.ctypes.data_as(C.c_void_p), want.ctypes.data_as(C.c_void_p), C.c_size_t(n),
                                lim_a.ctypes.data_as(C.c_void_p), bnd_a.ctypes.data_as(C.c_void_p))
    for off in (0, 1):  # aligned (vector path) and misaligned (scalar path) pointers
        dx, dy, dz = (dev(np.concatenate([np.zeros(off), a]))[off:] for a in (x, y, z))
        keys = torch.zeros(n + off, dtype=torch.uint32, device="cuda")[off:]
        capi().compute_sfc_keys(dx, dy, dz, keys, lim, bnd, kind=kind)
        assert np.array_equal(host(keys), want), (kind, off)


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref")
def test_bounding_boxes_double_coordinates_float_h():
    """computeBoundingBoxGpu<double, float> (focus/source_center_gpu.cu:91) vs the unmodified reference"""
    import ctypes as C

    from _libs import ref_lib
    n, nl = 50000, 1000
    rng = np.random.default_rng(3)
    x, y, z = (rng.random(n) for _ in range(3))
    h = (0.01 + 0.02 * rng.random(n)).astype(np.float32)
    layout = np.sort(np.concatenate([[0, n], rng.integers(0, n, nl - 1)])).astype(np.uint32)
    first, last = 17, nl - 5
    init = rng.random((nl, 3))
    want_c, want_s = init.copy(), np.zeros_like(init)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ref_lib().ref_bounding_boxes_df(P(x), P(y), P(z), P(h), P(layout), C.c_int(first), C.c_int(last), C.c_float(2.0),
                                    P(want_c), P(want_s))
    sc, ss = dev(init.copy()), dev(np.zeros_like(init))
    dx, dy, dz, dh, dl = dev(x), dev(y), dev(z), dev(h), dev(layout)
    st = capi().lib().cs_compute_bounding_boxes_df(
        C.c_void_p(dx.data_ptr()), C.c_void_p(dy.data_ptr()), C.c_void_p(dz.data_ptr()), C.c_void_p(dh.data_ptr()),
        C.c_void_p(dl.data_ptr()), C.c_int(first), C.c_int(last), C.c_float(2.0), C.c_void_p(sc.data_ptr()),
        C.c_void_p(ss.data_ptr()), None)
    assert st == 0
    torch.cuda.synchronize()
    assert np.array_equal(host(sc)[first:last], want_c[first:last])
    assert np.array_equal(host(ss)[first:last], want_s[first:last])


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref")
@pytest.mark.parametrize("pbc", [0, 1])
def test_find_neighbors_double_coordinates_float_h(pbc):
    """findNeighbors with Th = float, Tc = double (findneighbors.hpp:89-99): radiusSq is formed in float"""
    import ctypes as C

    from _libs import ref_lib
    n, ngmax = 30000, 96
    keys, (x, y, z), lim, _ = sorted_keys("u64d", n, "uniform")
    bnd = (pbc, pbc, pbc)
    lo, co = oracle().compute_octree("u64", keys, 16)
    to = oracle().build_octree("u64", lo)
    cen_o, siz_o = oracle().node_fp_centers("u64d", to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    rng = np.random.default_rng(4)
    h = (const_h(n, 40, np.float64, 8.0) * (0.6 + 0.8 * rng.random(n))).astype(np.float32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lim_a, bnd_a = np.array(lim, dtype=np.float64), np.array(bnd, dtype=np.int32)
    nb_o, nc_o = np.zeros(n * ngmax, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
    ref_lib().ref_find_neighbors_u64df(P(x), P(y), P(z), P(h), C.c_uint(0), C.c_uint(n), P(lim_a), P(bnd_a),
                                       C.c_int(to["numLeaves"]), C.c_int(to["numNodes"]), P(to["prefixes"]),
                                       P(to["childOffsets"]), P(to["parents"]), P(to["internalToLeaf"]),
                                       P(to["leafToInternal"]), P(to["levelRange"]), P(lo), P(layout_o), P(cen_o),
                                       P(siz_o), C.c_uint(ngmax), P(nb_o), P(nc_o))
    tree = capi().Octree(dev(lo))
    cen, siz = capi().compute_geo_centers(tree.prefixes, torch.float64, lim, bnd)
    dx, dy, dz, dh, dl = dev(x), dev(y), dev(z), dev(h), dev(layout_o)
    nb = torch.zeros(n * ngmax, dtype=torch.uint32, device="cuda")
    nc = torch.zeros(n, dtype=torch.uint32, device="cuda")
    lim_c, bnd_c = (C.c_double * 6)(*lim), (C.c_int * 3)(*bnd)
    V = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = capi().lib().cs_find_neighbors_df(V(dx), V(dy), V(dz), V(dh), C.c_uint32(0), C.c_uint32(n), lim_c, bnd_c,
                                           C.c_int(tree.num_leaves), V(tree.child_offsets), V(tree.parents),
                                           V(tree.internal_to_leaf), V(dl), V(cen), V(siz), C.c_uint32(ngmax), V(nb),
                                           V(nc), None)
    assert st == 0
    torch.cuda.synchronize()
    assert np.array_equal(host(nc), nc_o)
    m = np.arange(ngmax)[None, :] < np.minimum(nc_o, ngmax)[:, None]
    assert np.array_equal(host(nb).reshape(n, ngmax)[m], nb_o.reshape(n, ngmax)[m])
    assert nc_o.max() > 10


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref")
@pytest.mark.parametrize("frac", [0.002, 0.05])
@pytest.mark.parametrize("pbc,search", [(0, 1), (0, 2), (1, 1), (1, 2)])
def test_find_neighbors_particles_outside_their_leaf_boxes(pbc, search, frac):
    """Arrays that do not belong to the tree: a fraction of the particles is moved by up to a search radius AFTER keys,
    tree and layout were built, so they lie outside the box of their leaf (0.2 %: the group-steered search runs and
    takes its exact route on the marked leaves; 5 %: more than one leaf in 32 is marked, it declines and the per-lane
    search runs instead).  The reference still finds such a particle only
    from targets whose own walk enters its leaf (findneighbors.hpp:108-146), which no longer follows from the
    distance alone.  Both searches (1 = per-lane walks, 2 = group-steered with its stray-leaf preparation, forced) must
    return the reference's lists; the unperturbed case on the same tree is checked too."""
    import ctypes as C

    from _libs import ref_lib
    n, ngmax = 40000, 128
    keys, (x, y, z), lim, _ = sorted_keys("u64d", n, "uniform")
    bnd = (pbc, pbc, pbc)
    lo, co = oracle().compute_octree("u64", keys, 16)
    to = oracle().build_octree("u64", lo)
    cen_o, siz_o = oracle().node_fp_centers("u64d", to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    h = const_h(n, 40, np.float64, 8.0)
    rng = np.random.default_rng(12)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lim_a, bnd_a = np.array(lim, dtype=np.float64), np.array(bnd, dtype=np.int32)
    tree = capi().Octree(dev(lo))
    cen, siz = capi().compute_geo_centers(tree.prefixes, torch.float64, lim, bnd)
    dl = dev(layout_o)
    capi().tuning_set(2, search)
    try:
        for moved in (False, True):
            xs, ys, zs = x.copy(), y.copy(), z.copy()
            if moved:
                pick = rng.random(n) < frac
                for a in (xs, ys, zs):
                    a[pick] += (rng.random(int(pick.sum())) - 0.5) * 4.0 * h[0]
                    np.clip(a, lim[0] + 1e-9, lim[1] - 1e-9, out=a)
            nb_o, nc_o = np.zeros(n * ngmax, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
            ref_lib().ref_find_neighbors_u64d(P(xs), P(ys), P(zs), P(h), C.c_uint(0), C.c_uint(n), P(lim_a), P(bnd_a),
                                              C.c_int(to["numLeaves"]), C.c_int(to["numNodes"]), P(to["prefixes"]),
                                              P(to["childOffsets"]), P(to["parents"]), P(to["internalToLeaf"]),
                                              P(to["leafToInternal"]), P(to["levelRange"]), P(lo), P(layout_o),
                                              P(cen_o), P(siz_o), C.c_uint(ngmax), P(nb_o), P(nc_o))
            nb, nc = capi().find_neighbors(dev(xs), dev(ys), dev(zs), dev(h), 0, n, lim, bnd, tree, dl, cen, siz, ngmax)
            torch.cuda.synchronize()
            assert np.array_equal(host(nc), nc_o), (moved, int((host(nc) != nc_o).sum()))
            m = np.arange(ngmax)[None, :] < np.minimum(nc_o, ngmax)[:, None]
            assert np.array_equal(host(nb).reshape(n, ngmax)[m], nb_o.reshape(n, ngmax)[m]), moved
            assert nc_o.max() > 10
    finally:
        capi().tuning_set(2, 0)


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref")
@pytest.mark.parametrize("combo,factor,pbc,search", [("u64d", 1.3, 0, 1), ("u64d", 1.3, 1, 2), ("u64d", 0.8, 0, 2),
                                                     ("u64d", 0.8, 1, 1), ("u32f", 1.5, 0, 1), ("u32f", 0.7, 1, 1),
                                                     ("u64f", 2.0, 0, 1)])
def test_find_neighbors_search_ext_factor(combo, factor, pbc, search):
    """OctreeNsView::searchExtFactor (tree/octree.hpp:279-282): the continuation tests of the walk use the radius
    2h * factor, acceptance stays at 2h (findneighbors.hpp:99-146).  Particles are moved off their leaves so that the
    factor changes the result (a larger factor reaches strays that factor 1 misses, a smaller one drops neighbours
    whose leaf box is no longer entered); lists and counts equal the reference's for both searches."""
    import ctypes as C

    from _libs import ref_lib
    kt, T = key_of(combo), real_of(combo)
    n, ngmax = 30000, 128
    keys, (x, y, z), lim, _ = sorted_keys(combo, n, "uniform")
    bnd = (pbc, pbc, pbc)
    lo, co = oracle().compute_octree(kt, keys, 16)
    to = oracle().build_octree(kt, lo)
    cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd)
    layout_o = np.zeros(lo.size, dtype=np.uint32)
    layout_o[1:] = np.cumsum(co)
    h = const_h(n, 40, T, 8.0)
    rng = np.random.default_rng(21)
    pick = rng.random(n) < 0.01
    for a in (x, y, z):
        a[pick] += ((rng.random(int(pick.sum())) - 0.5) * 2.0 * h[0]).astype(T)
        np.clip(a, T(lim[0] + 1e-3), T(lim[1] - 1e-3), out=a)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lim_a, bnd_a = np.array(lim, dtype=np.float64), np.array(bnd, dtype=np.int32)

    def reference(f):
        nb_o, nc_o = np.zeros(n * ngmax, dtype=np.uint32), np.zeros(n, dtype=np.uint32)
        ref_lib().ref_set_search_ext_factor(C.c_float(f))
        try:
            getattr(ref_lib(), "ref_find_neighbors_" + combo)(
                P(x), P(y), P(z), P(h), C.c_uint(0), C.c_uint(n), P(lim_a), P(bnd_a), C.c_int(to["numLeaves"]),
                C.c_int(to["numNodes"]), P(to["prefixes"]), P(to["childOffsets"]), P(to["parents"]),
                P(to["internalToLeaf"]), P(to["leafToInternal"]), P(to["levelRange"]), P(lo), P(layout_o), P(cen_o),
                P(siz_o), C.c_uint(ngmax), P(nb_o), P(nc_o))
        finally:
            ref_lib().ref_set_search_ext_factor(C.c_float(1.0))
        return nb_o, nc_o

    nb_o, nc_o = reference(factor)
    _, nc_1 = reference(1.0)
    assert not np.array_equal(nc_o, nc_1), "the factor must change the result for the test to mean anything"
    tree = capi().Octree(dev(lo))
    cen, siz = capi().compute_geo_centers(tree.prefixes, torch.float32 if T == np.float32 else torch.float64, lim, bnd)
    capi().tuning_set(2, search)
    try:
        nb, nc = capi().find_neighbors(dev(x), dev(y), dev(z), dev(h), 0, n, lim, bnd, tree, dev(layout_o), cen, siz,
                                       ngmax, search_ext_factor=factor)
        torch.cuda.synchronize()
    finally:
        capi().tuning_set(2, 0)
    assert np.array_equal(host(nc), nc_o), int((host(nc) != nc_o).sum())
    m = np.arange(ngmax)[None, :] < np.minimum(nc_o, ngmax)[:, None]
    assert np.array_equal(host(nb).reshape(n, ngmax)[m], nb_o.reshape(n, ngmax)[m])
