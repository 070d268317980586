"""Pins the C oracle against the UNMODIFIED reference (oracle/_ref/libcstone_ref.so) on seeded random inputs.
Skipped when the reference library was never built (it is built wherever /root/reference exists and then travels
with the repo snapshot)."""
import numpy as np
import pytest

from _libs import COMBOS, KEYS, key_of, oracle, real_of, ref
from _util import const_h, gaussian_particles, uniform_particles

pytestmark = pytest.mark.skipif(ref() is None, reason="oracle/_ref/libcstone_ref.so not built")

BOXES = [
    ((0, 1, 0, 1, 0, 1), (0, 0, 0)),
    ((0, 1, 0, 1, 0, 1), (1, 1, 1)),
    ((-1.2, 1.3, -0.4, 2.2, -3.0, 1.0), (1, 0, 1)),
]


def particles(combo, n, lim, seed=7):
    T = real_of(combo)
    rng = np.random.default_rng(seed)
    out = []
    for d in range(3):
        lo, hi = lim[2 * d], lim[2 * d + 1]
        a = (lo + (hi - lo) * rng.random(n)).astype(T)
        np.clip(a, T(lo), np.nextafter(T(hi), T(lo)), out=a)
        out.append(a)
    return out


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("lim,bnd", BOXES)
def test_sfc_keys(combo, kind, lim, bnd):
    x, y, z = particles(combo, 20000, lim)
    ko = oracle().sfc_keys(combo, kind, x, y, z, lim, bnd)
    kr = ref().sfc_keys(combo, kind, x, y, z, lim, bnd)
    assert np.array_equal(ko, kr)
    # removeKey entries are left untouched (sfc/sfc.hpp:274)
    K = KEYS[key_of(combo)]
    remove = K(1) << K(30 if key_of(combo) == "u32" else 63)
    pre = np.zeros(x.size, dtype=K)
    pre[::7] = remove
    ko2 = oracle().sfc_keys(combo, kind, x, y, z, lim, bnd, keys=pre.copy())
    kr2 = ref().sfc_keys(combo, kind, x, y, z, lim, bnd, keys=pre.copy())
    assert np.array_equal(ko2, kr2)
    assert np.all(ko2[::7] == remove)


@pytest.mark.parametrize("kt", ["u32", "u64"])
def test_sort_by_key(kt):
    rng = np.random.default_rng(5)
    K = KEYS[kt]
    for n, hi in [(0, 10), (1, 10), (1000, 50), (50000, 2 ** 30)]:  # small hi => many duplicates => stability matters
        keys = rng.integers(0, hi, n).astype(K)
        vo = np.arange(n, dtype=np.uint32)
        vr = vo.copy()
        ko, kr = keys.copy(), keys.copy()
        oracle().sort_by_key(kt, ko, vo)
        ref().sort_by_key(kt, kr, vr)
        assert np.array_equal(ko, kr) and np.array_equal(vo, vr)
        assert np.array_equal(vo, np.argsort(keys, kind="stable").astype(np.uint32))


def _sorted_keys(combo, n, gaussian, seed=11):
    T = real_of(combo)
    lim, bnd = (-1, 1, -1, 1, -1, 1), (0, 0, 0)
    x, y, z = gaussian_particles(n, T, seed) if gaussian else uniform_particles(n, T, seed, -1, 1)
    keys = oracle().sfc_keys(combo, 0, x, y, z, lim, bnd)
    order = np.arange(n, dtype=np.uint32)
    oracle().sort_by_key(key_of(combo), keys, order)
    return keys, order, (x[order], y[order], z[order]), lim, bnd


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("gaussian", [False, True])
@pytest.mark.parametrize("bucket", [1, 16, 64])
def test_tree_build_and_link(combo, gaussian, bucket):
    kt = key_of(combo)
    keys, _, _, lim, bnd = _sorted_keys(combo, 30000, gaussian)
    lo, co = oracle().compute_octree(kt, keys, bucket)
    lr, cr = ref().compute_octree(kt, keys, bucket)
    assert np.array_equal(lo, lr) and np.array_equal(co, cr)

    # single update steps from a coarse tree, including the decision vector
    leaves = np.array([0, int(lo[-1])], dtype=KEYS[kt])
    counts = np.array([keys.size], dtype=np.uint32)
    for _ in range(3):
        oo, convo = oracle().rebalance_decision(kt, leaves, counts, bucket)
        orr, convr = ref().rebalance_decision(kt, leaves, counts, bucket)
        assert np.array_equal(oo, orr) and convo == convr
        l1, c1, v1 = oracle().update_octree(kt, keys, bucket, leaves, counts)
        l2, c2, v2 = ref().update_octree(kt, keys, bucket, leaves, counts)
        assert np.array_equal(l1, l2) and np.array_equal(c1, c2) and v1 == v2
        leaves, counts = l1, c1

    assert np.array_equal(oracle().compute_node_counts(kt, lo, keys[::3].copy(), 5),
                          ref().compute_node_counts(kt, lo, keys[::3].copy(), 5))

    to, tr = oracle().build_octree(kt, lo), ref().build_octree(kt, lo)
    for k in ("prefixes", "childOffsets", "parents", "levelRange", "internalToLeaf", "leafToInternal"):
        assert np.array_equal(to[k], tr[k]), k

    cen_o, siz_o = oracle().node_fp_centers(combo, to["prefixes"], lim, bnd)
    cen_r, siz_r = ref().node_fp_centers(combo, to["prefixes"], lim, bnd)
    assert np.array_equal(cen_o, cen_r) and np.array_equal(siz_o, siz_r)


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("pbc", [0, 1])
@pytest.mark.parametrize("gaussian", [False, True])
def test_find_neighbors(combo, pbc, gaussian):
    kt, T = key_of(combo), real_of(combo)
    n = 6000
    keys, _, (x, y, z), lim, _ = _sorted_keys(combo, n, gaussian)
    bnd = (pbc, pbc, pbc)
    leaves, counts = oracle().compute_octree(kt, keys, 16)
    tree = oracle().build_octree(kt, leaves)
    centers, sizes = oracle().node_fp_centers(combo, tree["prefixes"], lim, bnd)
    layout = np.zeros(leaves.size, dtype=np.uint32)
    layout[1:] = np.cumsum(counts)
    rng = np.random.default_rng(2)
    h = (const_h(n, 40, T, 8.0) * (0.6 + 0.8 * rng.random(n))).astype(T)
    for ngmax in (16, 150):
        nbo, nco = oracle().find_neighbors(combo, x, y, z, h, 100, n - 50, lim, bnd, tree, leaves, layout, centers,
                                           sizes, ngmax)
        nbr, ncr = ref().find_neighbors(combo, x, y, z, h, 100, n - 50, lim, bnd, tree, leaves, layout, centers,
                                        sizes, ngmax)
        assert np.array_equal(nco, ncr)
        m = np.arange(ngmax)[None, :] < np.minimum(nco, ngmax)[:, None]
        assert np.array_equal(nbo[m], nbr[m])
        assert nco.max() > ngmax if ngmax == 16 else True


@pytest.mark.parametrize("combo", COMBOS)
@pytest.mark.parametrize("pbc", [0, 1])
def test_halo_discovery(combo, pbc):
    kt, T = key_of(combo), real_of(combo)
    n = 20000
    keys, _, (x, y, z), lim, _ = _sorted_keys(combo, n, True)
    bnd = (pbc, pbc, pbc)
    leaves, counts = oracle().compute_octree(kt, keys, 8)
    tree = oracle().build_octree(kt, leaves)
    centers, sizes = oracle().node_fp_centers(combo, tree["prefixes"], lim, bnd)
    layout = np.zeros(leaves.size, dtype=np.uint32)
    layout[1:] = np.cumsum(counts)
    h = const_h(n, 30, T, 8.0)
    nl = leaves.size - 1
    for first, last in [(0, nl // 4), (nl // 3, 2 * nl // 3), (nl - 5, nl), (0, nl)]:
        init = centers[tree["leafToInternal"][tree["numInternal"]:]]
        sco, sso = oracle().bounding_boxes(combo, x, y, z, h, layout, first, last, 2.0, init)
        scr, ssr = ref().bounding_boxes(combo, x, y, z, h, layout, first, last, 2.0, init)
        assert np.array_equal(sco[first:last], scr[first:last]) and np.array_equal(sso[first:last], ssr[first:last])
        fo = oracle().find_halos(combo, tree, centers, sizes, leaves, sco, sso, lim, bnd, first, last)
        fr = ref().find_halos(combo, tree, centers, sizes, leaves, scr, ssr, lim, bnd, first, last)
        assert np.array_equal(fo, fr)
        if (first, last) != (0, nl):
            assert fo.sum() > 0
