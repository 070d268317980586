"""Pins the C oracle (oracle/liboracle.so) against the known-answer vectors held by the reference's own unit tests.
Citations are relative to /root/reference/test/unit."""
import ctypes as C

import numpy as np
import pytest

from _libs import KEYS, MAXLEVEL, oracle, orc_decode, orc_scalar
from _util import OctreeMaker, node_range

KT = ["u32", "u64"]


def pad(kt, prefix, length):
    return prefix << (3 * MAXLEVEL[kt] - length)


@pytest.mark.parametrize("kt", KT)
def test_imorton3d(kt):  # sfc/morton.cpp:26-34
    shift = MAXLEVEL[kt] - 3
    assert orc_scalar("imorton", kt, 5 << shift, 3 << shift, 6 << shift) == pad(kt, 0b101011110, 9)


def test_decode_morton():  # sfc/morton.cpp:42-73
    assert orc_decode("decode_morton", "u32", 340) == (5, 2, 4)
    m = (1 << 21) - 1
    assert orc_decode("decode_morton", "u64", 0x7FFFFFFFFFFFFFFF) == (m, m, m)
    assert orc_decode("decode_morton", "u64", 0x1249249241249249)[2] == (1 << 21) - 512 - 1
    assert orc_decode("decode_morton", "u64", 0b0111 << 60) == (1 << 20, 1 << 20, 1 << 20)
    assert orc_decode("decode_morton", "u64", 0b0011 << 60) == (0, 1 << 20, 1 << 20)


@pytest.mark.parametrize("kt", KT)
def test_hilbert_first_order_curve(kt):  # sfc/hilbert.cpp:84-112
    hilbert_to_morton = [0, 1, 3, 2, 6, 7, 5, 4]
    L1 = (1 << MAXLEVEL[kt]) // 2
    for xi in range(2):
        for yi in range(2):
            for zi in range(2):
                for off in (0, L1 - 1):
                    key = orc_scalar("ihilbert", kt, L1 * xi + off, L1 * yi + off, L1 * zi + off)
                    octant = (key >> (3 * (MAXLEVEL[kt] - 1))) & 7
                    assert hilbert_to_morton[octant] == 4 * xi + 2 * yi + zi


@pytest.mark.parametrize("kt", KT)
def test_hilbert_continuity(kt):  # sfc/hilbert.cpp:121-145
    for level in range(1, MAXLEVEL[kt]):
        for octant in range(8 if level > 1 else 7):
            last = (octant + 1) * node_range(kt, level) - 1
            a = orc_decode("decode_hilbert", kt, last)
            b = orc_decode("decode_hilbert", kt, last + 1)
            assert sum(abs(int(p) - int(q)) for p, q in zip(a, b)) == 1


@pytest.mark.parametrize("kt", KT)
def test_hilbert_inversion(kt):  # sfc/hilbert.cpp:154-182
    rng = np.random.default_rng(0)
    m = 1 << MAXLEVEL[kt]
    for x, y, z in rng.integers(0, m, size=(1000, 3)):
        key = orc_scalar("ihilbert", kt, int(x), int(y), int(z))
        assert orc_decode("decode_hilbert", kt, key) == (x, y, z)
    for x, y, z in [(0, 0, 0), (m - 1, m - 1, m - 1), (m - 1, 0, 0), (0, m - 1, 0), (0, 0, m - 1)]:
        key = orc_scalar("ihilbert", kt, x, y, z)
        assert orc_decode("decode_hilbert", kt, key) == (x, y, z)


@pytest.mark.parametrize("kt", KT)
def test_count_tree_nodes(kt):  # tree/csarray.cpp:57-76
    tree = OctreeMaker(kt).divide().divide(0).make_tree()
    t = [int(v) for v in tree]
    keys = np.array([t[1], t[1], t[1] + 10, t[1] + 100, t[2] - 1, t[2] + 1, t[11], t[11] + 2, t[12], t[12] + 1000,
                     t[12] + 2000, t[13] - 10, t[13], t[13] + 1], dtype=KEYS[kt])
    counts = oracle().compute_node_counts(kt, tree, keys)
    assert counts.tolist() == [0, 5, 1, 0, 0, 0, 0, 0, 0, 0, 0, 2, 4, 2, 0]


@pytest.mark.parametrize("kt", KT)
def test_count_first_last_node(kt):  # tree/csarray.cpp:80-98 (spanning tree {0,1,2^(3L)-1,2^(3L)} restated directly)
    tree = OctreeMaker(kt)
    path = []
    for _ in range(MAXLEVEL[kt]):
        tree.divide(*path)
        path.append(0)
    path = []
    for lvl in range(MAXLEVEL[kt]):
        if lvl:
            tree.divide(*path)
        path.append(7)
    leaves = tree.make_tree()
    assert leaves[1] == 1 and leaves[-2] == node_range(kt, 0) - 1
    keys = np.array([0, 0, node_range(kt, 0) - 1, node_range(kt, 0) - 1], dtype=KEYS[kt])
    counts = oracle().compute_node_counts(kt, leaves, keys)
    ref = np.zeros(leaves.size - 1, dtype=np.uint32)
    ref[0] = ref[-1] = 2
    assert np.array_equal(counts, ref)


@pytest.mark.parametrize("kt", KT)
def test_rebalance_decision(kt):  # tree/csarray.cpp:108-122
    tree = OctreeMaker(kt).divide().divide(0).make_tree()
    counts = np.array([1, 1, 1, 0, 0, 0, 0, 0, 2, 3, 4, 5, 6, 7, 8], dtype=np.uint32)
    ops, converged = oracle().rebalance_decision(kt, tree, counts, 4)
    assert ops.tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 8, 8, 8, 8]
    assert not converged


@pytest.mark.parametrize("kt", KT)
def test_rebalance_tree(kt):  # tree/csarray.cpp:193-206
    lib = oracle().lib
    tree = OctreeMaker(kt).divide().divide(0).make_tree()
    ops = np.array([1, 0, 0, 0, 0, 0, 0, 0, 1, 8, 1, 1, 1, 1, 8, 0], dtype=np.int32)
    new = np.zeros(64, dtype=KEYS[kt])
    f = getattr(lib, "orc_rebalance_tree_" + kt)
    f.restype = C.c_int
    n = f(tree.ctypes.data_as(C.c_void_p), C.c_int(15), ops.ctypes.data_as(C.c_void_p), new.ctypes.data_as(C.c_void_p))
    ref = OctreeMaker(kt).divide().divide(2).divide(7).make_tree()
    assert n == ref.size - 1
    assert np.array_equal(new[:n + 1], ref)


@pytest.mark.parametrize("kt", KT)
def test_octree_connectivity_and_upsweep(kt):  # tree/octree.cpp:27-74 (checkConnectivity) and :349-371 (upsweep)
    leaves = OctreeMaker(kt).divide().divide(0).divide(0, 2).divide(3).make_tree()
    o = oracle()
    t = o.build_octree(kt, leaves)
    nn, ni = t["numNodes"], t["numInternal"]
    assert ni == (leaves.size - 2) // 7
    co, par, pf = t["childOffsets"], t["parents"], t["prefixes"]
    assert np.all(np.diff(pf.astype(np.uint64)) > 0)  # level-major, key-minor order
    for i in range(nn):
        if co[i]:
            plen = int(pf[i]).bit_length() - 1
            for s in range(8):
                c = co[i] + s
                assert par[(c - 1) // 8] == i
                assert int(pf[c]).bit_length() - 1 == plen + 3
                assert int(pf[c]) >> 3 == int(pf[i]) and int(pf[c]) & 7 == s
        else:
            leaf = t["internalToLeaf"][i]
            assert 0 <= leaf < t["numLeaves"]
            assert t["leafToInternal"][ni + leaf] == i
    counts = np.zeros(nn, dtype=np.uint32)
    counts[t["leafToInternal"][ni:]] = 1
    o._fn("upsweep_counts_" + kt)(t["levelRange"].ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p),
                                  counts.ctypes.data_as(C.c_void_p))
    assert counts.tolist() == [29, 15, 1, 1, 8, 1, 1, 1, 1, 1, 1, 8] + [1] * 21


@pytest.mark.parametrize("kt", KT)
def test_compute_octree_invariants(kt):  # tree/csarray.cpp:303-348 (random Gaussian keys, invariants + counts)
    rng = np.random.default_rng(3)
    nr = node_range(kt, 0)
    g = rng.normal(nr / 2, nr / 5, 120000)
    keys = g[(g >= 0) & (g < nr - 1)][:100000].astype(KEYS[kt])
    keys.sort()
    for bucket in (64, 1024, 10000):
        leaves, counts = oracle().compute_octree(kt, keys, bucket)
        assert leaves[0] == 0 and int(leaves[-1]) == nr
        d = np.diff(leaves.astype(np.uint64)).astype(np.uint64)
        assert np.all(d > 0)
        for v in np.unique(d):  # every node range is a power of 8
            v = int(v)
            assert v & (v - 1) == 0 and (v.bit_length() - 1) % 3 == 0
        assert counts.sum() == keys.size
        assert counts.max() <= bucket
        ref = np.searchsorted(keys, leaves[1:], side="left") - np.searchsorted(keys, leaves[:-1], side="left")
        assert np.array_equal(counts, ref.astype(np.uint32))


# ---------------------------------------------------------------------------------------------- LET rebalance decisions
from _focus_vectors import ENFORCE_CASES, ESSENTIAL_CASES, decode_placeholder, octree_maker  # noqa: E402


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _node_arrays(orc, kt, cstree, leaf_counts, leaf_macs, internal_macs):
    tree = orc.build_octree(kt, cstree)
    nn, ni = tree["numNodes"], tree["numInternal"]
    l2i = tree["leafToInternal"][ni:]
    counts = np.zeros(nn, dtype=np.uint32)
    counts[l2i] = leaf_counts
    orc._fn("upsweep_counts_" + kt)(_p(tree["levelRange"]), _p(tree["childOffsets"]), _p(counts))
    macs = np.zeros(nn, dtype=np.uint8)
    macs[l2i] = leaf_macs
    for key, value in internal_macs:
        (idx,) = np.nonzero(tree["prefixes"] == key)[0]
        macs[idx] = value
    return tree, counts, macs, l2i


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("case", range(len(ESSENTIAL_CASES)))
def test_rebalance_decision_essential_vectors(kt, case):
    """test/unit/focus/octree_focus.cpp:70-186: rebalanceDecisionEssential + protectAncestors, bucketSize 1"""
    orc = oracle()
    paths, leaf_counts, leaf_macs, imacs, focus, want, want_conv = ESSENTIAL_CASES[case]
    cstree = octree_maker(kt, *paths)
    tree, counts, macs, l2i = _node_arrays(orc, kt, cstree, leaf_counts, leaf_macs, imacs)
    nn = tree["numNodes"]
    parents = np.ascontiguousarray(tree["parents"])
    ops = np.zeros(nn, dtype=np.int32)
    cast = C.c_uint64 if kt == "u64" else C.c_uint32
    orc._fn("rebalance_decision_essential_" + kt)(_p(tree["prefixes"]), C.c_int(nn), _p(tree["childOffsets"]),
                                                  _p(parents), _p(counts), _p(macs), cast(int(cstree[focus[0]])),
                                                  cast(int(cstree[focus[1]])), C.c_uint(1), _p(ops))
    conv = orc._fn("protect_ancestors_" + kt, C.c_int)(_p(tree["prefixes"]), C.c_int(nn), _p(parents), _p(ops))
    assert ops[l2i].tolist() == want
    assert bool(conv) == want_conv


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("case", range(len(ENFORCE_CASES)))
def test_key_enforcement_vectors(kt, case):
    """test/unit/focus/octree_focus.cpp:228-290: enforceKeySingle on the 17-node tree divide().divide(1)"""
    orc = oracle()
    start, codes, statuses, protect, want = ENFORCE_CASES[case]
    np_t, max_level = (np.uint32, 10) if kt == "u32" else (np.uint64, 21)
    tree = orc.build_octree(kt, octree_maker(kt, (), (1,)))
    assert tree["numNodes"] == 17
    parents = np.ascontiguousarray(tree["parents"])
    ops = np.array(start, dtype=np.int32)
    for code, status in zip(codes, statuses):
        key = np.array([decode_placeholder(code, max_level)], dtype=np_t)
        got = orc._fn("enforce_keys_" + kt, C.c_int)(_p(key), C.c_int(1), _p(tree["prefixes"]),
                                                     _p(tree["childOffsets"]), _p(parents), _p(ops))
        assert got == status
    if protect:
        orc._fn("protect_ancestors_" + kt, C.c_int)(_p(tree["prefixes"]), C.c_int(17), _p(parents), _p(ops))
    assert ops.tolist() == want


# ---------------------------------------------------------------------------------------------- halo discovery
def uniform_level_tree(kt, level):
    np_t, max_level = (np.uint32, 10) if kt == "u32" else (np.uint64, 21)
    n = 8 ** level
    step = 1 << (3 * (max_level - level))
    return (np.arange(n + 1, dtype=np.uint64) * step).astype(np_t)


def halo_flags_setup(backend, combo):
    kt = "u32" if combo.startswith("u32") else "u64"
    T = np.float32 if combo.endswith("f") else np.float64
    lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
    leaves = uniform_level_tree(kt, 2)                       # makeUniformNLevelTree(64, 1): 4 x 4 x 4 leaves
    tree = backend.build_octree(kt, leaves)
    cen, siz = backend.node_fp_centers(combo, tree["prefixes"], lim, bnd)
    l2i = tree["leafToInternal"][tree["numInternal"]:]
    cen3, siz3 = cen.reshape(-1, 3), siz.reshape(-1, 3)
    sc = np.ascontiguousarray(cen3[l2i]).astype(T)
    ss = np.ascontiguousarray(siz3[l2i] + T(0.1)).astype(T)  # size of one node is 0.25^3, search radius 0.1
    return kt, leaves, tree, cen, siz, sc, ss, lim, bnd


def all_to_all_flags(tree, leaves, cen, siz, sc, ss, first, last):
    """findHalosAll2All of test/unit/traversal/discovery.cpp:25-49 for an open box"""
    cen3, siz3 = cen.reshape(-1, 3), siz.reshape(-1, 3)
    flags = np.zeros(tree["numNodes"], dtype=np.uint8)
    max_level = 10 if leaves.dtype == np.uint32 else 21
    lo, hi = int(leaves[first]), int(leaves[last])
    for n in range(tree["numNodes"]):
        p = int(tree["prefixes"][n])
        length = p.bit_length() - 1
        k1 = (p ^ (1 << length)) << (3 * max_level - length)
        k2 = k1 + (1 << (3 * max_level - length))
        if lo <= k1 and k2 <= hi:
            continue                                          # sources inside the excluded range do not count
        for t in range(first, last):
            if all(abs(cen3[n][d] - sc[t][d]) - siz3[n][d] - ss[t][d] < 0 for d in range(3)):
                flags[n] = 1
                break
    return flags


@pytest.mark.parametrize("combo", ["u32f", "u64d"])
def test_find_halos_flags_21(combo):
    """test/unit/traversal/discovery.cpp:52-103: the surface between the first and the last 32 leaves of a 4x4x4 tree is
    16 leaves + 5 internal nodes, from either side, and equals the all-to-all collision search"""
    orc = oracle()
    kt, leaves, tree, cen, siz, sc, ss, lim, bnd = halo_flags_setup(orc, combo)
    for first, last in ((0, 32), (32, 64)):
        flags = orc.find_halos(combo, tree, cen, siz, leaves, sc, ss, lim, bnd, first, last)
        assert int(flags.sum()) == 21
        assert np.array_equal(flags, all_to_all_flags(tree, leaves, cen, siz, sc, ss, first, last))


# ---------------------------------------------------------------------------------------------- sort + gather
@pytest.mark.parametrize("kt", ["u32", "u64"])
def test_sort_by_key_and_gather_vectors(kt):
    """test/unit/primitives/gather.cpp:24-67: the ordering that sorts {2,1,5,4} is {1,0,3,2}; sorting
    {0,50,10,60,...} and gathering a value array through the ordering gives the listed reference"""
    orc = oracle()
    K = KEYS[kt]
    keys = np.array([2, 1, 5, 4], dtype=K)
    order = np.arange(4, dtype=np.uint32)
    orc.sort_by_key(kt, keys, order)
    assert order.tolist() == [1, 0, 3, 2] and keys.tolist() == [1, 2, 4, 5]

    keys = np.array([0, 50, 10, 60, 20, 70, 30, 80, 40, 90], dtype=K)
    order = np.arange(10, dtype=np.uint32)
    orc.sort_by_key(kt, keys, order)
    assert keys.tolist() == [0, 10, 20, 30, 40, 50, 60, 70, 80, 90]
    values = np.array([-2, -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11], dtype=np.float64)
    probe = values.copy()
    probe[2:12] = values[2:][order]                           # gatherCpu(ordering, values + 2, probe + 2)
    assert probe.tolist() == [-2, -1, 0, 2, 4, 6, 8, 1, 3, 5, 7, 9, 10, 11]


def test_groups_oracle_reference_known_answers():
    """oracle/groups_oracle.py pinned to test/unit_cuda/traversal/groups.cu: fixed groups :27-41, the two split counts
    {2, 2} and {64, 60} of groupSplitsKernel :252-271, and the boundaries (4, 6, 68, 75, 128) of computeGroupSplits
    :273-281"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import groups_oracle

    assert groups_oracle.fixed_groups(4, 34, 8).tolist() == [4, 12, 20, 28, 34]
    r = np.uint64(1) << np.uint64(63)
    o = r >> np.uint64(3)
    leaves = [np.uint64(0), o, np.uint64(2) * o]
    leaves += [np.uint64(2) * o + np.uint64(k) * (o >> np.uint64(3)) for k in range(1, 8)]
    leaves += [np.uint64(k) * o for k in range(3, 8)] + [r]
    leaves = np.array(leaves, dtype=np.uint64)
    first, last, G = 4, 128, 64
    counts = np.array([4, 1, 8, 8, 8, 8, 31, 8, 8, 8, 16, 16, 16, 0, 0], dtype=np.uint32)
    layout = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    x = np.arange(last, dtype=np.float64)
    y, z = x.copy(), x.copy()
    h = np.full(last, float(last))
    h[first + G + 6] = 0.99 * np.sqrt(3.0) / 2
    h[first + G + 7] = 1.01 * np.sqrt(3.0) / 2
    for a in (x, y, z):
        a[5] -= 0.01
    lim = (0, last, 0, last, 0, last)
    dist_crit = np.cbrt(float(last) ** 3 / 64)
    loose = groups_oracle.group_splits(first, last, x, y, z, h, leaves, layout, lim, G,
                                       np.float32(np.sqrt(3.0) / dist_crit * 1.01))
    assert loose.tolist() == [4, 6, 68, 75, 128]
    tight = groups_oracle.group_splits(first, last, x, y, z, h, leaves, layout, lim, G,
                                       np.float32(np.sqrt(3.0) / dist_crit * 0.99))
    assert int((tight < 68).sum()) == 64 and int((tight >= 68).sum()) - 1 == 60
