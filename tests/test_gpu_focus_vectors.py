"""The LET rebalance kernels against the reference's explicit known-answer vectors
(test/unit/focus/octree_focus.cpp:26-186 rebalanceDecisionEssential + protectAncestors, :228-290 enforceKeys), through
the C ABI entry points that stand in for rebalanceDecisionEssentialGpu / protectAncestorsGpu / enforceKeysGpu
(focus/rebalance_gpu.h:27-79)."""
import ctypes as C

import numpy as np
import pytest
import torch

from _focus_vectors import ENFORCE_CASES, ESSENTIAL_CASES, decode_placeholder, octree_maker

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = {"u32": (np.uint32, torch.uint32, 10), "u64": (np.uint64, torch.uint64, 21)}


def capi():
    from cstone_b200 import capi as c
    return c


def linked(kt, cstree):
    c = capi()
    _, torch_t, _ = KEYS[kt]
    leaves = torch.from_numpy(cstree.view(np.int64 if kt == "u64" else np.int32)).to(DEV).view(torch_t)
    return c.Octree(leaves)


def node_ops_essential(kt, tree, cstree, leaf_counts, leaf_macs, focus, bucket, internal_macs):
    c = capi()
    np_t, _, _ = KEYS[kt]
    nn, ni = tree.num_nodes, tree.num_internal
    l2i = tree.leaf_to_internal[ni:].long()
    counts = torch.zeros(nn, dtype=torch.int32, device=DEV)
    counts[l2i] = torch.tensor(leaf_counts, dtype=torch.int32, device=DEV)
    counts = counts.view(torch.uint32)
    c.upsweep_sum(kt, tree.level_range.cpu().numpy(), tree.child_offsets, counts)
    macs = torch.zeros(nn, dtype=torch.uint8, device=DEV)
    macs[l2i] = torch.tensor(leaf_macs, dtype=torch.uint8, device=DEV)
    prefixes = tree.prefixes.cpu().numpy()
    for key, value in internal_macs:
        (idx,) = np.nonzero(prefixes == np_t(key))[0]
        macs[idx] = value
    ops = torch.zeros(nn, dtype=torch.int32, device=DEV)
    lib = c.lib()
    cast = C.c_uint64 if kt == "u64" else C.c_uint32
    c._check(getattr(lib, "cs_rebalance_decision_essential_" + kt)(
        c._ptr(tree.prefixes), c._ptr(tree.child_offsets), c._ptr(tree.parents), c._ptr(counts), c._ptr(macs),
        cast(int(cstree[focus[0]])), cast(int(cstree[focus[1]])), C.c_uint32(bucket), c._ptr(ops), C.c_int(nn),
        c._stream()), "essential")
    conv = C.c_int(-1)
    c._check(getattr(lib, "cs_protect_ancestors_" + kt)(c._ptr(tree.prefixes), c._ptr(tree.parents), c._ptr(ops),
                                                        C.c_int(nn), C.byref(conv), c._stream()), "protect")
    return ops[l2i].cpu().tolist(), bool(conv.value)


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("case", range(len(ESSENTIAL_CASES)))
def test_rebalance_decision_essential_vectors(kt, case):
    """test/unit/focus/octree_focus.cpp:70-186, bucketSize 1"""
    paths, counts, macs, imacs, focus, want, want_conv = ESSENTIAL_CASES[case]
    cstree = octree_maker(kt, *paths)
    tree = linked(kt, cstree)
    got, conv = node_ops_essential(kt, tree, cstree, counts, macs, focus, 1, imacs)
    assert got == want
    assert conv == want_conv


def enforce(kt, tree, ops, code):
    c = capi()
    np_t, torch_t, max_level = KEYS[kt]
    key = torch.from_numpy(np.array([decode_placeholder(code, max_level)], dtype=np_t).view(
        np.int64 if kt == "u64" else np.int32)).to(DEV).view(torch_t)
    status = C.c_int(-1)
    c._check(getattr(c.lib(), "cs_enforce_keys_" + kt)(c._ptr(key), C.c_int(1), c._ptr(tree.prefixes),
                                                       c._ptr(tree.child_offsets), c._ptr(tree.parents), c._ptr(ops),
                                                       C.byref(status), c._stream()), "enforce")
    return status.value


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("case", range(len(ENFORCE_CASES)))
def test_key_enforcement_vectors(kt, case):
    """test/unit/focus/octree_focus.cpp:228-290: node ops of the 17-node tree divide().divide(1)"""
    c = capi()
    start, codes, statuses, protect, want = ENFORCE_CASES[case]
    tree = linked(kt, octree_maker(kt, (), (1,)))
    assert tree.num_nodes == 17
    ops = torch.tensor(start, dtype=torch.int32, device=DEV)
    for code, status in zip(codes, statuses):
        assert enforce(kt, tree, ops, code) == status
    if protect:
        conv = C.c_int(-1)
        c._check(getattr(c.lib(), "cs_protect_ancestors_" + kt)(c._ptr(tree.prefixes), c._ptr(tree.parents),
                                                                c._ptr(ops), C.c_int(17), C.byref(conv), c._stream()),
                 "protect")
    assert ops.cpu().tolist() == want


@pytest.mark.parametrize("kt", ["u32", "u64"])
@pytest.mark.parametrize("seed", [1, 2])
def test_rebalance_kernels_equal_oracle_on_random_trees(kt, seed):
    """rebalanceDecisionEssential, enforceKeys and protectAncestors on a real tree (Gaussian keys, bucket 16) with
    random MAC flags and a focus in the middle of the curve: node ops, convergence flag and resolution status equal
    the C restatement of focus/rebalance.hpp in oracle/"""
    from _libs import oracle

    c = capi()
    orc = oracle()
    np_t, torch_t, max_level = KEYS[kt]
    rng = np.random.default_rng(seed)
    bits = 3 * max_level
    keys = np.sort(np.clip(rng.normal(0.5, 0.12, 60000), 0, 0.999999) * float(1 << bits)).astype(np_t)
    leaves, leaf_counts = orc.compute_octree(kt, keys, 16)
    ot = orc.build_octree(kt, leaves)
    nn, ni, nl = ot["numNodes"], ot["numInternal"], ot["numLeaves"]
    l2i = ot["leafToInternal"][ni:]
    counts = np.zeros(nn, dtype=np.uint32)
    counts[l2i] = leaf_counts
    orc._fn("upsweep_counts_" + kt)(ot["levelRange"].ctypes.data_as(C.c_void_p),
                                    ot["childOffsets"].ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p))
    macs = (rng.random(nn) < 0.6).astype(np.uint8)
    focus = (int(leaves[nl // 3]), int(leaves[2 * nl // 3]))
    bucket = 8  # smaller than the bucket of the tree: splits, merges and stays all occur
    parents = np.ascontiguousarray(ot["parents"])
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    cast = C.c_uint64 if kt == "u64" else C.c_uint32

    want = np.zeros(nn, dtype=np.int32)
    orc._fn("rebalance_decision_essential_" + kt)(p_(ot["prefixes"]), C.c_int(nn), p_(ot["childOffsets"]), p_(parents),
                                                  p_(counts), p_(macs), cast(focus[0]), cast(focus[1]), C.c_uint(bucket),
                                                  p_(want))
    # mandatory keys: leaf boundaries inside the focus (present) and keys one and three levels below a leaf
    lv = leaves[nl // 3:nl // 3 + 40].astype(np_t)
    fine1 = (leaves[nl // 2:nl // 2 + 5] + (leaves[nl // 2 + 1:nl // 2 + 6] - leaves[nl // 2:nl // 2 + 5]) // 8).astype(np_t)
    fine3 = (leaves[nl // 2 + 9:nl // 2 + 12] + 1).astype(np_t)
    mandatory = np.concatenate([lv, fine1, fine3]).astype(np_t)
    want_status = orc._fn("enforce_keys_" + kt, C.c_int)(p_(mandatory), C.c_int(mandatory.size), p_(ot["prefixes"]),
                                                         p_(ot["childOffsets"]), p_(parents), p_(want))
    want_conv = orc._fn("protect_ancestors_" + kt, C.c_int)(p_(ot["prefixes"]), C.c_int(nn), p_(parents), p_(want))

    view = np.int32 if kt == "u32" else np.int64
    tree = linked(kt, leaves)
    assert np.array_equal(tree.prefixes.cpu().numpy().view(np_t) if False else
                          tree.prefixes.cpu().view(torch.int32 if kt == "u32" else torch.int64).numpy().view(np_t),
                          ot["prefixes"])
    d_counts = torch.from_numpy(counts.view(np.int32)).to(DEV).view(torch.uint32)
    d_macs = torch.from_numpy(macs).to(DEV)
    d_keys = torch.from_numpy(mandatory.view(view)).to(DEV).view(torch_t)
    ops = torch.zeros(nn, dtype=torch.int32, device=DEV)
    lib = c.lib()
    c._check(getattr(lib, "cs_rebalance_decision_essential_" + kt)(
        c._ptr(tree.prefixes), c._ptr(tree.child_offsets), c._ptr(tree.parents), c._ptr(d_counts), c._ptr(d_macs),
        cast(focus[0]), cast(focus[1]), C.c_uint32(bucket), c._ptr(ops), C.c_int(nn), c._stream()), "essential")
    status, conv = C.c_int(-1), C.c_int(-1)
    c._check(getattr(lib, "cs_enforce_keys_" + kt)(c._ptr(d_keys), C.c_int(mandatory.size), c._ptr(tree.prefixes),
                                                   c._ptr(tree.child_offsets), c._ptr(tree.parents), c._ptr(ops),
                                                   C.byref(status), c._stream()), "enforce")
    c._check(getattr(lib, "cs_protect_ancestors_" + kt)(c._ptr(tree.prefixes), c._ptr(tree.parents), c._ptr(ops),
                                                        C.c_int(nn), C.byref(conv), c._stream()), "protect")
    assert status.value == want_status
    assert bool(conv.value) == bool(want_conv)
    assert np.array_equal(ops.cpu().numpy(), want)
    assert len(set(want.tolist())) >= 3, "the case must exercise merges, stays and splits"
