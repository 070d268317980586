"""The LET rebalance kernels against the reference's explicit known-answer vectors
(test/unit/focus/octree_focus.cpp:26-186 rebalanceDecisionEssential + protectAncestors, :228-290 enforceKeys), through
the C ABI entry points that stand in for rebalanceDecisionEssentialGpu / protectAncestorsGpu / enforceKeysGpu
(focus/rebalance_gpu.h:27-79)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEYS = {"u32": (np.uint32, torch.uint32, 10), "u64": (np.uint64, torch.uint64, 21)}


def capi():
    from cstone_b200 import capi as c
    return c


def octree_maker(kt, *paths):
    """OctreeMaker (test/coord_samples... tree/cs_util.hpp:65-140): divide the node addressed by a path of octants"""
    np_t, _, max_level = KEYS[kt]
    leaves = [0, 1 << (3 * max_level)]
    for path in paths:
        key, level = 0, 0
        for digit in path:
            level += 1
            key += digit << (3 * (max_level - level))
        i = leaves.index(key)
        size = 1 << (3 * (max_level - level))
        assert leaves[i + 1] - key == size, "node to divide is not a leaf"
        leaves[i + 1:i + 1] = [key + s * (size // 8) for s in range(1, 8)]
    return np.array(leaves, dtype=np_t)


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


# (divisions, leaf counts, leaf macs, internal macs {prefix: mac}, focus leaf indices, expected leaf ops, converged)
ESSENTIAL_CASES = [
    (((), (0,), (7,)),
     [1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
     [0, 0, 1, 0, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0],
     [(1, 1), (0o10, 1), (0o17, 1)], (0, 8),
     [1, 1, 1, 8, 1, 1, 1, 1, 1, 1, 8, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], False),
    (((), (0,), (7,)),
     [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 2, 1, 0, 0, 0, 0],
     [0, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
     [(1, 1), (0o10, 1), (0o17, 1)], (0, 8),
     [1] * 22, True),
    (((), (0,), (7,)),
     [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 2, 1, 0, 0, 0, 0],
     [0, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0],
     [(1, 1), (0o10, 1), (0o17, 0)], (0, 8),
     [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], False),
    (((), (0,), (1,)),
     [1, 2, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 2, 1, 2, 1, 1, 2, 1, 1],
     [0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0],
     [(1, 1), (0o10, 1), (0o11, 0)], (2, 10),
     [1, 8, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 8, 1, 1, 1, 1, 1], False),
    (((), (6,), (7,)),
     [1] * 22,
     [1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
     [(1, 1), (0o16, 0), (0o17, 0)], (14, 22),
     [1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1], False),
]


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


def decode_placeholder(code, max_level):
    length = code.bit_length() - 1
    return (code ^ (1 << length)) << (3 * max_level - length)


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


CANCEL_MERGE, REBALANCE, FAILED = 1, 2, 3


@pytest.mark.parametrize("kt", ["u32", "u64"])
def test_key_enforcement_vectors(kt):
    """test/unit/focus/octree_focus.cpp:228-290: node ops of the 17-node tree divide().divide(1)"""
    c = capi()
    cstree = octree_maker(kt, (), (1,))
    tree = linked(kt, cstree)
    assert tree.num_nodes == 17
    start = [1, 1] + [0] * 15

    ops = torch.tensor(start, dtype=torch.int32, device=DEV)
    assert enforce(kt, tree, ops, 0o111) == CANCEL_MERGE
    assert ops.cpu().tolist() == [1] * 17

    ops = torch.tensor(start, dtype=torch.int32, device=DEV)
    assert enforce(kt, tree, ops, 0o1112) == REBALANCE
    assert ops.cpu().tolist() == [1] * 10 + [8] + [1] * 6

    ops = torch.tensor(start, dtype=torch.int32, device=DEV)
    assert enforce(kt, tree, ops, 0o101) == REBALANCE
    conv = C.c_int(-1)
    c._check(getattr(c.lib(), "cs_protect_ancestors_" + kt)(c._ptr(tree.prefixes), c._ptr(tree.parents), c._ptr(ops),
                                                            C.c_int(17), C.byref(conv), c._stream()), "protect")
    assert ops.cpu().tolist() == [1, 8] + [1] * 8 + [0] * 7

    ops = torch.tensor([1] * 10 + [0] * 7, dtype=torch.int32, device=DEV)
    assert enforce(kt, tree, ops, 0o101) == REBALANCE
    assert enforce(kt, tree, ops, 0o1011) == FAILED
    assert ops.cpu().tolist() == [1, 8] + [1] * 8 + [0] * 7
