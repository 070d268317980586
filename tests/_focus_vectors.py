"""The reference's explicit known-answer vectors for the LET rebalance decisions
(test/unit/focus/octree_focus.cpp:26-290), shared by the oracle (CPU) and the kernel (GPU) tests."""
import numpy as np

KEY_TYPES = {"u32": (np.uint32, 10), "u64": (np.uint64, 21)}


def octree_maker(kt, *paths):
    """OctreeMaker (test/coord_samples... tree/cs_util.hpp:65-140): divide the node addressed by a path of octants"""
    np_t, max_level = KEY_TYPES[kt]
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



def decode_placeholder(code, max_level):
    length = code.bit_length() - 1
    return (code ^ (1 << length)) << (3 * max_level - length)



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



CANCEL_MERGE, REBALANCE, FAILED = 1, 2, 3
#: node ops of the 17-node tree divide().divide(1): (start ops, [placeholder-bit keys to enforce], expected statuses,
#: protectAncestors afterwards, expected ops)   octree_focus.cpp:228-290
ENFORCE_CASES = [
    ([1, 1] + [0] * 15, [0o111], [CANCEL_MERGE], False, [1] * 17),
    ([1, 1] + [0] * 15, [0o1112], [REBALANCE], False, [1] * 10 + [8] + [1] * 6),
    ([1, 1] + [0] * 15, [0o101], [REBALANCE], True, [1, 8] + [1] * 8 + [0] * 7),
    ([1] * 10 + [0] * 7, [0o101, 0o1011], [REBALANCE, FAILED], False, [1, 8] + [1] * 8 + [0] * 7),
]
