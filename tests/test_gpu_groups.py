"""Target particle groups (-m gpu): cs_compute_fixed_groups / cs_group_splits_* against the numpy restatement
(oracle/groups_oracle.py) and the reference's own known answers (test/unit_cuda/traversal/groups.cu:27-41,205-282; the
unmodified test file itself runs inside tests/test_gpu_dropin.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
DEV = "cuda:0"


def capi():
    from cstone_b200 import capi as c
    return c


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view({8: np.int64, 4: np.int32}[a.dtype.itemsize])).to(DEV)


def gpu_group_splits(first, last, x, y, z, h, leaves, layout, lim, group_size, tol_factor):
    c = capi()
    sfx = {("float64", "float64"): "dd", ("float64", "float32"): "df", ("float32", "float32"): "ff"}[
        (x.dtype.name, h.dtype.name)]
    dx, dy, dz, dh, dl, dla = dev(x), dev(y), dev(z), dev(h), dev(leaves), dev(layout)
    lim_c, bnd_c = (C.c_double * 6)(*lim), (C.c_int * 3)(0, 0, 0)
    n = C.c_uint32(0)
    V = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = getattr(c.lib(), "cs_group_splits_begin_" + sfx)(
        C.c_uint32(first), C.c_uint32(last), V(dx), V(dy), V(dz), V(dh), V(dl), C.c_int(leaves.size - 1), V(dla), lim_c,
        bnd_c, C.c_uint32(group_size), C.c_float(tol_factor), C.byref(n), None)
    assert st == 0, c.lib().cs_last_error()
    groups = torch.zeros(n.value + 1, dtype=torch.int32, device=DEV)
    st = c.lib().cs_group_splits_finish(C.c_uint32(first), C.c_uint32(last), C.c_uint32(group_size), V(groups), None)
    assert st == 0
    torch.cuda.synchronize()
    return groups.cpu().numpy().view(np.uint32)


def test_fixed_groups_reference_vector():
    """test/unit_cuda/traversal/groups.cu:27-41: groupSize 8, first 4, last 34 -> {4, 12, 20, 28, 34}"""
    import groups_oracle
    g = torch.zeros(5, dtype=torch.int32, device=DEV)
    assert capi().lib().cs_compute_fixed_groups(C.c_uint32(4), C.c_uint32(34), C.c_uint32(8),
                                                C.c_void_p(g.data_ptr()), None) == 0
    torch.cuda.synchronize()
    assert g.cpu().tolist() == [4, 12, 20, 28, 34]
    assert groups_oracle.fixed_groups(4, 34, 8).tolist() == [4, 12, 20, 28, 34]


def _two_level_tree():
    """OctreeMaker{}.divide().divide(2): root split, then child 2 split (15 leaves), as in groups.cu:221"""
    r = np.uint64(1) << np.uint64(63)
    o = r >> np.uint64(3)
    leaves = [np.uint64(0), o, np.uint64(2) * o]
    leaves += [np.uint64(2) * o + np.uint64(k) * (o >> np.uint64(3)) for k in range(1, 8)]
    leaves += [np.uint64(k) * o for k in range(3, 8)] + [r]
    return np.array(leaves, dtype=np.uint64)


def test_group_splits_reference_known_answer():
    """test/unit_cuda/traversal/groups.cu:205-282: fixed groups (4, 68, 128) become (4, 6, 68, 75, 128): one cut by
    distance, one by interaction radius"""
    import groups_oracle
    first, last, G = 4, 128, 64
    leaves = _two_level_tree()
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
    tol = np.float32(np.sqrt(3.0) / dist_crit * 1.01)
    want = [4, 6, 68, 75, 128]
    assert groups_oracle.group_splits(first, last, x, y, z, h, leaves, layout, lim, G, tol).tolist() == want
    assert gpu_group_splits(first, last, x, y, z, h, leaves, layout, lim, G, tol).tolist() == want


@pytest.mark.parametrize("types", [(np.float64, np.float64), (np.float64, np.float32), (np.float32, np.float32)])
@pytest.mark.parametrize("group_size", [32, 64])
@pytest.mark.parametrize("first,n", [(0, 5000), (7, 4099), (3, 40), (0, 64)])
def test_group_splits_random_walks(types, group_size, first, n):
    """particles along a random walk whose steps are either well below or well above the distance criterion, radii
    either far above it or small enough to cut; leaves of three different sizes so that the smallest leaf of a group
    (taken from the first 32 particles of each lane, as the reference does) changes from group to group"""
    import groups_oracle
    Tc, Th = types
    rng = np.random.default_rng(n + group_size)
    last = first + n
    leaves = _two_level_tree()
    num_leaves = leaves.size - 1
    cuts = np.sort(rng.choice(np.arange(1, last), size=num_leaves - 1, replace=False))
    layout = np.concatenate([[0], cuts, [last]]).astype(np.uint32)
    crit_small = 1.0 / 4  # cbrt of the volume of the small leaves (1/64 of the unit box)
    tol = np.float32(0.05)
    step = np.where(rng.random(last) < 0.1, 6.0, 0.2) * crit_small * float(tol) / np.sqrt(3.0)
    walk = np.cumsum(step)
    x = (walk % 1.0).astype(Tc)
    y = ((0.5 * walk) % 1.0).astype(Tc)
    z = ((0.25 * walk) % 1.0).astype(Tc)
    h = np.where(rng.random(last) < 0.05, 1e-4, 10.0).astype(Th)
    lim = (0, 1, 0, 1, 0, 1)
    want = groups_oracle.group_splits(first, last, x, y, z, h, leaves, layout, lim, group_size, tol)
    got = gpu_group_splits(first, last, x, y, z, h, leaves, layout, lim, group_size, tol)
    assert np.array_equal(got, want), (got[:12], want[:12])
    assert want[0] == first and want[-1] == last and bool((np.diff(want.astype(np.int64)) > 0).all())
    assert want.size > -(-n // group_size) + 1, "the test is meant to produce splits"
