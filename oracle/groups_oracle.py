"""TEST INFRASTRUCTURE - numpy restatement of the reference's target-group construction, used only by tests/.

computeFixedGroups (traversal/groups_gpu.cu:27-56) and computeGroupSplits (traversal/groups_gpu.cu:59-146 with the
kernels of traversal/groups_gpu.cuh:37-217).  The reference implements these for the GPU only; its own known-answer
test (test/unit_cuda/traversal/groups.cu:205-282) is compiled unmodified against the library by oracle/Makefile
(_ref/ref_gpu_unit_tests), this restatement extends the comparison to random inputs.  cbrt comes from numpy here and
from the CUDA math library on the device: the tests keep distances away from the criterion by more than an ulp."""
import numpy as np

MAX_LEVEL = 21


def fixed_groups(first, last, group_size):
    """groups_gpu.cu:27-56"""
    n = -(-(last - first) // group_size)
    return np.array([first + g * group_size for g in range(n)] + [last], dtype=np.uint32)


def group_splits(first, last, x, y, z, h, leaves, layout, lim, group_size, tol_factor):
    """ascending group boundaries from `first` to `last` (groups_gpu.cu:59-108 + groups_gpu.cuh:37-217)"""
    Tc = x.dtype.type
    Th = h.dtype.type if h is not None else Tc
    num_leaves = leaves.size - 1
    lens = [Tc(lim[2 * d + 1]) - Tc(lim[2 * d]) for d in range(3)]
    ilen = [Tc(1) / l for l in lens]
    min_extent = min(lens)
    n_fixed = -(-(last - first) // group_size)
    out = []
    for g in range(n_fixed):
        idx = np.minimum(first + g * group_size + np.arange(group_size), last - 1)
        # volume of the leaf of the lane's FIRST segment particle (groups_gpu.cuh:171-186 uses leafIdx[0] for every k)
        leaf = np.searchsorted(layout[:num_leaves], idx[:32], side="right") - 1
        rng = (leaves[leaf + 1] - leaves[leaf]).astype(np.uint64)
        level = np.array([MAX_LEVEL - (int(r).bit_length() - 1) // 3 for r in rng])
        cube = (1 << MAX_LEVEL) >> level
        half_unit = Th(0.5) * Th(1.0 / (1 << MAX_LEVEL)) * Th(1)
        size = cube.astype(Th) * half_unit
        vol = Th(8) * size * size * size
        node_volume = Th(min(Th(1), vol.min()))
        root = np.cbrt(node_volume)
        dist_crit = Tc(root * np.float32(tol_factor)) if Th is np.float32 else Tc(Th(root) * Th(np.float32(tol_factor)))
        crit_sq = Tc(dist_crit * dist_crit)
        px, py, pz = x[idx] * ilen[0], y[idx] * ilen[1], z[idx] * ilen[2]
        pr = (Tc(2) * h[idx].astype(Tc) / min_extent) if h is not None else np.ones(group_size, dtype=Tc)
        nx, ny, nz = (np.concatenate([a[1:], a[-1:]]) for a in (px, py, pz))
        dx, dy, dz = nx - px, ny - py, nz - pz
        d2 = dx * dx + (dy * dy + dz * dz)
        rr = pr * pr
        thr = np.where(rr < crit_sq, rr, crit_sq)
        split = d2 > thr
        out.append(first + g * group_size)
        out.extend(first + g * group_size + p + 1 for p in np.nonzero(split)[0])
    out.append(last)
    return np.array(out, dtype=np.uint32)
