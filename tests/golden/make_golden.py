"""Generates tests/golden/domain_*.npz from the UNMODIFIED reference (oracle/_ref/libcstone_ref.so =
/root/reference headers, execution::Cpu, single rank through the thread-backed MPI stand-in).

Each file holds the inputs of one case and the complete observable state of cstone::Domain after 1, 2 and 3 calls to
sync() (particles drift between calls exactly as in oracle/ref_api.cpp::domainRank).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs  # noqa: E402
from _util import gaussian_particles, plummer_particles, uniform_particles  # noqa: E402

CASES = {
    # name: (combo, distribution, n, lim, bnd, bucket, bucketFocus, h-neighbours)
    "uniform_u64d_open": ("u64d", "uniform", 4000, (0, 1, 0, 1, 0, 1), (0, 0, 0), 64, 8, 30),
    "gaussian_u64d_pbc": ("u64d", "gaussian", 4000, (-1, 1, -1, 1, -1, 1), (1, 1, 1), 32, 8, 20),
    "uniform_u32f_open": ("u32f", "uniform", 4000, (0, 1, 0, 1, 0, 1), (0, 0, 0), 16, 16, 30),
    "plummer_u64f_open": ("u64f", "plummer", 4000, (-1, 1, -1, 1, -1, 1), (0, 0, 0), 64, 16, 10),
}
MOVES = [0.01, 0.03]
KEEP = ["keys", "x", "y", "z", "h", "focus_leaves", "global_leaves", "focus_counts", "layout", "prefixes",
        "child_offsets", "parents", "internal_to_leaf", "leaf_to_internal", "level_range", "centers", "sizes", "flags",
        "neighbors_count"]


def inputs(combo, dist, n, lim):
    T = _libs.real_of(combo)
    if dist == "uniform":
        x, y, z = uniform_particles(n, T, 42, lim[0], lim[1])
    elif dist == "gaussian":
        x, y, z = gaussian_particles(n, T, 42, lim[0], lim[1])
    else:
        x, y, z = plummer_particles(n, T, 42)
    vol = 1.0 if dist == "uniform" else 0.5
    return x, y, z


def main():
    assert _libs.ref_lib() is not None, "build oracle/_ref first (make -C oracle ref)"
    for name, (combo, dist, n, lim, bnd, bucket, bucket_focus, ng) in CASES.items():
        T = _libs.real_of(combo)
        x, y, z = inputs(combo, dist, n, lim)
        h = np.full(n, 0.5 * np.cbrt(3.0 * ng / (4 * np.pi * n)) * (lim[1] - lim[0]), dtype=T)
        moves = np.array(MOVES, dtype=T)
        out = dict(x0=x, y0=y, z0=z, h0=h, moves=moves, lim=np.array(lim, dtype=np.float64),
                   bnd=np.array(bnd, dtype=np.int32), bucket=bucket, bucket_focus=bucket_focus, ngmax=64,
                   combo=np.array(combo))
        for ns in (1, 2, 3):
            r = _libs.ref_domain_run(combo, 1, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, [0, n], num_syncs=ns,
                                     ngmax=64, moves=moves)[0]
            for k in KEEP:
                out[f"s{ns}_{k}"] = r[k]
            nc = r["neighbors_count"]
            nb = r["neighbors"].reshape(nc.size, 64)
            m = np.arange(64)[None, :] < np.minimum(nc, 64)[:, None]
            out[f"s{ns}_neighbors_flat"] = nb[m]
            out[f"s{ns}_box"] = r["box"]
            out[f"s{ns}_start_end"] = np.array([r["start"], r["end"]], dtype=np.int64)
        path = os.path.join(HERE, f"domain_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
