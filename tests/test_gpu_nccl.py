"""Multi-rank Domain over the REAL transport: one process per GPU, NCCL collectives and the peer-memory particle
exchange (CUDA IPC), launched with torch.distributed.run.  Needs at least two GPUs (skipped otherwise; the same C++
code is covered on one GPU by tests/test_gpu_multirank.py through the local communicator).  Every array is compared
bit for bit with the unmodified reference run with the same number of ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from _libs import ref, ref_domain_run
from _util import const_h, uniform_particles

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("P,pbc,peer_push", [(2, 1, True), (2, 0, False)])
def test_nccl_domain_matches_reference(tmp_path, P, pbc, peer_push):
    if torch.cuda.device_count() < P:
        pytest.skip(f"needs {P} GPUs")
    n_per, bucket, bucket_focus, num_syncs = 20000, 64, 8, 2
    n = n_per * P
    x, y, z = uniform_particles(n, np.float64, 21)
    h = const_h(n, 40, np.float64, 1.0)
    lim, bnd = (0, 1, 0, 1, 0, 1), (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    inputs = tmp_path / "inputs.npz"
    np.savez(inputs, x=x, y=y, z=z, h=h, offsets=np.array(offsets), bucket=bucket, bucket_focus=bucket_focus,
             lim=np.array(lim, dtype=np.float64), bnd=np.array(bnd), num_syncs=num_syncs)
    env = dict(os.environ)
    if not peer_push:
        env["CSB_NO_PEER_PUSH"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P}", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "_nccl_worker.py"),
           str(tmp_path), str(inputs)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]

    want = ref_domain_run("u64d", P, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=num_syncs,
                          ngmax=64)
    for rank in range(P):
        g, w = np.load(tmp_path / f"rank{rank}.npz"), want[rank]
        assert (int(g["start"]), int(g["end"])) == (w["start"], w["end"]), rank
        for k in ("keys", "x", "y", "z", "h", "focus_leaves", "layout", "global_leaves"):
            assert np.array_equal(g[k], w[k]), (rank, k)
        assert bool(g["halo_field_ok"]), rank
        assert np.array_equal(g["nc"], w["neighbors_count"]), rank
        m = np.arange(64)[None, :] < np.minimum(w["neighbors_count"], 64)[:, None]
        assert np.array_equal(g["nb"][m], w["neighbors"].reshape(-1, 64)[m]), rank
        assert int(g["bytes_sent"]) > 0
