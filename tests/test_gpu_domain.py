"""GPU parity of the Domain (cs_domain_* C ABI) against the reference's own cstone::Domain<KeyType,T,Cpu>:
  - golden fixtures generated from the unmodified reference (tests/golden/make_golden.py), three consecutive syncs;
  - when oracle/_ref/libcstone_ref.so is present, a live comparison at a larger size.
Everything is compared bit-for-bit: keys, particle order, box, global/focus leaf arrays, counts, layout, the linked
focus octree, node centres/sizes, halo flags and neighbour lists."""
import glob
import os

import numpy as np
import pytest
import torch

from _libs import key_of, real_of, ref, ref_domain_run
from _util import gaussian_particles, uniform_particles

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "domain_*.npz")))


def capi():
    from cstone_b200 import capi as c
    return c


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.cpu().numpy()


def drift(dom, d):
    """the particle update of oracle/ref_api.cpp::domainRank, applied to the domain-owned device arrays"""
    T = torch.float32 if dom.real == "f" else torch.float64
    keys = dom.field("keys").view(torch.int64 if dom.kt == "u64" else torch.int32)
    box = dom.box
    d = torch.tensor(d, dtype=T, device=DEV)
    sl = slice(dom.start_index, dom.end_index)  # only the assigned particles move, halos are refreshed by the sync
    keys = keys[sl]
    for name, shift, lo, hi in (("x", 3, box[0], box[1]), ("y", 6, box[2], box[3]), ("z", 9, box[4], box[5])):
        a = dom.field(name)[sl]
        digit = ((keys >> shift) & 7).to(T)
        a += d * (torch.tensor(0.5, dtype=T, device=DEV) - digit / torch.tensor(7, dtype=T, device=DEV))
        lo_t = torch.tensor(lo, dtype=T, device=DEV)
        hi_t = torch.nextafter(torch.tensor(hi, dtype=T, device=DEV), lo_t)
        torch.clamp(a, min=lo_t, max=hi_t, out=a)


def compare_state(dom, want, tag):
    assert (dom.start_index, dom.end_index) == (want["start"], want["end"]), tag
    assert dom.n_particles_with_halos == want["keys"].size, tag
    T = real_of(dom.combo)
    assert np.array_equal(np.array(dom.box, dtype=T), want["box"].astype(T)), (tag, dom.box, want["box"])
    pairs = [("keys", "keys"), ("x", "x"), ("y", "y"), ("z", "z"), ("h", "h"), ("focus_leaves", "focus_leaves"),
             ("global_leaves", "global_leaves"), ("focus_leaf_counts", "focus_counts"), ("layout", "layout"),
             ("prefixes", "prefixes"), ("child_offsets", "child_offsets"), ("internal_to_leaf", "internal_to_leaf"),
             ("leaf_to_internal", "leaf_to_internal"), ("level_range", "level_range"), ("halo_flags", "flags")]
    for ours, theirs in pairs:
        got = host(dom.field(ours))
        assert got.shape == want[theirs].shape, (tag, ours, got.shape, want[theirs].shape)
        assert np.array_equal(got, want[theirs]), (tag, ours)
    nn = want["prefixes"].size
    assert np.array_equal(host(dom.field("parents"))[:(nn - 1) // 8], want["parents"]), tag
    assert np.array_equal(host(dom.field("geo_centers")).ravel(), want["centers"]), tag
    assert np.array_equal(host(dom.field("geo_sizes")).ravel(), want["sizes"]), tag


def compare_neighbors(dom, want_nc, want_flat, ngmax, tag):
    nb, nc = dom.find_neighbors(ngmax)
    nb, nc = host(nb), host(nc)
    assert np.array_equal(nc, want_nc), tag
    m = np.arange(ngmax)[None, :] < np.minimum(nc, ngmax)[:, None]
    assert np.array_equal(nb[m], want_flat), tag


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_domain_against_reference_golden(path):
    g = np.load(path)
    combo = str(g["combo"])
    ngmax = int(g["ngmax"])
    dom = capi().Domain(0, 1, int(g["bucket"]), int(g["bucket_focus"]), 0.5, g["lim"], g["bnd"], key=key_of(combo),
                        real=combo[3])
    for ns in (1, 2, 3):
        if ns == 1:
            dom.sync(dev(g["x0"]), dev(g["y0"]), dev(g["z0"]), dev(g["h0"]))
        else:
            drift(dom, float(g["moves"][ns - 2]))
            dom.sync()
        want = {k[3:]: g[k] for k in g.files if k.startswith(f"s{ns}_")}
        want["start"], want["end"] = [int(v) for v in want["start_end"]]
        compare_state(dom, want, f"sync {ns}")
        compare_neighbors(dom, want["neighbors_count"], want["neighbors_flat"], ngmax, f"sync {ns}")


@pytest.mark.skipif(ref() is None, reason="oracle/_ref/libcstone_ref.so not built")
@pytest.mark.parametrize("combo,dist,pbc,bucket,bucket_focus",
                         [("u64d", "uniform", 0, 64, 64), ("u64d", "gaussian", 1, 128, 16),
                          ("u32f", "gaussian", 0, 64, 8), ("u64f", "uniform", 1, 1024, 32)])
def test_domain_against_reference_live(combo, dist, pbc, bucket, bucket_focus):
    T = real_of(combo)
    n = 200000
    lim, bnd = (-1, 1, -1, 1, -1, 1), (pbc, pbc, pbc)
    x, y, z = gaussian_particles(n, T, 5) if dist == "gaussian" else uniform_particles(n, T, 5, -1, 1)
    h = np.full(n, 0.5 * np.cbrt(3.0 * 40 * 8 / (4 * np.pi * n)), dtype=T)
    moves = np.array([0.02, 0.01], dtype=T)
    ngmax = 100
    dom = capi().Domain(0, 1, bucket, bucket_focus, 0.5, lim, bnd, key=key_of(combo), real=combo[3])
    for ns in (1, 2, 3):
        if ns == 1:
            dom.sync(dev(x), dev(y), dev(z), dev(h))
        else:
            drift(dom, float(moves[ns - 2]))
            dom.sync()
        r = ref_domain_run(combo, 1, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, [0, n], num_syncs=ns,
                           ngmax=ngmax, moves=moves)[0]
        compare_state(dom, r, f"sync {ns}")
        nc = r["neighbors_count"]
        m = np.arange(ngmax)[None, :] < np.minimum(nc, ngmax)[:, None]
        compare_neighbors(dom, nc, r["neighbors"].reshape(nc.size, ngmax)[m], ngmax, f"sync {ns}")


def test_domain_host_input_and_remove_key():
    """host-buffer entry (the e2e path) and removeKey handling (sfc/sfc.hpp:274, domain.hpp:467-468)"""
    n = 50000
    x, y, z = uniform_particles(n, np.float64, 9)
    h = np.full(n, 0.01)
    keys = np.zeros(n, dtype=np.uint64)
    keys[::10] = np.uint64(1) << np.uint64(63)
    lim, bnd = (0, 1, 0, 1, 0, 1), (1, 1, 1)
    dom = capi().Domain(0, 1, 64, 64, 0.5, lim, bnd)
    pinned = [torch.from_numpy(a).pin_memory() for a in (x, y, z, h)]
    dom.sync(*pinned, keys=torch.from_numpy(keys).pin_memory())
    keep = np.ones(n, dtype=bool)
    keep[::10] = False
    assert dom.end_index - dom.start_index == keep.sum() == dom.n_particles_with_halos
    got = np.sort(host(dom.field("x")))
    assert np.array_equal(got, np.sort(x[keep]))
    k = host(dom.field("keys"))
    assert np.all(k[1:] >= k[:-1])
    out = [torch.empty(dom.n_particles_with_halos, dtype=torch.float64).pin_memory() for _ in range(4)]
    kout = torch.empty(dom.n_particles_with_halos, dtype=torch.uint64).pin_memory()
    dom.download(*out, kout)
    torch.cuda.synchronize()
    assert np.array_equal(out[0].numpy(), host(dom.field("x")))
    assert np.array_equal(kout.numpy(), k)


def test_domain_rejects_unsupported_configurations():
    with pytest.raises(capi().CstoneError):
        capi().Domain(0, 1, 8, 64, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0))  # bucketSize < bucketSizeFocus
    with pytest.raises(capi().CstoneError):
        capi().Domain(2, 2, 64, 64, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0))  # rank out of range
    # a multi-rank domain without a communicator must refuse to sync
    dom = capi().Domain(0, 2, 64, 64, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0))
    v = torch.rand(100, dtype=torch.float64, device=DEV)
    with pytest.raises(capi().CstoneError):
        dom.sync(v, v, v, v)
