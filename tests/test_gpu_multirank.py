"""Multi-rank Domain parity (-m gpu): P ranks run as threads of this process over the library's thread-backed
communicator (cs_comm_create_local), all on cuda:0, and are compared rank by rank with the UNMODIFIED reference run
with P ranks (oracle/_ref, thread-backed MPI stand-in).  The same C++ code runs over NCCL with one process per GPU
(bench.py --gpus N)."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

from _libs import key_of, real_of, ref, ref_domain_run, ref_lib
from _util import const_h, gaussian_particles, uniform_particles

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def capi():
    from cstone_b200 import capi as c
    return c


def run_ranks(P, fn, world=None):
    """run fn(rank) on P threads (each with its own CUDA stream); re-raise the first failure.  A failing rank aborts
    the local world so that the ranks waiting for it at a collective return an error instead of blocking forever."""
    errors = [None] * P
    results = [None] * P

    def body(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream(device=DEV)):
                results[r] = fn(r)
                torch.cuda.current_stream().synchronize()
        except BaseException as e:  # noqa: BLE001
            errors[r] = e
            if world is not None:
                world.abort()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    if any(t.is_alive() for t in threads) and world is not None:
        world.abort()
        for t in threads:
            t.join(timeout=10)
    first = [e for e in errors if e is not None and "another rank failed" not in str(e)]
    if first:
        raise first[0]
    for e in errors:
        if e is not None:
            raise e
    for t in threads:
        assert not t.is_alive(), "rank thread hung"
    return results


def make_particles(combo, n, dist, seed):
    T = real_of(combo)
    if dist == "gaussian":
        x, y, z = gaussian_particles(n, T, seed)
        lim = (-1, 1, -1, 1, -1, 1)
    else:
        x, y, z = uniform_particles(n, T, seed)
        lim = (0, 1, 0, 1, 0, 1)
    return x, y, z, lim


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("combo,P,dist,pbc,bucket,bucket_focus", [
    ("u64d", 2, "uniform", 0, 64, 8),
    ("u64d", 3, "gaussian", 1, 128, 16),
    ("u64d", 4, "uniform", 1, 64, 8),
    ("u64f", 4, "gaussian", 0, 64, 64),
    ("u32f", 2, "uniform", 0, 32, 8),
    ("u64d", 8, "uniform", 0, 256, 16),
])
def test_assigned_particles_match_reference(combo, P, dist, pbc, bucket, bucket_focus):
    """after sync every rank owns exactly the reference's particles in the reference's order: keys, x, y, z, h of
    [startIndex, endIndex), the global tree and the coordinate box are bit-identical (two consecutive syncs)"""
    n_per = 6000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, dist, 3)
    T = real_of(combo)
    h = const_h(n, 40, T, 8.0 if dist == "gaussian" else 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    for num_syncs in (1, 2):
        want = ref_domain_run(combo, P, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=num_syncs)
        world = capi().LocalWorld(P)

        def rank_body(r):
            c = capi()
            comm = world.comm(r)
            dom = c.Domain(r, P, bucket, bucket_focus, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV,
                           comm=comm)
            sl = slice(offsets[r], offsets[r + 1])
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
            dom.sync(to(x), to(y), to(z), to(h))
            for _ in range(num_syncs - 1):
                dom.sync()
            s, e = dom.start_index, dom.end_index
            out = {k: dom.field(k)[s:e].cpu().numpy() for k in ("keys", "x", "y", "z", "h")}
            out["global_leaves"] = dom.field("global_leaves").cpu().numpy()
            out["box"] = dom.box
            dom.close()
            comm.close()
            return out

        got = run_ranks(P, rank_body, world)
        for r in range(P):
            w = want[r]
            s, e = w["start"], w["end"]
            assert np.array_equal(got[r]["global_leaves"], w["global_leaves"]), (r, "global leaves")
            assert np.array_equal(np.asarray(got[r]["box"]), w["box"]), (r, "box")
            for k in ("keys", "x", "y", "z", "h"):
                assert np.array_equal(got[r][k], w[k][s:e]), (r, k, num_syncs)


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("combo,P,dist,pbc,bucket,bucket_focus", [
    ("u64d", 2, "uniform", 0, 64, 8),
    ("u64d", 2, "uniform", 1, 64, 8),
    ("u64d", 3, "gaussian", 1, 128, 16),
    ("u64d", 4, "uniform", 1, 64, 8),
    ("u64f", 4, "gaussian", 0, 64, 64),
    ("u32f", 2, "uniform", 0, 32, 8),
    ("u64d", 8, "uniform", 0, 256, 16),
    ("u64d", 5, "gaussian", 0, 1024, 32),
])
def test_full_domain_with_halos_matches_reference(combo, P, dist, pbc, bucket, bucket_focus):
    """everything Domain::sync leaves behind on every rank is bit-identical with the reference: start/end/size, the
    complete key and coordinate arrays including halos, the LET (leaves, counts, linked tree, halo flags), the layout"""
    n_per = 5000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, dist, 5)
    T = real_of(combo)
    h = const_h(n, 40, T, 8.0 if dist == "gaussian" else 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    tree_fields = ("focus_leaves", "layout", "prefixes", "child_offsets", "internal_to_leaf", "leaf_to_internal",
                   "level_range")
    for num_syncs in (1, 3):
        want = ref_domain_run(combo, P, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=num_syncs,
                              ngmax=64)
        world = capi().LocalWorld(P)

        def rank_body(r):
            c = capi()
            comm = world.comm(r)
            dom = c.Domain(r, P, bucket, bucket_focus, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV,
                           comm=comm)
            sl = slice(offsets[r], offsets[r + 1])
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
            dom.sync(to(x), to(y), to(z), to(h))
            for _ in range(num_syncs - 1):
                dom.sync()
            out = {k: dom.field(k).cpu().numpy() for k in ("keys", "x", "y", "z", "h") + tree_fields}
            out["focus_counts"] = dom.field("focus_leaf_counts").cpu().numpy()
            out["flags"] = dom.field("halo_flags").cpu().numpy()
            out["parents"] = dom.field("parents").cpu().numpy()
            out["start"], out["end"], out["size"] = dom.start_index, dom.end_index, dom.n_particles_with_halos
            nb, nc = dom.find_neighbors(64)
            out["nb"], out["nc"] = nb.cpu().numpy(), nc.cpu().numpy()
            dom.close()
            comm.close()
            return out

        got = run_ranks(P, rank_body, world)
        for r in range(P):
            w, g = want[r], got[r]
            assert (g["start"], g["end"], g["size"]) == (w["start"], w["end"], w["keys"].size), (r, num_syncs)
            for k in tree_fields + ("focus_counts",):
                assert np.array_equal(g[k], w[k]), (r, k, num_syncs)
            nn = w["prefixes"].size
            assert np.array_equal(g["parents"][:(nn - 1) // 8], w["parents"]), (r, "parents")
            assert np.array_equal(g["flags"], w["flags"]), (r, "halo flags", num_syncs)
            for k in ("keys", "x", "y", "z", "h"):
                assert np.array_equal(g[k], w[k]), (r, k, num_syncs)
            assert np.array_equal(g["nc"], w["neighbors_count"]), (r, "neighbour counts")
            m = np.arange(64)[None, :] < np.minimum(w["neighbors_count"], 64)[:, None]
            assert np.array_equal(g["nb"][m], w["neighbors"].reshape(-1, 64)[m]), (r, "neighbour lists")


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("combo,P,dist,pbc,bucket,bucket_focus", [
    ("u64d", 2, "uniform", 1, 64, 8),
    ("u64d", 4, "gaussian", 0, 128, 16),
    ("u64f", 3, "uniform", 0, 64, 16),
])
def test_drifting_particles_match_reference(combo, P, dist, pbc, bucket, bucket_focus):
    """particles move between syncs (the deterministic drift of oracle/ref_api.cpp), so later syncs re-assign and
    exchange particles, the LET changes under the peers' feet (treelet keys get rejected) and halos are re-discovered;
    after four syncs every rank still holds the reference's arrays bit for bit"""
    from test_gpu_domain import drift

    n_per = 4000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, dist, 11)
    T = real_of(combo)
    h = const_h(n, 40, T, 8.0 if dist == "gaussian" else 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    moves = np.array([0.02, 0.05, 0.01], dtype=T)
    num_syncs = moves.size + 1
    want = ref_domain_run(combo, P, bucket, bucket_focus, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=num_syncs,
                          moves=moves)
    world = capi().LocalWorld(P)

    def rank_body(r):
        c = capi()
        comm = world.comm(r)
        dom = c.Domain(r, P, bucket, bucket_focus, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV,
                       comm=comm)
        sl = slice(offsets[r], offsets[r + 1])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
        dom.sync(to(x), to(y), to(z), to(h))
        for s in range(1, num_syncs):
            drift(dom, float(moves[s - 1]))
            dom.sync()
        out = {k: dom.field(k).cpu().numpy() for k in ("keys", "x", "y", "z", "h", "focus_leaves", "layout")}
        out["start"], out["end"] = dom.start_index, dom.end_index
        dom.close()
        comm.close()
        return out

    got = run_ranks(P, rank_body, world)
    for r in range(P):
        w, g = want[r], got[r]
        assert (g["start"], g["end"]) == (w["start"], w["end"]), r
        for k in ("focus_leaves", "layout", "keys", "x", "y", "z", "h"):
            assert np.array_equal(g[k], w[k]), (r, k)


@pytest.mark.parametrize("combo,P,pbc", [("u64d", 2, 0), ("u64d", 4, 1), ("u32f", 3, 0)])
def test_exchange_halos_of_client_fields(combo, P, pbc):
    """Domain::exchangeHalos (domain.hpp:332-337): fields the client computed for its assigned particles arrive in the
    halo rows of the ranks that need them.  The fields are functions of the coordinates, and the halo coordinates were
    already proven identical with the reference, so the expected halo values follow from them exactly."""
    n_per = 5000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, "uniform", 9)
    T = real_of(combo)
    tT = torch.float32 if T == np.float32 else torch.float64
    h = const_h(n, 40, T, 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    world = capi().LocalWorld(P)

    def make_fields(dom):
        fx, fy, fz = dom.field("x"), dom.field("y"), dom.field("z")
        rho = fx * 2 + fy                                        # one real per particle
        vel = torch.stack([fz, fx - fy, fy * fz], dim=1).contiguous()   # three reals per particle
        tag = (fx * 1000).to(torch.int32)                        # one 32-bit integer per particle
        return rho, vel, tag

    def rank_body(r):
        c = capi()
        comm = world.comm(r)
        dom = c.Domain(r, P, 64, 8, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV, comm=comm)
        sl = slice(offsets[r], offsets[r + 1])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
        dom.sync(to(x), to(y), to(z), to(h))
        want = make_fields(dom)
        s, e, nh = dom.start_index, dom.end_index, dom.n_particles_with_halos
        fields = []
        for w in want:
            f = torch.full_like(w, -7)
            f[s:e] = w[s:e]
            fields.append(f)
        dom.exchange_halos(*fields)
        torch.cuda.current_stream().synchronize()
        ok = all(torch.equal(f, w) for f, w in zip(fields, want))
        num_halos = nh - (e - s)
        assert tT == want[0].dtype
        dom.close()
        comm.close()
        return ok, num_halos

    res = run_ranks(P, rank_body, world)
    assert all(ok for ok, _ in res)
    assert all(nh > 0 for _, nh in res), "the test needs halos to be meaningful"


@pytest.mark.parametrize("combo,P,pbc,dist", [("u64d", 1, 0, "uniform"), ("u64d", 2, 0, "uniform"),
                                              ("u64d", 4, 1, "gaussian"), ("u32f", 3, 0, "uniform"),
                                              ("u64f", 8, 0, "uniform")])
def test_reapply_sync_moves_client_fields_with_their_particles(combo, P, pbc, dist):
    """Domain::reapplySync (domain.hpp:297-329, the ExchangeLog replay of index_ranges.hpp:186-210): fields that did not
    take part in sync() are sent through the recorded exchange and reordered afterwards.  The fields are functions of
    the coordinates computed BEFORE the sync (in the old order, on the old owners); after sync + reapplySync every
    assigned row must equal the same function of the row's new coordinates, which sync() moved itself and which were
    proven identical with the reference above.  Covers the first sync (input order -> SFC order), later syncs with
    particles that migrate between ranks, and two replays of the same log."""
    from test_gpu_domain import drift

    n_per = 6000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, dist, 21)
    T = real_of(combo)
    h = const_h(n, 40, T, 8.0 if dist == "gaussian" else 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    world = capi().LocalWorld(P)

    def make_fields(fx, fy, fz):
        rho = fx * 2 + fy
        vel = torch.stack([fz, fx - fy, fy * fz], dim=1).contiguous()
        tag = (fx * 100000).to(torch.int32)
        quad = torch.stack([fx, fy, fz, fx + fz], dim=1).contiguous()
        return rho, vel, tag, quad

    def rank_body(r):
        c = capi()
        comm = world.comm(r)
        dom = c.Domain(r, P, 64, 8, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV, comm=comm)
        sl = slice(offsets[r], offsets[r + 1])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
        dx, dy, dz = to(x), to(y), to(z)
        before = make_fields(dx, dy, dz)
        dom.sync(dx, dy, dz, to(h))
        moved = 0
        for step in range(4):
            after = dom.reapply_sync(*before)
            again = dom.reapply_sync(*before[:2])
            s, e = dom.start_index, dom.end_index
            want = make_fields(dom.field("x"), dom.field("y"), dom.field("z"))
            for a, w in zip(after, want):
                assert a.shape[0] == dom.n_particles_with_halos
                assert torch.equal(a[s:e], w[s:e]), (r, step)
            for a, b in zip(after, again):
                assert torch.equal(a[s:e], b[s:e]), (r, step)
            if step == 3:
                break
            drift(dom, 0.08)
            # the client's fields in the layout it holds now (assigned + halo rows), changed along with the particles
            before = make_fields(dom.field("x").clone(), dom.field("y").clone(), dom.field("z").clone())
            n_before = dom.end_index - dom.start_index
            dom.sync()
            moved += abs((dom.end_index - dom.start_index) - n_before)
        dom.close()
        comm.close()
        return moved

    run_ranks(P, rank_body, world)


def test_reapply_sync_needs_a_sync_to_replay():
    c = capi()
    dom = c.Domain(0, 1, 64, 8, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0), key="u64", real="d", device=DEV)
    with pytest.raises(RuntimeError):
        dom.reapply_sync(torch.zeros(4, device=DEV))
    dom.close()


@pytest.mark.skipif(ref() is None, reason="needs oracle/_ref (built where /root/reference exists)")
@pytest.mark.parametrize("combo,P,pbc,factor", [("u64d", 3, 0, 1.5), ("u64f", 2, 1, 2.0)])
def test_halo_factor_matches_reference(combo, P, pbc, factor):
    """Domain::setHaloFactor (domain.hpp:365): the halo search boxes are built with factor * 2h
    (octree_focus_mpi.hpp:542), so more leaves are flagged and more halo particles arrive.  Start/end/size, halo flags,
    layout and the complete arrays with halos equal the reference's run with the same factor - and differ from the run
    with factor 1."""
    n_per = 5000
    n = n_per * P
    x, y, z, lim = make_particles(combo, n, "uniform", 17)
    T = real_of(combo)
    h = const_h(n, 40, T, 1.0)
    bnd = (pbc, pbc, pbc)
    offsets = [n_per * r for r in range(P + 1)]
    base = ref_domain_run(combo, P, 64, 8, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=2)
    ref_lib().ref_set_halo_factor(C.c_float(factor))
    try:
        want = ref_domain_run(combo, P, 64, 8, 0.5, lim, bnd, x, y, z, h, offsets, num_syncs=2)
    finally:
        ref_lib().ref_set_halo_factor(C.c_float(1.0))
    assert any(want[r]["keys"].size > base[r]["keys"].size for r in range(P)), "the factor must deepen the halo layer"
    world = capi().LocalWorld(P)

    def rank_body(r):
        c = capi()
        comm = world.comm(r)
        dom = c.Domain(r, P, 64, 8, 0.5, lim, bnd, key=key_of(combo), real=combo[-1], device=DEV, comm=comm)
        dom.set_halo_factor(factor)
        sl = slice(offsets[r], offsets[r + 1])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a[sl])).to(DEV)  # noqa: E731
        dom.sync(to(x), to(y), to(z), to(h))
        dom.sync()
        out = {k: dom.field(k).cpu().numpy() for k in ("keys", "x", "y", "z", "h", "layout", "halo_flags")}
        out["start"], out["end"], out["size"] = dom.start_index, dom.end_index, dom.n_particles_with_halos
        dom.close()
        comm.close()
        return out

    got = run_ranks(P, rank_body, world)
    for r in range(P):
        w, g = want[r], got[r]
        assert (g["start"], g["end"], g["size"]) == (w["start"], w["end"], w["keys"].size), r
        assert np.array_equal(g["halo_flags"], w["flags"]), (r, "halo flags")
        for k in ("layout", "keys", "x", "y", "z", "h"):
            assert np.array_equal(g[k], w[k]), (r, k)
