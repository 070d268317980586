"""CPU tests of the multi-rank HOST logic (no GPU, no compute calls into the CUDA kernels):

* the SFC domain decomposition functions exported by the C ABI (cs_uniform_bins, cs_initial_global_tree_*,
  cs_spanning_tree_u64, cs_exchange_buffer_layout) against the reference's known-answer vectors
  (test/unit/domain/domaindecomp.cpp:22-71, domain/buffer_description.hpp:98-125);
* the N > 1 path with world_size 2 over `gloo`: two processes build the replicated global tree exactly the way
  GlobalAssignment::assign does (domain/assignment.hpp:92-144) - local node counts from the oracle, the sum over ranks
  through torch.distributed (the role ncclAllReduce has on the GPUs), counts = max(local, sum), rebalance until every
  count fits the bucket - then cut the curve with cs_uniform_bins.  Global leaves and the assigned particle numbers
  of both ranks must equal what the UNMODIFIED reference computes with 2 ranks (oracle/_ref).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
LIB = os.path.join(ROOT, "cornerstone-octree_b200", "cstone_b200", "libcstone_b200.so")


def host_lib():
    lib = C.CDLL(LIB)
    lib.cs_initial_global_tree_u64.restype = C.c_long
    lib.cs_initial_global_tree_u32.restype = C.c_long
    lib.cs_spanning_tree_u64.restype = C.c_long
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def uniform_bins(lib, counts, num_bins):
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    bins = np.zeros(num_bins + 1, dtype=np.int32)
    bin_counts = np.zeros(num_bins, dtype=np.uint32)
    assert lib.cs_uniform_bins(_p(counts), C.c_size_t(counts.size), C.c_int(num_bins), _p(bins), _p(bin_counts)) == 0
    return bins, bin_counts


def initial_tree(lib, num_ranks, dtype=np.uint64):
    f = lib.cs_initial_global_tree_u64 if dtype == np.uint64 else lib.cs_initial_global_tree_u32
    n = f(C.c_int(num_ranks), None, C.c_long(0))
    leaves = np.zeros(n, dtype=dtype)
    assert f(C.c_int(num_ranks), _p(leaves), C.c_long(n)) == n
    return leaves


# ---------------------------------------------------------------------------------------------- known answers
def test_uniform_bins_reference_vectors():
    """test/unit/domain/domaindecomp.cpp:22-71"""
    lib = host_lib()
    umax = 0xFFFFFFFF
    bins, cnt = uniform_bins(lib, [umax - 10, 5, 5, umax - 11, 5, 6], 2)
    assert bins.tolist() == [0, 3, 6] and cnt.tolist() == [umax, umax]
    bins, _ = uniform_bins(lib, [5, 5, 5, 15, 1, 0], 2)
    assert bins.tolist() == [0, 3, 6]
    _, cnt = uniform_bins(lib, [15, 0, 1, 5, 5, 5], 2)
    assert (cnt.min(), cnt.max()) == (15, 16)
    _, cnt = uniform_bins(lib, [4, 3, 4, 3, 4, 3, 4, 3, 4, 3], 7)
    assert (cnt.min(), cnt.max()) == (3, 7)


def test_make_sfc_assignment_vector():
    """makeSfcAssignment (domaindecomp.cpp:73-88): counts {5,5,5,5} over leaves {0,10,20,30,40}, 2 ranks -> {0,20,40}"""
    lib = host_lib()
    bins, cnt = uniform_bins(lib, [5, 5, 5, 5], 2)
    leaves = np.array([0, 10, 20, 30, 40], dtype=np.uint64)
    assert leaves[bins].tolist() == [0, 20, 40]
    assert bins.tolist() == [0, 2, 4] and cnt.tolist() == [10, 10]


@pytest.mark.parametrize("num_ranks", [1, 2, 3, 8, 13])
@pytest.mark.parametrize("dtype", [np.uint64, np.uint32])
def test_initial_global_tree_is_a_cornerstone_array(num_ranks, dtype):
    """GlobalAssignment ctor (assignment.hpp:62-66): computeSpanningTree(initialDomainSplits(P, log8ceil(100 P)));
    invariants of tree/csarray.hpp:10-34 and the initialDomainSplit test (domaindecomp.cpp:134-141)"""
    lib = host_lib()
    leaves = initial_tree(lib, num_ranks, dtype)
    bits = 63 if dtype == np.uint64 else 30
    assert leaves[0] == 0 and int(leaves[-1]) == 1 << bits
    d = np.diff(leaves.astype(np.uint64))
    assert (d > 0).all()
    log2 = np.log2(d.astype(np.float64))
    assert np.array_equal(log2, np.round(log2)) and (log2.astype(np.int64) % 3 == 0).all()
    # every node starts at a multiple of its own size
    assert ((leaves[:-1].astype(np.uint64) % d) == 0).all()
    # the splits are taken at level log8ceil(100 P): at least one leaf per rank, at most the full grid of that level
    level = int(np.ceil(np.log(100 * num_ranks) / np.log(8)))
    assert num_ranks <= leaves.size - 1 <= 8 ** level


def test_spanning_tree_sizes():
    """computeSpanningTree (tree/csarray.hpp:483-510; sizes as in test/unit/tree/csarray.cpp:350-374): resolving key 1
    takes 7 siblings on each of the 21 levels plus the two cells of size 1; a key on a level-1 boundary one split"""
    lib = host_lib()
    out = np.zeros(1024, dtype=np.uint64)
    keys = np.array([0, 1, 1 << 63], dtype=np.uint64)
    n = lib.cs_spanning_tree_u64(_p(keys), C.c_long(3), _p(out), C.c_long(out.size))
    assert n == 7 * 21 + 2 and out[0] == 0 and out[1] == 1 and int(out[n - 1]) == 1 << 63
    keys = np.array([0, 1 << 60, 1 << 63], dtype=np.uint64)
    n = lib.cs_spanning_tree_u64(_p(keys), C.c_long(3), _p(out), C.c_long(out.size))
    assert out[:n].tolist() == [i << 60 for i in range(9)]


@pytest.mark.parametrize("start,end,size,present,assigned,want", [
    # incoming fit in the head gap: receive right before start (buffer_description.hpp:98-125)
    (10, 20, 30, 8, 12, (30, 6, 6, 20)),
    # head too small, tail fits: receive at end
    (2, 20, 30, 8, 12, (30, 20, 2, 24)),
    # neither fits: buffer grows to end + incoming, receive at end
    (1, 20, 21, 10, 16, (26, 20, 1, 26)),
    # nothing incoming
    (0, 16, 16, 16, 16, (16, 0, 0, 16)),
])
def test_exchange_buffer_layout(start, end, size, present, assigned, want):
    lib = host_lib()
    out = np.zeros(4, dtype=np.uint32)
    assert lib.cs_exchange_buffer_layout(C.c_uint32(start), C.c_uint32(end), C.c_uint32(size), C.c_uint32(present),
                                         C.c_uint32(assigned), _p(out)) == 0
    assert tuple(out.tolist()) == want


# ---------------------------------------------------------------------------------------------- gloo, world_size 2
def _rank_particles(rank, n_per):
    rng = np.random.default_rng(42 + rank)
    return tuple(rng.random(n_per) for _ in range(3))


def _gloo_worker(rank, world, port, n_per, bucket, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    import _libs

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = host_lib()
        orc = _libs.oracle()
        bnd = (0, 0, 0)
        x, y, z = _rank_particles(rank, n_per)
        # makeGlobalBox (sfc/box_mpi.hpp:51-105): open dimensions take the extent of all particles, one MIN reduction
        # over {min, -max}
        ext = torch.tensor([x.min(), -x.max(), y.min(), -y.max(), z.min(), -z.max()], dtype=torch.float64)
        dist.all_reduce(ext, op=dist.ReduceOp.MIN)
        lim = tuple(float(v) * (1 if i % 2 == 0 else -1) for i, v in enumerate(ext))
        keys = orc.sfc_keys("u64d", 0, x, y, z, lim, bnd)
        keys.sort()

        def global_counts(leaves):
            local = orc.compute_node_counts("u64", leaves, keys)
            total = torch.from_numpy(local.astype(np.int64))
            dist.all_reduce(total, op=dist.ReduceOp.SUM)          # ncclAllReduce on the GPUs
            summed = np.minimum(total.numpy(), 0xFFFFFFFF).astype(np.uint32)
            return np.maximum(local, summed)                     # update_mpi.hpp:86-97

        # GlobalAssignment ctor + assign (assignment.hpp:53-74, 92-144)
        leaves = initial_tree(lib, world)
        counts = np.full(leaves.size - 1, bucket - 1, dtype=np.uint32)

        def update(leaves, counts):
            ops, conv = orc.rebalance_decision("u64", leaves, counts, bucket)
            ops_scan = np.zeros(ops.size + 1, dtype=np.int32)
            ops_scan[:-1] = ops
            new_n = int(ops.sum())
            new_leaves = np.zeros(new_n + 1, dtype=np.uint64)
            f = orc._fn("rebalance_tree_u64", C.c_int)
            assert f(_p(leaves), C.c_int(ops.size), _p(ops_scan), _p(new_leaves)) == new_n
            return new_leaves, global_counts(new_leaves), conv

        leaves, counts, _ = update(leaves, counts)
        while True:  # first call: iterate until the largest count fits the bucket
            leaves, counts, _ = update(leaves, counts)
            if counts.max() <= bucket:
                break
        bins, bin_counts = uniform_bins(lib, counts, world)
        boundaries = leaves[bins]
        send = np.searchsorted(keys, boundaries, side="left")     # createSendRanges (domaindecomp.hpp:178-191)
        sent = torch.from_numpy(np.diff(send).astype(np.int64))
        everyone = [torch.zeros_like(sent) for _ in range(world)]
        dist.all_gather(everyone, sent)
        received = sum(int(t[rank]) for t in everyone)
        np.savez(os.path.join(result_dir, f"rank{rank}.npz"), leaves=leaves, counts=counts, bins=bins,
                 bin_counts=bin_counts, received=received, boundaries=boundaries)
    finally:
        dist.destroy_process_group()


def test_global_assignment_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp

    import _libs

    world, n_per, bucket = 2, 20000, 64
    port = 29000 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(world, port, n_per, bucket, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # replicated state is identical on both ranks and conserves the particles
    for k in ("leaves", "counts", "bins", "bin_counts", "boundaries"):
        assert np.array_equal(res[0][k], res[1][k]), k
    assert int(res[0]["counts"].astype(np.int64).sum()) == world * n_per
    assert [int(r["received"]) for r in res] == res[0]["bin_counts"].tolist()

    if _libs.ref() is None:
        pytest.skip("oracle/_ref not built: comparison with the unmodified reference skipped")
    xs, ys, zs = (np.concatenate(a) for a in zip(*[_rank_particles(r, n_per) for r in range(world)]))
    h = np.full(world * n_per, 0.01)
    want = _libs.ref_domain_run("u64d", world, bucket, 8, 0.5, (0, 1, 0, 1, 0, 1), (0, 0, 0), xs, ys, zs, h,
                                [0, n_per, 2 * n_per], num_syncs=1)
    assert np.array_equal(res[0]["leaves"], want[0]["global_leaves"])
    for r in range(world):
        assert int(res[r]["received"]) == want[r]["end"] - want[r]["start"], r
