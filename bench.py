#!/usr/bin/env python
"""Benchmark of the Cornerstone domain-sync hot path on B200 (BASELINE.json metric:
"Mparticles/s, Domain::sync + findNeighbors").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config uniform|plummer|morton]

One "step" = one pass of the hot path over one batch of synthetic particles.  Workloads (BASELINE.json `configs`):

  uniform (default)  N=1: configs[1], 64 Mi uniform particles, 64-bit Hilbert, double: first Domain::sync (cold trees,
                     unsorted input) + findNeighbors.  N>1: configs[3], ONE Domain over N ranks (64 Mi particles per
                     GPU drawn over the whole periodic box, so the first sync moves (N-1)/N of them): global-tree
                     ncclAllReduce, exchangeParticles, LET and exchangeHalos over NCCL / peer memory (weak scaling).
  plummer            configs[2], 64 Mi Plummer-sphere particles (deep tree), 64-bit Hilbert, double: first Domain::sync
                     + halo discovery over the first quarter of the leaves (test/performance/octree.cu:110-143).
  morton             configs[4], 16 Mi uniform particles, 32-bit Morton keys, float, ng ~ 300: standalone tree build +
                     findNeighbors.

Prints ONE JSON line (rank 0).  `value` is measured with inputs resident in HBM; `e2e` goes through the same C-ABI
calls but starts from pinned HOST buffers and ends with the results' D2H copies inside the timed region.
`checks` holds the parity evidence of the run: a bounded sample of the workload (N>1: a reduced-size Domain over the
SAME communicator, before the timed region) is run on the GPU and by the unmodified reference (oracle/_ref) and every
result array is compared through position-weighted 64-bit digests (equal digests <=> equal arrays), plus
size-independent invariants of the full-size state.
`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified reference headers) on
the host cores: the FULL workload at N=1, a reduced-size P=N run (ranks as threads) for N>1.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Mparticles/s, Domain::sync + findNeighbors"
UNIT = "Mparticles/s"
BUCKET = 64
NG0 = 100
NGMAX = 150
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
MASK64 = (1 << 64) - 1
DIGEST_C1, DIGEST_C2 = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F  # oracle/ref_api.cpp


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if k in d:
                    return float(d[k]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def h_for(n, ng):
    import numpy as np
    return 0.5 * float(np.cbrt(3.0 * ng / (4 * np.pi * n)))


# ------------------------------------------------------------------------------------------------ workloads
def workload(config, n, world):
    """the description both arms print as `config` (identical for --impl ours and --impl reference) and the
    parameters of the run"""
    if config == "uniform":
        bucket = BUCKET if world == 1 else max(BUCKET, (n * world) // (100 * world))
        w = dict(key="u64", real="d", kind=0, bucket=bucket, bucket_focus=BUCKET, ng=NG0, ngmax=NGMAX,
                 lim=(0, 1, 0, 1, 0, 1), bnd=(0, 0, 0) if world == 1 else (1, 1, 1))
        box = "open" if world == 1 else "periodic"
        text = (f"{n} uniform-random particles per GPU, 64-bit Hilbert, double, bucketSize={bucket}, "
                f"bucketSizeFocus={BUCKET}, {box} box: first Domain::sync (cold trees, unsorted input) + "
                f"findNeighbors(ng~{NG0}, ngmax={NGMAX})")
        par = ("single-rank Domain" if world == 1 else
               f"one Domain over {world} ranks (SFC ranges): ncclAllReduce of global node counts, exchangeParticles "
               f"through peer memory, LET treelets / exchangeHalos over ncclSend/Recv")
    elif config == "plummer":
        w = dict(key="u64", real="d", kind=0, bucket=BUCKET, bucket_focus=BUCKET, ng=0, ngmax=0,
                 lim=(-1, 1, -1, 1, -1, 1), bnd=(0, 0, 0))
        text = (f"{n} Plummer-sphere particles (test/coord_samples/plummer.hpp distribution), 64-bit Hilbert, double, "
                f"bucketSize={BUCKET}, open box: first Domain::sync (cold trees, unsorted input) + halo discovery "
                f"(search boxes + findHalos, first quarter of the leaves as own range), h=0.01")
        par = "single-rank Domain"
    elif config == "morton":
        w = dict(key="u32", real="f", kind=1, bucket=BUCKET, bucket_focus=BUCKET, ng=300, ngmax=384,
                 lim=(0, 1, 0, 1, 0, 1), bnd=(0, 0, 0))
        text = (f"{n} uniform-random particles, 32-bit Morton keys, float, bucketSize={BUCKET}: keys + sort + gather + "
                f"computeOctree + link + centres + findNeighbors(ng~300, ngmax=384)")
        par = "single GPU, standalone stage calls"
    else:
        raise SystemExit(f"unknown --config {config}")
    w["config"] = {"workload": text, "name": config, "particles_per_gpu": n, "parallelism": par,
                   "l2_policy": "inputs (>= 64 MB per array at the quoted size, 512 MB for the 64 Mi workloads) "
                                "exceed or rival the 126 MB L2; every step rewrites > 1 GB between uses"}
    return w


def make_particles(config, n, seed, n_for_h=None):
    """host (numpy) particles of a workload; rank r of a multi-rank run uses seed 42 + r"""
    import numpy as np
    rng = np.random.default_rng(seed)
    if config == "uniform":
        x, y, z = (rng.random(n) for _ in range(3))
        h = np.full(n, h_for(n_for_h or n, NG0))
    elif config == "plummer":
        # test/coord_samples/plummer.hpp:15-78 (radius cut at 100, scaled by 3 pi / 16, centred) with numpy's generator
        parts, need = [], n
        while need > 0:
            u = np.maximum(rng.random(int(need * 1.01) + 1024), 1e-300)
            r = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
            r = r[r < 100.0][:need]
            parts.append(r)
            need -= r.size
        r = np.concatenate(parts)
        zc = (1.0 - 2.0 * rng.random(n)) * r
        th = 2 * np.pi * rng.random(n)
        rho = np.sqrt(np.maximum(r * r - zc * zc, 0.0))
        conv = 3.0 * np.pi / 16.0
        x, y, z = rho * np.cos(th) * conv, rho * np.sin(th) * conv, zc * conv
        x, y, z = x - x.mean(), y - y.mean(), z - z.mean()
        h = np.full(n, 0.01)
    else:
        x, y, z = (np.minimum(rng.random(n, dtype=np.float32), np.float32(0.99999994)) for _ in range(3))
        h = np.full(n, h_for(n_for_h or n, 300), dtype=np.float32)
    return x, y, z, h


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([f.strip() for f in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) < 6:
                continue
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for nme, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ digests (GPU side)
def wsum(t):
    """sum_i bits(t[i]) * (i + 1) mod 2^64 of a device tensor (the digest of oracle/ref_api.cpp:weightedSum)"""
    import torch
    t = t.reshape(-1)
    if t.numel() == 0:
        return 0
    if t.element_size() == 8:
        v = t.view(torch.int64)
    elif t.element_size() == 4:
        v = t.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    else:
        v = t.to(torch.int64)
    acc, chunk = 0, 1 << 26
    for c0 in range(0, v.numel(), chunk):
        c1 = min(c0 + chunk, v.numel())
        idx = torch.arange(c0 + 1, c1 + 1, dtype=torch.int64, device=t.device)
        acc = (acc + int((v[c0:c1] * idx).sum())) & MASK64  # int64 arithmetic wraps: exactly mod 2^64
    return acc


def list_digest(nb, nc, ngmax, chunk_rows=1 << 19):
    """(sum of counts, digest of the neighbour lists) as oracle/ref_api.cpp:neighborsChunked computes them"""
    import torch

    def s64(c):
        return c - (1 << 64) if c >= (1 << 63) else c

    n = nc.numel()
    nbv = nb[: n * ngmax].view(n, ngmax)
    k = (torch.arange(1, ngmax + 1, dtype=torch.int64, device=nb.device) * s64(DIGEST_C2))[None, :]
    kk = torch.arange(ngmax, dtype=torch.int64, device=nb.device)[None, :]
    acc, total = 0, 0
    for c0 in range(0, n, chunk_rows):
        c1 = min(c0 + chunk_rows, n)
        cnt = nc[c0:c1].to(torch.int64)
        total += int(cnt.sum())
        rows = (torch.arange(c0 + 1, c1 + 1, dtype=torch.int64, device=nb.device) * s64(DIGEST_C1))[:, None]
        val = (nbv[c0:c1].to(torch.int64) + 1) * (rows + k)
        val = torch.where(kk < torch.clamp(cnt, max=ngmax)[:, None], val, torch.zeros_like(val))
        acc = (acc + int(val.sum())) & MASK64
    return total, acc


def domain_digest(dom):
    """digests of everything Domain::sync leaves behind (names as in tests/_libs.py:DIGEST_SLOTS)"""
    d = {"keys": wsum(dom.field("keys")), "x": wsum(dom.field("x")), "y": wsum(dom.field("y")),
         "z": wsum(dom.field("z")), "h": wsum(dom.field("h")), "leaves": wsum(dom.field("focus_leaves")),
         "num_leaves": dom.num_focus_leaves, "layout": wsum(dom.field("layout")), "start": dom.start_index,
         "end": dom.end_index, "size": dom.n_particles_with_halos, "num_nodes": dom.num_focus_nodes,
         "prefixes": wsum(dom.field("prefixes")), "child_offsets": wsum(dom.field("child_offsets")),
         "centers": wsum(dom.field("geo_centers")), "sizes": wsum(dom.field("geo_sizes")),
         "leaf_counts": wsum(dom.field("focus_leaf_counts"))}
    return d


def compare_digests(got, want, names=None):
    names = names or [k for k in got if k in want]
    bad = [k for k in names if int(got[k]) != int(want[k])]
    return {"arrays_compared": names, "mismatches": bad, "identical": not bad}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def _ref_env(threads=None, ranks=1):
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the reference arm uses every host core
    os.environ["OMP_NUM_THREADS"] = str(threads or os.cpu_count())
    os.environ.setdefault("OMP_PROC_BIND", "spread" if ranks == 1 else "false")
    if ranks > 1:
        # ranks are threads of this process, each with its own OpenMP team: spinning teams starve each other
        # (measured: Domain::sync 15x slower with the default active wait policy)
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    import _libs
    if _libs.ref_lib() is None:
        raise RuntimeError("oracle/_ref/libcstone_ref.so is missing (it is built where /root/reference exists and "
                           "travels with the snapshot)")
    import ctypes as C
    try:
        # the OpenMP runtime may have been initialised (by torch) before the variable was set
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(threads or os.cpu_count()))
    except Exception:
        pass
    return _libs


def ref_step(config, w, x, y, z, h, offsets, threads):
    """one step of the workload with the unmodified reference; returns (seconds, per-phase seconds, digests of rank 0,
    all ranks' results)"""
    import _libs
    P = len(offsets) - 1
    if config == "morton":
        out = _libs.ref_bench_tree_neighbors(w["key"] + w["real"], w["kind"], x, y, z, h, w["bucket"], w["lim"], w["bnd"],
                                             w["ngmax"])
        return out["t_build"] + out["t_neighbors"], {"build": out["t_build"], "neighbors": out["t_neighbors"]}, [out]
    out = _libs.ref_bench_run(w["key"] + w["real"], P, w["bucket"], w["bucket_focus"], 0.5, w["lim"], w["bnd"], x, y, z,
                              h, offsets, num_syncs=1, ngmax=w["ngmax"], threads=threads,
                              halo_quarter=(config == "plummer"))
    t_sync = max(o["t_sync"][0] for o in out)
    t_nb = max(o["t_neighbors"] for o in out)
    t_halo = max(o["t_halos"] for o in out)
    return t_sync + t_nb + t_halo, {"sync": t_sync, "neighbors": t_nb, "halos": t_halo}, out


def run_reference(args, n, world, budget_s, steps, warmup):
    """the reference's own CPU path on `n` particles per rank (world ranks as threads); returns the cpu_baseline object
    plus the digests of the last step"""
    import numpy as np
    cores = os.cpu_count()
    threads = max(1, cores // world)
    _libs = _ref_env(threads, world)
    w = workload(args.config, n, world)  # bucketSize = N_total / (100 P) follows the size that actually runs
    xs, ys, zs, hs, offsets = [], [], [], [], [0]
    for r in range(world):
        x, y, z, h = make_particles(args.config, n, 42 + r, n_for_h=n * world)
        xs.append(x), ys.append(y), zs.append(z), hs.append(h)
        offsets.append(offsets[-1] + n)
    x, y, z, h = (np.concatenate(a) if world > 1 else a[0] for a in (xs, ys, zs, hs))
    del xs, ys, zs, hs
    values, phases, out = [], None, None
    t_begin = time.perf_counter()
    done = 0
    for it in range(warmup + steps):
        dt, phases, out = ref_step(args.config, w, x, y, z, h, offsets, threads)
        if it >= warmup:
            values.append(world * n / dt / 1e6)
            done += 1
        # bounded: stop early when the next step would overrun the time budget (at least one timed step)
        elapsed = time.perf_counter() - t_begin
        if done >= 1 and elapsed + dt * 1.2 > budget_s:
            break
    v = statistics.median(values)
    full = n == args.n
    sample = (f"{'the full workload: ' if full else 'bounded sample: '}{n} particles per rank x {world} rank(s) "
              f"(ranks = threads of one process over the in-process MPI stand-in), {threads} OpenMP thread(s) per rank; "
              f"phases of the last step [s]: " + ", ".join(f"{k} {v_:.3f}" for k, v_ in phases.items()))
    cb = {"value": v, "unit": UNIT, "cores": min(cores, threads * world), "kind": "reference", "sample": sample,
          "particles_per_rank": n, "ranks": world, "steps_timed": done}
    return cb, out


# ------------------------------------------------------------------------------------------------ our arm: helpers
def stage_rooflines(n, dev, x, y, z, h, key_dtype, kind, reps=3):
    """the HBM-bound stage kernels timed alone with CUDA events on the launching stream (same kernels the Domain
    launches), against their algorithmic bytes (SURVEY.md 8d / DESIGN.md)"""
    import ctypes as C

    import torch

    from cstone_b200 import capi

    lim, bnd = (0, 1, 0, 1, 0, 1), (0, 0, 0)
    if float(x.min()) < 0:
        m = float(max(x.abs().max(), y.abs().max(), z.abs().max())) * 1.0001
        lim = (-m, m, -m, m, -m, m)
    res = {}

    def timeit(name, fn, setup=None):
        ms = []
        for _ in range(reps + 1):
            if setup:
                setup()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        res[name] = statistics.median(ms[1:])

    keys = torch.zeros(n, dtype=key_dtype, device=dev)
    timeit("keys", lambda: capi.compute_sfc_keys(x, y, z, keys, lim, bnd, kind=kind))
    unsorted = keys.clone()
    order = capi.sequence(0, n, dev)
    kt = "u64" if key_dtype == torch.uint64 else "u32"
    tmp_bytes = getattr(capi.lib(), "cs_sort_by_key_temp_bytes_" + kt)(C.c_size_t(n))
    key_buf, val_buf = torch.empty_like(keys), torch.empty_like(order)
    tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)

    def sort_setup():
        keys.copy_(unsorted)
        order.copy_(capi.sequence(0, n, dev))

    def sort_call():
        capi._check(getattr(capi.lib(), "cs_sort_by_key_" + kt)(capi._ptr(keys), capi._ptr(order), C.c_size_t(n),
                                                                 capi._ptr(key_buf), capi._ptr(val_buf), capi._ptr(tmp),
                                                                 C.c_size_t(tmp_bytes), capi._stream()), "sort")

    timeit("sort", sort_call, sort_setup)
    outs = [torch.empty_like(x) for _ in range(4)]
    src_a = (C.c_void_p * 4)(*[t.data_ptr() for t in (x, y, z, h)])
    dst_a = (C.c_void_p * 4)(*[t.data_ptr() for t in outs])
    timeit("gather", lambda: capi._check(capi.lib().cs_gather_arrays4(capi._ptr(order), C.c_size_t(n), C.c_size_t(n),
                                                                      src_a, dst_a, C.c_int(x.element_size()),
                                                                      capi._stream()), "gather_arrays4"))
    leaves, counts = capi.compute_octree(keys, BUCKET)
    timeit("counts", lambda: capi.compute_node_counts(leaves, keys))
    res["num_leaves"] = leaves.numel() - 1
    del keys, unsorted, order, key_buf, val_buf, tmp, outs
    return res


def stage_table(st, n, kbytes, rbytes, peak):
    passes = kbytes  # 8-bit digits over all key bits
    alg = {"keys": (3.0 * rbytes + 2 * kbytes) * n, "sort": (kbytes + passes * 2.0 * (kbytes + 4)) * n,
           "gather": (4.0 + 8 * rbytes) * n, "counts": float(kbytes) * n + (kbytes + 4.0) * st["num_leaves"]}
    stages = {}
    for name, nbytes in alg.items():
        gbs = nbytes / (st[name] * 1e-3) / 1e9
        stages[name] = {"ms": round(st[name], 4), "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4),
                        "algorithmic_bytes": nbytes}
    return stages, alg


def sort_roofline(stages, n, kbytes, peak, peak_kind):
    per_particle = kbytes + kbytes * 2 * (kbytes + 4)
    return {"bound": "hbm", "kernel": f"onesweepKernel<u{8 * kbytes},values> x{kbytes} (+ radixHistogramKernel)",
            "achieved": stages["sort"]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": stages["sort"]["frac"],
            # not measured by this run (ncu cannot run inside the bench): the ratio of DRAM traffic to algorithmic bytes
            # comes from the committed ncu --set full capture of the same launches at 64 Mi u64 keys
            # (profiles/r2_final_full_summary.txt: 8 x (812.2 MB read + 811.1 MB written) + histogram 536.9 MB read
            # + 4.8 MB written = 13.528 GB against 13.422 GB algorithmic) and is scaled with n
            "traffic": 1.008 * per_particle * n,
            "traffic_source": "ncu --set full of the sort launches (profiles/r2_final_full_summary.txt): dram read + "
                              "write = 1.008 x algorithmic bytes at 64 Mi u64 keys; scaled with n, not re-measured here",
            "peak_source": peak_kind, "algorithmic_bytes_per_particle": per_particle,
            "launch_group_ms": stages["sort"]["ms"],
            "note": "dominant HBM-bound kernel group of the step; traversal kernels are under `stages`"}


def pinned(a):
    import torch
    return torch.from_numpy(a).pin_memory()


# ------------------------------------------------------------------------------------------------ our arm: uniform
def run_uniform(args):
    import numpy as np
    import torch

    from cstone_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    w = workload("uniform", n, world)
    lim, bnd, bucket = w["lim"], w["bnd"], w["bucket"]
    # every rank draws its particles over the WHOLE box (BASELINE configs[3]): the first sync moves (P-1)/P of them
    hx, hy, hz, hh = (pinned(a) for a in make_particles("uniform", n, 42 + rank, n_for_h=n * world))
    x, y, z, h = (t.to(dev) for t in (hx, hy, hz, hh))

    comm = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Comm.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = capi.Comm.nccl(rank, world, bytes(uid.cpu().numpy().tobytes()))

    checks = {}

    # ---- N>1: bit-exact parity of the REAL transport (NCCL + peer-memory exchange) before anything is timed: a
    # reduced-size Domain over the same communicator against the unmodified reference run with P = N thread-ranks
    if world > 1:
        checks["transport_parity"] = multirank_parity(args, comm, rank, world, dev, dist)

    dom = capi.Domain(rank, world, bucket, BUCKET, 0.5, lim, bnd, key="u64", real="d", device=str(dev), comm=comm)
    # the assigned share is n +- the granularity of the global leaves; halos add a shell around it
    cap = n if world == 1 else int(n * 1.05)
    cap_halo = n if world == 1 else int(n * 1.6)
    nb = torch.empty(cap * NGMAX, dtype=torch.uint32, device=dev)
    nc = torch.empty(cap, dtype=torch.uint32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    sync_ms, nb_ms = [], []

    def step(record=False):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        dom.reset()
        dom.sync(x, y, z, h)
        if dom.end_index - dom.start_index > cap or dom.n_particles_with_halos > cap_halo:
            raise SystemExit(f"rank {rank}: assigned {dom.end_index - dom.start_index} / with halos "
                             f"{dom.n_particles_with_halos} exceed the bench buffers ({cap}, {cap_halo})")
        e1.record()
        dom.find_neighbors(NGMAX, nb, nc)
        e2.record()
        if record:
            sync_ms.append((e0, e1))
            nb_ms.append((e1, e2))

    # ---- device-resident timing: cold Domain::sync (fresh trees) + findNeighbors on unsorted input
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = capi.kernel_launch_count()
    sent0 = comm.bytes_sent if comm is not None else 0
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = capi.kernel_launch_count() - launches0
    bytes_per_step = ((comm.bytes_sent if comm is not None else 0) - sent0) / args.steps
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    cold_sync = statistics.median(a.elapsed_time(b) for a, b in sync_ms)
    nb_time = statistics.median(a.elapsed_time(b) for a, b in nb_ms)
    num_leaves = dom.num_focus_leaves

    # ---- invariants of the synchronised state (outside the timed region; size independent): every particle arrived
    # exactly once (wrapping integer sums of the coordinate bit patterns are order independent and exact), keys are
    # sorted on every rank and the ranks' key ranges are ordered along the curve
    def bits_sum(t):
        return t.view(torch.int64).sum(dtype=torch.int64)

    s_, e_ = dom.start_index, dom.end_index
    keys_assigned = dom.field("keys")[s_:e_].view(torch.int64)
    chk = torch.stack([bits_sum(x) + bits_sum(y) + bits_sum(z),
                       bits_sum(dom.field("x")[s_:e_]) + bits_sum(dom.field("y")[s_:e_]) + bits_sum(dom.field("z")[s_:e_]),
                       torch.tensor(e_ - s_, dtype=torch.int64, device=dev)])
    ends = torch.stack([keys_assigned[0], keys_assigned[-1]]) if e_ > s_ else torch.zeros(2, dtype=torch.int64, device=dev)
    sorted_ok = torch.tensor([int(bool((keys_assigned[1:] >= keys_assigned[:-1]).all()))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(chk)  # sums wrap consistently on every rank
        all_ends = [torch.zeros_like(ends) for _ in range(world)]
        dist.all_gather(all_ends, ends)
        dist.all_reduce(sorted_ok, op=dist.ReduceOp.MIN)
        flat = torch.cat(all_ends)
        ranges_ordered = bool((flat[1:] >= flat[:-1]).all())
    else:
        ranges_ordered = True
    checks.update({"every_particle_arrived_once": bool(chk[0] == chk[1]) and int(chk[2]) == world * n,
                   "keys_sorted_on_every_rank": bool(sorted_ok.item()), "rank_key_ranges_ordered": ranges_ordered})
    del keys_assigned
    # digests of the full-size result (N=1: comparable with the `digest` of `--impl reference`, same particles)
    digest = None
    if world == 1:
        digest = domain_digest(dom)
        digest["nc_sum"], digest["lists"] = list_digest(nb, nc[: e_ - s_], NGMAX)
        digest = {k: (f"{v:#018x}" if v > 1 << 40 else v) for k, v in digest.items()}

    # ---- steady state: re-sync the (already SFC-ordered) domain arrays in place, one tree update per call
    steady = []
    for _ in range(4):
        a, b = ev(), ev()
        a.record()
        dom.sync()
        b.record()
        torch.cuda.synchronize()
        steady.append(a.elapsed_time(b))
    steady_sync = statistics.median(steady[1:])

    # ---- end to end: pinned host inputs -> C ABI (H2D inside) -> results back in pinned host memory
    out_host = [torch.empty(cap_halo, dtype=torch.float64).pin_memory() for _ in range(4)]
    keys_host = torch.empty(cap_halo, dtype=torch.uint64).pin_memory()
    nc_host = torch.empty(cap, dtype=torch.uint32).pin_memory()
    side = torch.cuda.Stream(device=dev)

    def e2e_step():
        main = torch.cuda.current_stream()
        dom.reset()
        dom.sync(hx, hy, hz, hh)
        # the synchronised arrays travel back on a second stream while the neighbour search runs
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dom.download(*out_host, keys_host)
        dom.find_neighbors(NGMAX, nb, nc)
        nc_host.copy_(nc, non_blocking=True)
        main.wait_stream(side)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    e2e_value = world * n / (e2e_ms * 1e-3) / 1e6
    h2d = 4 * 8 * n
    d2h = (4 * 8 + 8) * dom.n_particles_with_halos + 4 * cap
    mean_nc = float(nc_host[: 1 << 20].to(torch.int64).sum()) / float(min(n, 1 << 20))
    exchange = None
    if world > 1:
        stats = torch.tensor([bytes_per_step, dom.end_index - dom.start_index, dom.n_particles_with_halos,
                              dom.num_focus_leaves, dom.num_global_leaves], dtype=torch.float64, device=dev)
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
        exchange = {"bytes_sent_per_gpu_per_step": [int(t[0].item()) for t in allstats],
                    "assigned_per_gpu": [int(t[1].item()) for t in allstats],
                    "with_halos_per_gpu": [int(t[2].item()) for t in allstats],
                    "focus_leaves_per_gpu": [int(t[3].item()) for t in allstats],
                    "global_leaves": int(allstats[0][4].item())}
    dom.close()
    del nb, nc, out_host, keys_host
    torch.cuda.empty_cache()

    if rank != 0:
        return None

    # ---- per-stage rooflines
    peak, peak_kind = hbm_peak()
    st = stage_rooflines(n, dev, x, y, z, h, torch.uint64, 0)
    stages, _ = stage_table(st, n, 8, 8, peak)
    nb_bytes = (4.0 * min(mean_nc, NGMAX) + 4 + 32) * n
    stages["neighbors"] = {"ms": round(nb_time, 3), "achieved_gbs": round(nb_bytes / (nb_time * 1e-3) / 1e9, 1),
                           "frac": round(nb_bytes / (nb_time * 1e-3) / 1e9 / peak, 4),
                           "note": "traversal + distance tests, instruction-issue bound (profiles/); bytes = list "
                                   "output + particle reads"}
    stages["domain_sync_cold"] = {"ms": round(cold_sync, 3)}
    stages["domain_sync_steady"] = {"ms": round(steady_sync, 3)}
    del x, y, z, h
    torch.cuda.empty_cache()

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / f64 coordinates", "data": "synthetic", "config": w["config"],
        "results": {"focus_leaves": num_leaves, "mean_neighbors": round(mean_nc, 2), "exchange": exchange,
                    "digest": digest},
        "checks": checks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(e2e_ms, 3),
                "note": "neighbour lists stay in HBM for the device-side consumer; keys, x,y,z,h (copied on a second "
                        "stream during the neighbour search) and counts return"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": sort_roofline(stages, n, 8, peak, peak_kind),
        "stages": stages,
    }
    return line


def multirank_parity(args, comm, rank, world, dev, dist):
    """reduced-size Domain (sync x2 + findNeighbors + client-field halo exchange) over the bench's own communicator,
    compared rank by rank with the unmodified reference run with P = world thread-ranks (rank 0 computes it)"""
    import numpy as np
    import torch

    from cstone_b200 import capi

    n_s, ngmax = args.parity_n, 64
    w = workload("uniform", n_s, world)
    names = ["keys", "x", "y", "z", "h", "leaves", "num_leaves", "layout", "start", "end", "size", "num_nodes",
             "prefixes", "child_offsets", "centers", "sizes", "leaf_counts", "nc_sum", "lists"]
    px, py, pz, ph = make_particles("uniform", n_s, 1042 + rank, n_for_h=n_s * world)
    dom = capi.Domain(rank, world, w["bucket"], w["bucket_focus"], 0.5, w["lim"], w["bnd"], key="u64", real="d",
                      device=str(dev), comm=comm)
    dom.sync(*(torch.from_numpy(a).to(dev) for a in (px, py, pz, ph)))
    nbs, ncs = dom.find_neighbors(ngmax)
    got = domain_digest(dom)
    got["nc_sum"], got["lists"] = list_digest(nbs.reshape(-1), ncs, ngmax)
    # client-field halo exchange through the recorded pattern
    rho = dom.field("x") * 2 + dom.field("y")
    f = torch.full_like(rho, -7)
    f[dom.start_index:dom.end_index] = rho[dom.start_index:dom.end_index]
    dom.exchange_halos(f)
    halo_ok = bool(torch.equal(f, rho))
    dom.close()
    del nbs, ncs

    want = torch.zeros((world, len(names)), dtype=torch.int64, device=dev)
    ref_ok = torch.ones(1, dtype=torch.int64, device=dev)
    if rank == 0:
        try:
            # a fresh process: this one's OpenMP runtime was initialised by torch under the launcher's settings
            env = {k: v for k, v in os.environ.items() if not k.startswith(("OMP_", "RANK", "WORLD_SIZE", "LOCAL_RANK"))}
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "parity-worker", "--gpus",
                                str(world), "--parity-n", str(n_s)], env=env, capture_output=True, text=True,
                               timeout=900)
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-500:])
            out = json.loads(r.stdout.strip().splitlines()[-1])
            for r_ in range(world):
                for k, nme in enumerate(names):
                    v = int(out[r_][nme])
                    want[r_, k] = v - (1 << 64) if v >= (1 << 63) else v
        except Exception as e:  # the checker being unavailable must not hide the GPU numbers
            ref_ok[0] = 0
            print(f"transport parity: reference unavailable: {e}", file=sys.stderr)
    dist.broadcast(want, 0)
    dist.broadcast(ref_ok, 0)
    bad = [nme for k, nme in enumerate(names) if (int(want[rank, k]) & MASK64) != (int(got[nme]) & MASK64)]
    flag = torch.tensor([0 if bad else 1, 1 if halo_ok else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not ref_ok.item():
        return {"bit_identical_vs_reference": None, "note": "oracle/_ref unavailable on this host"}
    if bad:
        print(f"rank {rank}: transport parity mismatch in {bad}", file=sys.stderr)
    return {"bit_identical_vs_reference": bool(flag[0].item()), "halo_exchange_of_client_field_ok": bool(flag[1].item()),
            "particles_per_rank": n_s, "ranks": world, "arrays_compared": names,
            "transport": "NCCL collectives + CUDA-IPC peer-memory exchangeParticles (the bench's communicator)",
            "reference": f"oracle/_ref Domain<uint64_t,double> with {world} thread-ranks, same particles"}


# ------------------------------------------------------------------------------------------------ our arm: plummer
def halo_discovery_quarter(capi, torch, dom, bnd):
    """search boxes + findHalos on the synchronised focus tree, own range = first quarter of the leaves"""
    nl = dom.num_focus_leaves
    tree = capi.Octree(dom.field("focus_leaves").clone())
    cen, siz = dom.field("geo_centers"), dom.field("geo_sizes")
    layout = dom.field("layout")
    first, last = 0, nl // 4
    sx, sy, sz, sh = (dom.field(k) for k in ("x", "y", "z", "h"))
    init = cen[tree.leaf_to_internal[tree.num_internal:].long()].contiguous()

    def halos():
        sc, ss = capi.compute_bounding_boxes(sx, sy, sz, sh, layout, first, last, 2.0, init)
        return capi.find_halos(tree, cen, siz, sc, ss, dom.box, bnd, first, last)

    return tree, halos, first, last


def run_plummer(args):
    import torch

    from cstone_b200 import capi

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n = args.n
    w = workload("plummer", n, 1)
    lim, bnd = w["lim"], w["bnd"]
    hx, hy, hz, hh = (pinned(a) for a in make_particles("plummer", n, 42))
    x, y, z, h = (t.to(dev) for t in (hx, hy, hz, hh))
    dom = capi.Domain(0, 1, BUCKET, BUCKET, 0.5, lim, bnd, key="u64", real="d", device=str(dev))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    sync_ms, halo_ms = [], []
    state = {}

    def step(record=False):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        dom.reset()
        dom.sync(x, y, z, h)
        e1.record()
        tree, halos, first, last = halo_discovery_quarter(capi, torch, dom, bnd)
        state["flags"] = halos()
        state["tree"], state["range"] = tree, (first, last)
        e2.record()
        if record:
            sync_ms.append((e0, e1))
            halo_ms.append((e1, e2))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = capi.kernel_launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = capi.kernel_launch_count() - launches0
    ms_per_step = e0.elapsed_time(e1) / args.steps
    value = n / (ms_per_step * 1e-3) / 1e6
    cold_sync = statistics.median(a.elapsed_time(b) for a, b in sync_ms)
    halo_time = statistics.median(a.elapsed_time(b) for a, b in halo_ms)

    keys = dom.field("keys").view(torch.int64)
    counts = dom.field("focus_leaf_counts").to(torch.int64)
    leaves = dom.field("focus_leaves").view(torch.int64)
    level = (63 - torch.log2((leaves[1:] - leaves[:-1]).to(torch.float64))) / 3
    flags, tree, (first, last) = state["flags"], state["tree"], state["range"]
    own_nodes = tree.leaf_to_internal[tree.num_internal:][first:last].long()

    def bits_sum(t):
        return int(t.view(torch.int64).sum(dtype=torch.int64))

    checks = {"keys_sorted": bool((keys[1:] >= keys[:-1]).all()), "counts_sum_equals_n": int(counts.sum()) == n,
              "leaf_counts_within_bucket_or_max_depth": bool(((counts <= BUCKET) | (level >= 21)).all()),
              "every_particle_arrived_once": all(bits_sum(a) == bits_sum(dom.field(k))
                                                 for a, k in ((x, "x"), (y, "y"), (z, "z"))),
              "no_halo_flag_inside_own_range": int(flags[own_nodes].sum()) == 0,
              "some_halos_found": int(flags.sum()) > 0}
    results = {"focus_leaves": dom.num_focus_leaves, "max_leaf_level": int(level.max()),
               "particles_per_leaf": round(n / dom.num_focus_leaves, 2), "halo_nodes": int(flags.sum())}
    digest = domain_digest(dom)
    digest["halo_flags"], digest["halo_count"] = wsum(flags), int(flags.sum())
    results["digest"] = {k: (f"{v:#018x}" if v > 1 << 40 else v) for k, v in digest.items()}

    steady = []
    for _ in range(4):
        a, b = ev(), ev()
        a.record()
        dom.sync()
        b.record()
        torch.cuda.synchronize()
        steady.append(a.elapsed_time(b))

    # ---- end to end: pinned host inputs, synchronised arrays + halo flags back in pinned host memory
    out_host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
    keys_host = torch.empty(n, dtype=torch.uint64).pin_memory()
    flags_host = torch.empty(dom.num_focus_nodes + 1024, dtype=torch.uint8).pin_memory()

    def e2e_step():
        dom.reset()
        dom.sync(hx, hy, hz, hh)
        dom.download(*out_host, keys_host)
        _, halos, _, _ = halo_discovery_quarter(capi, torch, dom, bnd)
        fl = halos()
        flags_host[: fl.numel()].copy_(fl, non_blocking=True)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    dom.close()

    peak, peak_kind = hbm_peak()
    st = stage_rooflines(n, dev, x, y, z, h, torch.uint64, 0)
    stages, _ = stage_table(st, n, 8, 8, peak)
    stages["domain_sync_cold"] = {"ms": round(cold_sync, 3)}
    stages["domain_sync_steady"] = {"ms": round(statistics.median(steady[1:]), 3)}
    stages["halo_discovery_quarter_range"] = {"ms": round(halo_time, 3),
                                              "note": "computeBoundingBoxes + findHalos, traversal (latency bound)"}
    return {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 keys / f64 coordinates", "data": "synthetic", "config": w["config"],
        "results": results, "checks": checks,
        "e2e": {"value": round(n / (e2e_ms * 1e-3) / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": 4 * 8 * n,
                "d2h_bytes_per_step": 5 * 8 * n + dom_nodes_bytes(results), "ms_per_step": round(e2e_ms, 3),
                "note": "x, y, z, h, keys and the halo flags return"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": sort_roofline(stages, n, 8, peak, peak_kind),
        "stages": stages,
    }


def dom_nodes_bytes(results):
    nl = results["focus_leaves"]
    return nl + (nl - 1) // 7


# ------------------------------------------------------------------------------------------------ our arm: morton
def run_morton(args):
    import torch

    from cstone_b200 import capi

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n = args.n
    w = workload("morton", n, 1)
    lim, bnd, ngmax = w["lim"], w["bnd"], w["ngmax"]
    hx, hy, hz, hh = (pinned(a) for a in make_particles("morton", n, 42))
    x, y, z, h = (t.to(dev) for t in (hx, hy, hz, hh))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    nb = torch.empty(n * ngmax, dtype=torch.uint32, device=dev)
    nc = torch.empty(n, dtype=torch.uint32, device=dev)
    keys = torch.zeros(n, dtype=torch.uint32, device=dev)
    s = {}

    def build(xi, yi, zi, hi_):
        keys.zero_()
        capi.compute_sfc_keys(xi, yi, zi, keys, lim, bnd, kind=1)
        order = capi.sequence(0, n, dev)
        capi.sort_by_key(keys, order)
        sx, sy, sz, sh = capi.gather_arrays4(order, [xi, yi, zi, hi_])
        leaves, counts = capi.compute_octree(keys, BUCKET)
        tree = capi.Octree(leaves)
        cen, siz = capi.compute_geo_centers(tree.prefixes, torch.float32, lim, bnd, kind=1)
        layout = capi.exclusive_scan(torch.cat([counts, torch.zeros(1, dtype=torch.uint32, device=dev)]))
        s.update(sx=sx, sy=sy, sz=sz, sh=sh, tree=tree, cen=cen, siz=siz, layout=layout, counts=counts, leaves=leaves)

    def search():
        capi.find_neighbors(s["sx"], s["sy"], s["sz"], s["sh"], 0, n, lim, bnd, s["tree"], s["layout"], s["cen"],
                            s["siz"], ngmax, nb, nc)

    build_ms, nb_ms = [], []

    def step(record=False):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        build(x, y, z, h)
        e1.record()
        search()
        e2.record()
        if record:
            build_ms.append((e0, e1))
            nb_ms.append((e1, e2))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = capi.kernel_launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = capi.kernel_launch_count() - launches0
    ms_per_step = e0.elapsed_time(e1) / args.steps
    build_time = statistics.median(a.elapsed_time(b) for a, b in build_ms)
    nb_time = statistics.median(a.elapsed_time(b) for a, b in nb_ms)

    ksort = keys.view(torch.int32)
    mean_nc = float(nc.to(torch.float64).mean())
    checks = {"keys_sorted": bool((ksort[1:] >= ksort[:-1]).all()),
              "counts_sum_equals_n": int(s["counts"].to(torch.int64).sum()) == n,
              "lists_ascending_without_self": lists_sorted_no_self(torch, nb, nc, ngmax, 1 << 20)}
    digest = {"keys": wsum(keys), "x": wsum(s["sx"]), "y": wsum(s["sy"]), "z": wsum(s["sz"]), "h": wsum(s["sh"]),
              "leaves": wsum(s["leaves"]), "num_leaves": s["tree"].num_leaves, "layout": wsum(s["layout"]),
              "num_nodes": s["tree"].num_nodes, "prefixes": wsum(s["tree"].prefixes),
              "child_offsets": wsum(s["tree"].child_offsets[: s["tree"].num_nodes]), "centers": wsum(s["cen"]),
              "sizes": wsum(s["siz"]), "leaf_counts": wsum(s["counts"])}
    digest["nc_sum"], digest["lists"] = list_digest(nb, nc, ngmax)
    results = {"leaves": s["tree"].num_leaves, "mean_neighbors": round(mean_nc, 2),
               "digest": {k: (f"{v:#018x}" if v > 1 << 40 else v) for k, v in digest.items()}}

    # ---- end to end: pinned host inputs; sorted coordinates, keys and neighbour counts back in pinned host memory
    out_host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(4)]
    keys_host = torch.empty(n, dtype=torch.uint32).pin_memory()
    nc_host = torch.empty(n, dtype=torch.uint32).pin_memory()
    dx, dy, dz, dh = (torch.empty_like(t) for t in (x, y, z, h))

    def e2e_step():
        for d_, h_ in zip((dx, dy, dz, dh), (hx, hy, hz, hh)):
            d_.copy_(h_, non_blocking=True)
        build(dx, dy, dz, dh)
        search()
        for o_, k_ in zip(out_host, ("sx", "sy", "sz", "sh")):
            o_.copy_(s[k_], non_blocking=True)
        keys_host.copy_(keys, non_blocking=True)
        nc_host.copy_(nc, non_blocking=True)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps

    peak, peak_kind = hbm_peak()
    st = stage_rooflines(n, dev, x, y, z, h, torch.uint32, 1)
    stages, _ = stage_table(st, n, 4, 4, peak)
    nb_bytes = (4.0 * min(mean_nc, ngmax) + 4 + 16) * n
    stages["build"] = {"ms": round(build_time, 3)}
    stages["neighbors"] = {"ms": round(nb_time, 3), "achieved_gbs": round(nb_bytes / (nb_time * 1e-3) / 1e9, 1),
                           "frac": round(nb_bytes / (nb_time * 1e-3) / 1e9 / peak, 4),
                           "note": "traversal + distance tests, instruction-issue bound; bytes = list output + "
                                   "particle reads"}
    return {
        "metric": METRIC, "value": round(n / (ms_per_step * 1e-3) / 1e6, 2), "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 keys / f32 coordinates", "data": "synthetic",
        "config": w["config"], "results": results, "checks": checks,
        "e2e": {"value": round(n / (e2e_ms * 1e-3) / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": 4 * 4 * n,
                "d2h_bytes_per_step": 6 * 4 * n, "ms_per_step": round(e2e_ms, 3),
                "note": "neighbour lists stay in HBM for the device-side consumer; sorted x,y,z,h, keys and counts "
                        "return"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": sort_roofline(stages, n, 4, peak, peak_kind),
        "stages": stages,
    }


def lists_sorted_no_self(torch, nb, nc, ngmax, rows):
    rows = min(rows, nc.numel())
    v = nb[: rows * ngmax].view(rows, ngmax).to(torch.int64)
    m = torch.arange(ngmax, device=nb.device)[None, :] < torch.clamp(nc[:rows].to(torch.int64), max=ngmax)[:, None]
    asc = bool(((v[:, 1:] > v[:, :-1]) | ~m[:, 1:]).all())
    no_self = not bool(((v == torch.arange(rows, device=nb.device)[:, None]) & m).any())
    return asc and no_self


# ------------------------------------------------------------------------------------------------ sample parity (N=1)
def sample_parity(args):
    """the GPU path and the unmodified reference on the same bounded sample of the workload: every result array is
    compared through its digest; the reference's wall time is the cpu_baseline.  Runs after the timed regions."""
    import numpy as np
    import torch

    from cstone_b200 import capi

    cfg = args.config
    n_s = min(args.ref_n, args.n)
    w = workload(cfg, args.n, 1)
    fake = argparse.Namespace(**vars(args))
    cb, out = run_reference(fake, n_s, 1, budget_s=60.0, steps=1, warmup=0)
    want = out[0]["digest"]
    dev = torch.device("cuda", 0)
    px, py, pz, ph = make_particles(cfg, n_s, 42)
    dx, dy, dz, dh = (torch.from_numpy(a).to(dev) for a in (px, py, pz, ph))
    if cfg == "morton":
        keys = torch.zeros(n_s, dtype=torch.uint32, device=dev)
        capi.compute_sfc_keys(dx, dy, dz, keys, w["lim"], w["bnd"], kind=1)
        order = capi.sequence(0, n_s, dev)
        capi.sort_by_key(keys, order)
        sx, sy, sz, sh = capi.gather_arrays4(order, [dx, dy, dz, dh])
        leaves, counts = capi.compute_octree(keys, BUCKET)
        tree = capi.Octree(leaves)
        cen, siz = capi.compute_geo_centers(tree.prefixes, torch.float32, w["lim"], w["bnd"], kind=1)
        layout = capi.exclusive_scan(torch.cat([counts, torch.zeros(1, dtype=torch.uint32, device=dev)]))
        nb, nc = capi.find_neighbors(sx, sy, sz, sh, 0, n_s, w["lim"], w["bnd"], tree, layout, cen, siz, w["ngmax"])
        got = {"keys": wsum(keys), "x": wsum(sx), "y": wsum(sy), "z": wsum(sz), "h": wsum(sh), "leaves": wsum(leaves),
               "num_leaves": tree.num_leaves, "layout": wsum(layout), "num_nodes": tree.num_nodes,
               "prefixes": wsum(tree.prefixes), "child_offsets": wsum(tree.child_offsets[: tree.num_nodes]),
               "centers": wsum(cen), "sizes": wsum(siz), "leaf_counts": wsum(counts)}
        got["nc_sum"], got["lists"] = list_digest(nb.reshape(-1), nc, w["ngmax"])
    else:
        dom = capi.Domain(0, 1, w["bucket"], w["bucket_focus"], 0.5, w["lim"], w["bnd"], key="u64", real="d",
                          device=str(dev))
        dom.sync(dx, dy, dz, dh)
        got = domain_digest(dom)
        if cfg == "plummer":
            _, halos, _, _ = halo_discovery_quarter(capi, torch, dom, w["bnd"])
            fl = halos()
            got["halo_flags"], got["halo_count"] = wsum(fl), int(fl.sum())
        else:
            nb, nc = dom.find_neighbors(w["ngmax"])
            got["nc_sum"], got["lists"] = list_digest(nb.reshape(-1), nc, w["ngmax"])
        dom.close()
    cmp_ = compare_digests(got, want)
    cmp_.update({"particles": n_s, "reference": "oracle/_ref (unmodified reference headers), same particles"})
    return cb, cmp_


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "parity-worker"])
    ap.add_argument("--config", default="uniform", choices=["uniform", "plummer", "morton"])
    ap.add_argument("--n", type=int, default=0, help="particles per GPU (default: the size BASELINE.json quotes)")
    ap.add_argument("--ref-n", type=int, default=16 * 1024 * 1024,
                    help="particles of the bounded CPU sample (cpu_baseline of our arm; --impl reference at N>1)")
    ap.add_argument("--parity-n", type=int, default=1 << 20, help="particles per rank of the N>1 transport parity run")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="wall-time budget of the --impl reference steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.n <= 0:
        args.n = 16 * 1024 * 1024 if args.config == "morton" else 64 * 1024 * 1024
    if args.config != "uniform" and (args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1):
        raise SystemExit("--config plummer/morton are single-GPU workloads (BASELINE.json configs[2], configs[4])")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line (the JSON); everything libraries print (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()

    if args.impl == "parity-worker":
        # child of multirank_parity: the reference Domain with P = --gpus thread-ranks on the parity particles
        import numpy as np
        P, n_s = args.gpus, args.parity_n
        _libs = _ref_env(max(1, os.cpu_count() // P), P)
        w = workload("uniform", n_s, P)
        parts = [make_particles("uniform", n_s, 1042 + r, n_for_h=n_s * P) for r in range(P)]
        cat = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        out = _libs.ref_bench_run("u64d", P, w["bucket"], w["bucket_focus"], 0.5, w["lim"], w["bnd"], *cat,
                                  [n_s * r for r in range(P + 1)], num_syncs=1, ngmax=64,
                                  threads=max(1, os.cpu_count() // P))
        emit([o["digest"] for o in out])
        return

    if args.impl == "reference":
        if rank != 0:
            return
        n_gpus = max(args.gpus, world)
        # N=1: the full workload.  N>1: 8 x 64 Mi particles do not fit a CPU run of a few minutes: reduced size, P = N
        n_ref = args.n if n_gpus == 1 else min(args.ref_n // 8, args.n)
        steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
        cb, out = run_reference(args, n_ref, n_gpus, args.ref_budget_s, steps, warmup)
        w = workload(args.config, args.n, n_gpus)
        digest = {k: (f"{v:#018x}" if v > 1 << 40 else v) for k, v in out[0]["digest"].items()} if n_gpus == 1 else None
        dtype = "u32 keys / f32 coordinates" if args.config == "morton" else "u64 keys / f64 coordinates"
        line = {"impl": "reference", "metric": METRIC, "value": round(cb["value"], 4), "unit": UNIT,
                "n_gpus": n_gpus, "steps": cb["steps_timed"], "warmup": warmup, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "ms_per_step": round(n_ref * n_gpus / cb["value"] / 1e3, 3), "config": w["config"],
                "results": {"digest": digest}, "cpu_baseline": cb,
                "e2e": {"value": round(cb["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    args.warmup = max(args.warmup, 3)
    line = {"uniform": run_uniform, "plummer": run_plummer, "morton": run_morton}[args.config](args)
    if line is None:
        return
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb, cmp_ = sample_parity(args)
            line["cpu_baseline"] = cb
            line["checks"]["sample_bit_identical_vs_reference"] = cmp_
        except Exception as e:  # the checker being unavailable must not hide the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": str(e)[:300]}
    emit(line)


if __name__ == "__main__":
    main()
