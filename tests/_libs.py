"""ctypes bindings used by the tests: the C oracle (oracle/liboracle.so) and, when it was built, the unmodified
reference behind oracle/_ref/libcstone_ref.so.  Both are TEST INFRASTRUCTURE: the product (cstone_b200) never
imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcstone_ref.so")

KEYS = {"u32": np.uint32, "u64": np.uint64}
REALS = {"f": np.float32, "d": np.float64}
MAXLEVEL = {"u32": 10, "u64": 21}
COMBOS = ["u32f", "u64f", "u64d"]


def key_of(combo):
    return combo[:3]


def real_of(combo):
    return REALS[combo[3]]


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _build_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), ORACLE_SO])


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        _build_oracle()
        _oracle = C.CDLL(ORACLE_SO)
    return _oracle


def ref_lib():
    """the reference itself; None when oracle/_ref was never built (no /root/reference on this machine)"""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = C.CDLL(REF_SO)
    return _ref


def box_args(lim, bnd):
    lim = np.ascontiguousarray(lim, dtype=np.float64)
    bnd = np.ascontiguousarray(bnd, dtype=np.int32)
    assert lim.size == 6 and bnd.size == 3
    return lim, bnd


class _Backend:
    """common numpy-level API over liboracle.so (prefix 'orc_') and libcstone_ref.so (prefix 'ref_')"""

    def __init__(self, lib, prefix):
        self.lib = lib
        self.prefix = prefix

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    # ---- keys
    def sfc_keys(self, combo, kind, x, y, z, lim, bnd, keys=None):
        K = KEYS[key_of(combo)]
        n = x.size
        if keys is None:
            keys = np.zeros(n, dtype=K)
        lim, bnd = box_args(lim, bnd)
        self._fn("sfc_keys_" + combo)(C.c_int(kind), _p(x), _p(y), _p(z), _p(keys), C.c_size_t(n), _p(lim), _p(bnd))
        return keys

    def sort_by_key(self, kt, keys, values):
        self._fn("sort_by_key_" + kt)(_p(keys), _p(values), C.c_size_t(keys.size))

    def compute_octree(self, kt, keys, bucket, cap=None):
        K = KEYS[kt]
        cap = cap or max(64, 4 * keys.size // max(1, bucket) * 8 + 4096)
        leaves = np.zeros(cap + 1, dtype=K)
        counts = np.zeros(cap, dtype=np.uint32)
        n = self._fn("compute_octree_" + kt, C.c_long)(_p(keys), C.c_size_t(keys.size), C.c_uint(bucket), _p(leaves),
                                                       _p(counts), C.c_long(cap))
        assert n > 0, n
        return leaves[:n + 1].copy(), counts[:n].copy()

    def update_octree(self, kt, keys, bucket, leaves, counts, cap=None):
        K = KEYS[kt]
        nl = counts.size
        cap = cap or (4096 * nl + 16)
        lv = np.zeros(cap + 1, dtype=K)
        ct = np.zeros(cap, dtype=np.uint32)
        lv[:nl + 1] = leaves
        ct[:nl] = counts
        conv = C.c_int(0)
        n = self._fn("update_octree_" + kt, C.c_long)(_p(keys), C.c_size_t(keys.size), C.c_uint(bucket), _p(lv), _p(ct),
                                                      C.c_long(nl), C.c_long(cap), C.byref(conv))
        assert n > 0, n
        return lv[:n + 1].copy(), ct[:n].copy(), bool(conv.value)

    def compute_node_counts(self, kt, leaves, keys, max_count=0xFFFFFFFF):
        nl = leaves.size - 1
        counts = np.zeros(nl, dtype=np.uint32)
        self._fn("compute_node_counts_" + kt)(_p(leaves), _p(counts), C.c_int(nl), _p(keys), C.c_size_t(keys.size),
                                              C.c_uint(max_count))
        return counts

    def rebalance_decision(self, kt, leaves, counts, bucket):
        nl = leaves.size - 1
        ops = np.zeros(nl + 1, dtype=np.int32)
        conv = self._fn("rebalance_decision_" + kt, C.c_int)(_p(leaves), _p(counts), C.c_int(nl), C.c_uint(bucket),
                                                             _p(ops))
        return ops[:nl], bool(conv)

    def build_octree(self, kt, leaves):
        K = KEYS[kt]
        nl = leaves.size - 1
        ni = (nl - 1) // 7
        nn = nl + ni
        out = dict(
            prefixes=np.zeros(nn, dtype=K),
            childOffsets=np.zeros(nn, dtype=np.int32),
            parents=np.zeros(max(1, (nn - 1) // 8), dtype=np.int32),
            levelRange=np.zeros(MAXLEVEL[kt] + 2, dtype=np.int32),
            internalToLeaf=np.zeros(nn, dtype=np.int32),
            leafToInternal=np.zeros(nn, dtype=np.int32),
        )
        self._fn("build_octree_" + kt)(_p(leaves), C.c_int(nl), _p(out["prefixes"]), _p(out["childOffsets"]),
                                       _p(out["parents"]), _p(out["levelRange"]), _p(out["internalToLeaf"]),
                                       _p(out["leafToInternal"]))
        out["parents"] = out["parents"][:(nn - 1) // 8]
        out["numLeaves"], out["numInternal"], out["numNodes"] = nl, ni, nn
        return out

    def node_fp_centers(self, combo, prefixes, lim, bnd, kind=0):
        T = real_of(combo)
        n = prefixes.size
        centers = np.zeros((n, 3), dtype=T)
        sizes = np.zeros((n, 3), dtype=T)
        lim, bnd = box_args(lim, bnd)
        if self.prefix == "orc_":
            self._fn("node_fp_centers_" + combo)(C.c_int(kind), _p(prefixes), C.c_size_t(n), _p(centers), _p(sizes),
                                                 _p(lim), _p(bnd))
        else:
            assert kind == 0, "the reference decodes node boxes as Hilbert only (SURVEY H8)"
            self._fn("node_fp_centers_" + combo)(_p(prefixes), C.c_size_t(n), _p(centers), _p(sizes), _p(lim), _p(bnd))
        return centers, sizes

    def bounding_boxes(self, combo, x, y, z, h, layout, first, last, scale, init_centers):
        T = real_of(combo)
        sc = np.ascontiguousarray(init_centers, dtype=T).copy()
        ss = np.zeros_like(sc)
        if self.prefix == "orc_":
            f = self._fn("bounding_boxes_" + combo)
        else:
            f = self._fn("bounding_boxes_" + combo[3])
        sarg = C.c_float(scale) if T == np.float32 else C.c_double(scale)
        f(_p(x), _p(y), _p(z), _p(h), _p(layout), C.c_int(first), C.c_int(last), sarg, _p(sc), _p(ss))
        return sc, ss

    def find_halos(self, combo, tree, centers, sizes, leaves, sc, ss, lim, bnd, first, last, flags=None):
        if flags is None:
            flags = np.zeros(tree["numNodes"], dtype=np.uint8)
        lim, bnd = box_args(lim, bnd)
        self._fn("find_halos_" + combo)(_p(tree["prefixes"]), _p(tree["childOffsets"]), _p(tree["parents"]),
                                        _p(centers), _p(sizes), _p(leaves), _p(sc), _p(ss), _p(lim), _p(bnd),
                                        C.c_int(first), C.c_int(last), _p(flags))
        return flags

    def find_neighbors(self, combo, x, y, z, h, first, last, lim, bnd, tree, leaves, layout, centers, sizes, ngmax):
        nloc = last - first
        nb = np.zeros(nloc * ngmax, dtype=np.uint32)
        nc = np.zeros(nloc, dtype=np.uint32)
        lim, bnd = box_args(lim, bnd)
        if self.prefix == "orc_":
            self._fn("find_neighbors_" + combo)(_p(x), _p(y), _p(z), _p(h), C.c_uint(first), C.c_uint(last), _p(lim),
                                                _p(bnd), _p(tree["childOffsets"]), _p(tree["parents"]),
                                                _p(tree["internalToLeaf"]), _p(layout), _p(centers), _p(sizes),
                                                C.c_uint(ngmax), _p(nb), _p(nc))
        else:
            self._fn("find_neighbors_" + combo)(_p(x), _p(y), _p(z), _p(h), C.c_uint(first), C.c_uint(last), _p(lim),
                                                _p(bnd), C.c_int(tree["numLeaves"]), C.c_int(tree["numNodes"]),
                                                _p(tree["prefixes"]), _p(tree["childOffsets"]), _p(tree["parents"]),
                                                _p(tree["internalToLeaf"]), _p(tree["leafToInternal"]),
                                                _p(tree["levelRange"]), _p(leaves), _p(layout), _p(centers), _p(sizes),
                                                C.c_uint(ngmax), _p(nb), _p(nc))
        return nb.reshape(nloc, ngmax), nc


def oracle():
    return _Backend(oracle_lib(), "orc_")


def ref():
    lib = ref_lib()
    return _Backend(lib, "ref_") if lib is not None else None


# ---- scalar helpers of the oracle (used by the golden-vector tests)
def orc_scalar(name, kt, *args):
    lib = oracle_lib()
    f = getattr(lib, f"orc_{name}_{kt}")
    f.restype = C.c_uint32 if kt == "u32" else C.c_uint64
    return f(*args)


def orc_decode(name, kt, key):
    lib = oracle_lib()
    f = getattr(lib, f"orc_{name}_{kt}")
    x, y, z = C.c_uint(), C.c_uint(), C.c_uint()
    f((C.c_uint32 if kt == "u32" else C.c_uint64)(key), C.byref(x), C.byref(y), C.byref(z))
    return x.value, y.value, z.value


# ---- the reference Domain driver (threads as ranks), see oracle/ref_api.cpp
def ref_domain_run(combo, P, bucket, bucket_focus, theta, lim, bnd, x, y, z, h, offsets, num_syncs=1, ngmax=0,
                   moves=None):
    lib = ref_lib()
    assert lib is not None
    lim, bnd = box_args(lim, bnd)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    f = getattr(lib, "ref_domain_run_" + combo)
    f.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float] + [C.c_void_p] * 7 + [C.c_int, C.c_uint, C.c_void_p]
    f(P, bucket, bucket_focus, theta, _p(lim), _p(bnd), _p(x), _p(y), _p(z), _p(h), _p(offsets), num_syncs, ngmax,
      _p(moves))
    K = KEYS[key_of(combo)]
    T = real_of(combo)
    fields = dict(keys=(np.uint64, K), x=(np.float64, T), y=(np.float64, T), z=(np.float64, T), h=(np.float64, T),
                  focus_leaves=(np.uint64, K), global_leaves=(np.uint64, K), focus_counts=(np.uint32, np.uint32),
                  layout=(np.uint32, np.uint32), prefixes=(np.uint64, K), child_offsets=(np.int32, np.int32),
                  parents=(np.int32, np.int32), internal_to_leaf=(np.int32, np.int32),
                  leaf_to_internal=(np.int32, np.int32), level_range=(np.int32, np.int32),
                  centers=(np.float64, T), sizes=(np.float64, T), neighbors=(np.uint32, np.uint32),
                  neighbors_count=(np.uint32, np.uint32), flags=(np.uint8, np.uint8))
    out = []
    for r in range(P):
        d = {}
        for name, (ft, tt) in fields.items():
            g = getattr(lib, "ref_domain_get_" + name)
            g.restype = C.c_long
            g.argtypes = [C.c_int, C.c_void_p, C.c_long]
            n = g(r, None, 0)
            buf = np.zeros(n, dtype=ft)
            g(r, _p(buf), n)
            d[name] = buf.astype(tt)
        se = np.zeros(2, dtype=np.uint32)
        box = np.zeros(6)
        ts = np.zeros(8)
        tn = C.c_double()
        lib.ref_domain_get_info(C.c_int(r), _p(se), _p(box), _p(ts), C.byref(tn))
        d.update(start=int(se[0]), end=int(se[1]), box=box, t_sync=ts, t_neighbors=tn.value)
        out.append(d)
    return out


# ---- bench / full-size parity drivers of oracle/ref_api.cpp: wall times + position-weighted digests, nothing copied out
DIGEST_SLOTS = ["keys", "x", "y", "z", "h", "leaves", "num_leaves", "layout", "nc_sum", "lists", "start", "end", "size",
                "num_nodes", "prefixes", "child_offsets", "centers", "sizes", "halo_flags", "halo_count", "leaf_counts"]
DIGEST_C1, DIGEST_C2 = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F


def _digest_dict(raw):
    return {name: int(raw[i]) for i, name in enumerate(DIGEST_SLOTS)}


def ref_bench_run(combo, P, bucket, bucket_focus, theta, lim, bnd, x, y, z, h, offsets, num_syncs=1, ngmax=0,
                  chunk=1 << 20, threads=0, halo_quarter=False):
    """P thread-ranks of the reference Domain: num_syncs x sync (+ standalone halo discovery over the first quarter of
    the leaves, + findNeighbors in chunks); returns per rank {t_sync[4], t_neighbors, t_halos, digest{}}"""
    lib = ref_lib()
    assert lib is not None
    lim, bnd = box_args(lim, bnd)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    f = getattr(lib, "ref_bench_run_" + combo)
    f.argtypes = ([C.c_int, C.c_uint, C.c_uint, C.c_float] + [C.c_void_p] * 7 +
                  [C.c_int, C.c_uint, C.c_size_t, C.c_int, C.c_int])
    f(P, bucket, bucket_focus, theta, _p(lim), _p(bnd), _p(x), _p(y), _p(z), _p(h), _p(offsets), num_syncs, ngmax,
      chunk, threads, int(halo_quarter))
    out = []
    for r in range(P):
        times = np.zeros(6)
        dg = np.zeros(24, dtype=np.uint64)
        lib.ref_bench_get(C.c_int(r), _p(times), _p(dg))
        out.append(dict(t_sync=times[:4].copy(), t_neighbors=float(times[4]), t_halos=float(times[5]),
                        digest=_digest_dict(dg)))
    return out


def ref_bench_tree_neighbors(combo, kind, x, y, z, h, bucket, lim, bnd, ngmax, chunk=1 << 20):
    """standalone keys(kind) -> sort -> gather -> computeOctree -> link -> centres -> findNeighbors with the reference;
    returns {t_build, t_neighbors, digest{}}"""
    lib = ref_lib()
    assert lib is not None
    lim, bnd = box_args(lim, bnd)
    times = np.zeros(2)
    dg = np.zeros(24, dtype=np.uint64)
    f = getattr(lib, "ref_bench_tree_neighbors_" + combo)
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_size_t,
                                                  C.c_void_p, C.c_void_p]
    f(kind, _p(x), _p(y), _p(z), _p(h), x.size, bucket, _p(lim), _p(bnd), ngmax, chunk, _p(times), _p(dg))
    return dict(t_build=float(times[0]), t_neighbors=float(times[1]), digest=_digest_dict(dg))


def ref_plummer(n, T=np.float64):
    """the reference's Plummer sphere generator (test/coord_samples/plummer.hpp, srand48(42))"""
    lib = ref_lib()
    assert lib is not None
    x, y, z = (np.zeros(n, dtype=T) for _ in range(3))
    f = lib.ref_plummer_d if T == np.float64 else lib.ref_plummer_f
    f.argtypes = [C.c_size_t] + [C.c_void_p] * 3
    f(n, _p(x), _p(y), _p(z))
    return x, y, z
