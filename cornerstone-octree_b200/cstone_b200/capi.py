"""ctypes binding of libcstone_b200.so (the C ABI declared in include/cstone_b200.h).

PyTorch is used only as the owner of device memory and streams: tensors are passed to the C ABI as raw device
pointers.  There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcstone_b200.so")

KEY_DTYPES = {"u32": torch.uint32, "u64": torch.uint64}
REAL_DTYPES = {"f": torch.float32, "d": torch.float64}
MAXLEVEL = {"u32": 10, "u64": 21}


class CstoneError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CstoneError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              "g.build()'` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.cs_last_error.restype = C.c_char_p
        _lib.cs_kernel_launch_count.restype = C.c_uint64
        for name in ("cs_sort_by_key_temp_bytes_u32", "cs_sort_by_key_temp_bytes_u64", "cs_scan_temp_bytes",
                     "cs_build_octree_temp_bytes_u32", "cs_build_octree_temp_bytes_u64", "cs_node_ops_temp_bytes"):
            getattr(_lib, name).restype = C.c_size_t
        # experiments: CSB_TUNING="knob=value,knob=value" selects kernel variants (csb::TuningKnob) for this process
        for item in filter(None, os.environ.get("CSB_TUNING", "").split(",")):
            knob, value = item.split("=")
            tuning_set(int(knob), int(value))
    return _lib


def tuning_set(knob, value):
    """experiment hook, see csb::TuningKnob in csrc/common.cuh"""
    _check(_lib.cs_tuning_set(C.c_int(knob), C.c_int(value)), "cs_tuning_set")


def _check(status, what):
    if status != 0:
        raise CstoneError(f"{what} failed with status {status}: {lib().cs_last_error().decode()}")


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensors only"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _box(lim, bnd):
    lim_a = (C.c_double * 6)(*[float(v) for v in lim])
    bnd_a = (C.c_int * 3)(*[int(v) for v in bnd])
    return lim_a, bnd_a


def key_suffix(t):
    return {torch.uint32: "u32", torch.int32: "u32", torch.uint64: "u64", torch.int64: "u64"}[t.dtype]


def real_suffix(t):
    return {torch.float32: "f", torch.float64: "d"}[t.dtype]


def kernel_launch_count():
    return int(lib().cs_kernel_launch_count())


# ---------------------------------------------------------------- SFC keys
def compute_sfc_keys(x, y, z, keys, lim, bnd, kind=0, n=None):
    """keys[i] = sfc3D(x[i], y[i], z[i], box) in place (removeKey entries are preserved)"""
    n = x.numel() if n is None else n
    combo = key_suffix(keys) + real_suffix(x)
    lim_a, bnd_a = _box(lim, bnd)
    f = getattr(lib(), "cs_compute_sfc_keys_" + combo)
    _check(f(C.c_int(kind), _ptr(x), _ptr(y), _ptr(z), _ptr(keys), C.c_size_t(n), lim_a, bnd_a, _stream()),
           "cs_compute_sfc_keys_" + combo)
    return keys


# ---------------------------------------------------------------- sort / gather / scan
def sort_by_key(keys, values=None):
    """stable ascending sort of keys (and values) in place"""
    n = keys.numel()
    kt = key_suffix(keys)
    tmp_bytes = getattr(lib(), "cs_sort_by_key_temp_bytes_" + kt)(C.c_size_t(n))
    key_buf = torch.empty_like(keys)
    val_buf = torch.empty_like(values) if values is not None else None
    tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=keys.device)
    f = getattr(lib(), "cs_sort_by_key_" + kt)
    _check(f(_ptr(keys), _ptr(values), C.c_size_t(n), _ptr(key_buf), _ptr(val_buf), _ptr(tmp), C.c_size_t(tmp_bytes),
             _stream()), "cs_sort_by_key_" + kt)
    return keys, values


def sequence(start, n, device):
    out = torch.empty(n, dtype=torch.uint32, device=device)
    _check(lib().cs_sequence_u32(C.c_uint32(start), C.c_size_t(n), _ptr(out), _stream()), "cs_sequence_u32")
    return out


def gather(ordering, src):
    dst = torch.empty(ordering.numel(), dtype=src.dtype, device=src.device)
    _check(lib().cs_gather(_ptr(ordering), C.c_size_t(ordering.numel()), _ptr(src), _ptr(dst),
                           C.c_int(src.element_size()), _stream()), "cs_gather")
    return dst


def gather4(ordering, srcs):
    n = ordering.numel()
    dsts = [torch.empty(n, dtype=s.dtype, device=s.device) for s in srcs]
    src_a = (C.c_void_p * 4)(*[s.data_ptr() for s in srcs])
    dst_a = (C.c_void_p * 4)(*[d.data_ptr() for d in dsts])
    _check(lib().cs_gather4(_ptr(ordering), C.c_size_t(n), src_a, dst_a, C.c_int(srcs[0].element_size()), _stream()),
           "cs_gather4")
    return dsts


def gather_arrays4(ordering, srcs):
    """gatherArrays(x,y,z,h): record-packed gather of four equally long arrays (cs_gather_arrays4)"""
    n = ordering.numel()
    dsts = [torch.empty(n, dtype=s.dtype, device=s.device) for s in srcs]
    src_a = (C.c_void_p * 4)(*[s.data_ptr() for s in srcs])
    dst_a = (C.c_void_p * 4)(*[d.data_ptr() for d in dsts])
    _check(lib().cs_gather_arrays4(_ptr(ordering), C.c_size_t(n), C.c_size_t(srcs[0].numel()), src_a, dst_a,
                                   C.c_int(srcs[0].element_size()), _stream()), "cs_gather_arrays4")
    return dsts


def exclusive_scan(values):
    n = values.numel()
    out = torch.empty_like(values)
    tmp = torch.empty(lib().cs_scan_temp_bytes(C.c_size_t(n)), dtype=torch.uint8, device=values.device)
    _check(lib().cs_exclusive_scan_u32(_ptr(values), _ptr(out), C.c_size_t(n), _ptr(tmp), _stream()),
           "cs_exclusive_scan_u32")
    return out


# ---------------------------------------------------------------- csarray
def compute_node_counts(leaves, keys, max_count=0xFFFFFFFF, n=None):
    kt = key_suffix(leaves)
    nl = leaves.numel() - 1
    n = keys.numel() if n is None else n
    counts = torch.empty(nl, dtype=torch.uint32, device=leaves.device)
    f = getattr(lib(), "cs_compute_node_counts_" + kt)
    _check(f(_ptr(leaves), _ptr(counts), C.c_int(nl), _ptr(keys), C.c_size_t(n), C.c_uint32(max_count), _stream()),
           "cs_compute_node_counts_" + kt)
    return counts


def compute_node_ops(leaves, counts, bucket):
    """returns (scanned nodeOps[numLeaves+1], newNumLeaves, converged)"""
    kt = key_suffix(leaves)
    nl = leaves.numel() - 1
    ops = torch.empty(nl + 1, dtype=torch.int32, device=leaves.device)
    tmp = torch.empty(lib().cs_node_ops_temp_bytes(C.c_int(nl)), dtype=torch.uint8, device=leaves.device)
    new_n, conv = C.c_int(0), C.c_int(0)
    f = getattr(lib(), "cs_compute_node_ops_" + kt)
    _check(f(_ptr(leaves), C.c_int(nl), _ptr(counts), C.c_uint32(bucket), _ptr(ops), _ptr(tmp), C.byref(new_n),
             C.byref(conv), _stream()), "cs_compute_node_ops_" + kt)
    return ops, new_n.value, bool(conv.value)


def rebalance_tree(leaves, ops, new_num_leaves):
    kt = key_suffix(leaves)
    nl = leaves.numel() - 1
    new_leaves = torch.empty(new_num_leaves + 1, dtype=leaves.dtype, device=leaves.device)
    f = getattr(lib(), "cs_rebalance_tree_" + kt)
    _check(f(_ptr(leaves), C.c_int(nl), C.c_int(new_num_leaves), _ptr(ops), _ptr(new_leaves), _stream()),
           "cs_rebalance_tree_" + kt)
    return new_leaves


def compute_octree(keys, bucket, capacity=None, n=None):
    """converged cornerstone leaf array + counts for sorted keys"""
    kt = key_suffix(keys)
    n = keys.numel() if n is None else n
    capacity = capacity or max(4096, 16 * (n // max(1, bucket)) + 4096)
    while True:
        leaves = torch.empty(capacity + 1, dtype=keys.dtype, device=keys.device)
        counts = torch.empty(capacity, dtype=torch.uint32, device=keys.device)
        nl = C.c_int(0)
        f = getattr(lib(), "cs_compute_octree_" + kt)
        st = f(_ptr(keys), C.c_size_t(n), C.c_uint32(bucket), _ptr(leaves), _ptr(counts), C.c_int(capacity),
               C.byref(nl), _stream())
        if st == 3:
            capacity = int(nl.value * 1.5) + 16
            continue
        _check(st, "cs_compute_octree_" + kt)
        return leaves[:nl.value + 1].clone(), counts[:nl.value].clone()


# ---------------------------------------------------------------- octree
class Octree:
    """device-resident OctreeData (tree/octree.hpp:285-360)"""

    def __init__(self, leaves):
        kt = key_suffix(leaves)
        dev = leaves.device
        nl = leaves.numel() - 1
        self.num_leaves = nl
        self.num_internal = (nl - 1) // 7
        self.num_nodes = nl + self.num_internal
        nn = self.num_nodes
        self.leaves = leaves
        self.prefixes = torch.empty(nn, dtype=leaves.dtype, device=dev)
        self.child_offsets = torch.empty(nn + 1, dtype=torch.int32, device=dev)
        self.parents = torch.empty(max(1, (nn - 1) // 8), dtype=torch.int32, device=dev)
        self.level_range = torch.empty(MAXLEVEL[kt] + 2, dtype=torch.int32, device=dev)
        self.internal_to_leaf = torch.empty(nn, dtype=torch.int32, device=dev)
        self.leaf_to_internal = torch.empty(nn, dtype=torch.int32, device=dev)
        tmp_bytes = getattr(lib(), "cs_build_octree_temp_bytes_" + kt)(C.c_int(nl))
        tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
        f = getattr(lib(), "cs_build_octree_" + kt)
        _check(f(_ptr(leaves), C.c_int(nl), _ptr(self.prefixes), _ptr(self.child_offsets), _ptr(self.parents),
                 _ptr(self.level_range), _ptr(self.internal_to_leaf), _ptr(self.leaf_to_internal), _ptr(tmp),
                 C.c_size_t(tmp_bytes), _stream()), "cs_build_octree_" + kt)


def compute_geo_centers(prefixes, real_dtype, lim, bnd, kind=0):
    n = prefixes.numel()
    combo = key_suffix(prefixes) + {torch.float32: "f", torch.float64: "d"}[real_dtype]
    centers = torch.empty((n, 3), dtype=real_dtype, device=prefixes.device)
    sizes = torch.empty((n, 3), dtype=real_dtype, device=prefixes.device)
    lim_a, bnd_a = _box(lim, bnd)
    f = getattr(lib(), "cs_compute_geo_centers_" + combo)
    _check(f(C.c_int(kind), _ptr(prefixes), C.c_int(n), _ptr(centers), _ptr(sizes), lim_a, bnd_a, _stream()),
           "cs_compute_geo_centers_" + combo)
    return centers, sizes


def upsweep_sum(kt, level_range_host, child_offsets, counts):
    lr = (C.c_int * len(level_range_host))(*[int(v) for v in level_range_host])
    _check(lib().cs_upsweep_sum(C.c_int(MAXLEVEL[kt]), lr, _ptr(child_offsets), _ptr(counts), _stream()),
           "cs_upsweep_sum")
    return counts


# ---------------------------------------------------------------- halos
def compute_bounding_boxes(x, y, z, h, layout, first, last, scale, search_centers):
    """in: search_centers[leaf] = init point; out: (centers, sizes) of the per-leaf search boxes"""
    sfx = real_suffix(x)
    sc = search_centers.clone()
    ss = torch.zeros_like(sc)
    scale_arg = C.c_float(scale) if sfx == "f" else C.c_double(scale)
    f = getattr(lib(), "cs_compute_bounding_boxes_" + sfx)
    _check(f(_ptr(x), _ptr(y), _ptr(z), _ptr(h), _ptr(layout), C.c_int(first), C.c_int(last), scale_arg, _ptr(sc),
             _ptr(ss), _stream()), "cs_compute_bounding_boxes_" + sfx)
    return sc, ss


def find_halos(tree, centers, sizes, sc, ss, lim, bnd, first, last, flags=None):
    combo = key_suffix(tree.prefixes) + real_suffix(centers)
    if flags is None:
        flags = torch.zeros(tree.num_nodes, dtype=torch.uint8, device=centers.device)
    lim_a, bnd_a = _box(lim, bnd)
    f = getattr(lib(), "cs_find_halos_" + combo)
    _check(f(_ptr(tree.prefixes), _ptr(tree.child_offsets), _ptr(tree.parents), _ptr(centers), _ptr(sizes),
             _ptr(tree.leaves), _ptr(sc), _ptr(ss), lim_a, bnd_a, C.c_int(first), C.c_int(last), _ptr(flags),
             _stream()), "cs_find_halos_" + combo)
    return flags


# ---------------------------------------------------------------- neighbours
def find_neighbors(x, y, z, h, first, last, lim, bnd, tree, layout, centers, sizes, ngmax, neighbors=None,
                   counts=None, search_ext_factor=1.0):
    sfx = real_suffix(x)
    nloc = last - first
    if neighbors is None:
        neighbors = torch.zeros(nloc * ngmax, dtype=torch.uint32, device=x.device)
    if counts is None:
        counts = torch.zeros(nloc, dtype=torch.uint32, device=x.device)
    lim_a, bnd_a = _box(lim, bnd)
    if search_ext_factor != 1.0:  # OctreeNsView::searchExtFactor
        f = getattr(lib(), "cs_find_neighbors_ext_" + sfx)
        _check(f(_ptr(x), _ptr(y), _ptr(z), _ptr(h), C.c_uint32(first), C.c_uint32(last), lim_a, bnd_a,
                 C.c_int(tree.num_leaves), _ptr(tree.child_offsets), _ptr(tree.parents), _ptr(tree.internal_to_leaf),
                 _ptr(layout), _ptr(centers), _ptr(sizes), C.c_uint32(ngmax), _ptr(neighbors), _ptr(counts),
                 C.c_float(search_ext_factor), _stream()), "cs_find_neighbors_ext_" + sfx)
        return neighbors, counts
    f = getattr(lib(), "cs_find_neighbors_" + sfx)
    _check(f(_ptr(x), _ptr(y), _ptr(z), _ptr(h), C.c_uint32(first), C.c_uint32(last), lim_a, bnd_a,
             C.c_int(tree.num_leaves), _ptr(tree.child_offsets), _ptr(tree.parents), _ptr(tree.internal_to_leaf), _ptr(layout), _ptr(centers),
             _ptr(sizes), C.c_uint32(ngmax), _ptr(neighbors), _ptr(counts), _stream()), "cs_find_neighbors_" + sfx)
    return neighbors.view(nloc, ngmax), counts


# ---------------------------------------------------------------- Domain
FIELDS = ["x", "y", "z", "h", "keys", "focus_leaves", "focus_leaf_counts", "focus_node_counts", "layout", "prefixes",
          "child_offsets", "parents", "level_range", "internal_to_leaf", "leaf_to_internal", "geo_centers",
          "geo_sizes", "halo_flags", "global_leaves", "global_counts", "global_prefixes", "global_child_offsets"]


class _View:
    """zero-copy torch view of a device array owned by the C library"""

    def __init__(self, ptr, shape, dtype, device):
        self.ptr, self.shape, self.dtype, self.device = ptr, shape, dtype, device

    @property
    def __cuda_array_interface__(self):
        # exported as a signed integer type of the same width and re-viewed by the caller: the unsigned 32/64-bit
        # dtypes are not accepted by every torch version on this path
        typestr = {1: "|u1", 4: "<i4", 8: "<i8"}[torch.empty(0, dtype=self.dtype).element_size()]
        return {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 2}


class LocalWorld:
    """ranks as threads of this process (cs_local_world_*): lets multi-rank domains run on one GPU"""

    def __init__(self, size):
        lib().cs_local_world_create.restype = C.c_void_p
        self.size = size
        self.handle = lib().cs_local_world_create(C.c_int(size))
        if not self.handle:
            raise CstoneError(lib().cs_last_error().decode())

    def abort(self):
        """wake the ranks that wait for a rank that has failed (they return an error)"""
        if getattr(self, "handle", None):
            lib().cs_local_world_abort(C.c_void_p(self.handle))

    def comm(self, rank):
        lib().cs_comm_create_local.restype = C.c_void_p
        h = lib().cs_comm_create_local(C.c_void_p(self.handle), C.c_int(rank))
        if not h:
            raise CstoneError(lib().cs_last_error().decode())
        return Comm(h, rank, self.size, keepalive=self)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().cs_local_world_destroy(C.c_void_p(self.handle))
                self.handle = None
        except Exception:
            pass


class Comm:
    """what MPI_Comm is to the reference's Domain: cs_comm_t (local threads or NCCL)"""

    def __init__(self, handle, rank, size, keepalive=None):
        self.handle, self.rank, self.size, self._keepalive = handle, rank, size, keepalive

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_ubyte * 128)()
        _check(lib().cs_nccl_unique_id(buf), "cs_nccl_unique_id")
        return bytes(buf)

    @staticmethod
    def nccl(rank, size, unique_id):
        """collective over all ranks: one process per GPU, the current CUDA device is the rank's GPU"""
        lib().cs_comm_create_nccl.restype = C.c_void_p
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        h = lib().cs_comm_create_nccl(C.c_int(rank), C.c_int(size), buf)
        if not h:
            raise CstoneError(lib().cs_last_error().decode())
        return Comm(h, rank, size)

    @property
    def bytes_sent(self):
        lib().cs_comm_bytes_sent.restype = C.c_uint64
        return int(lib().cs_comm_bytes_sent(C.c_void_p(self.handle)))

    def close(self):
        if getattr(self, "handle", None):
            lib().cs_comm_destroy(C.c_void_p(self.handle))
            self.handle = None


class Domain:
    """host-side mirror of cstone::Domain<KeyType, T, Gpu> (domain/domain.hpp:38-664) over the cs_domain_* C ABI"""

    def __init__(self, rank, num_ranks, bucket_size, bucket_size_focus, theta, lim, bnd, key="u64", real="d",
                 device="cuda:0", comm=None):
        self.combo = key + real
        self.comm = comm
        self.kt, self.real = key, real
        self.device = torch.device(device)
        lim_a, bnd_a = _box(lim, bnd)
        f = getattr(lib(), "cs_domain_create_" + self.combo)
        f.restype = C.c_void_p
        with torch.cuda.device(self.device):
            self.handle = f(C.c_int(rank), C.c_int(num_ranks), C.c_uint(bucket_size), C.c_uint(bucket_size_focus),
                            C.c_float(theta), lim_a, bnd_a)
        if not self.handle:
            # the reference throws std::runtime_error from the constructor (domain.hpp:81-85)
            raise CstoneError(lib().cs_last_error().decode())
        lib().cs_domain_ptr.restype = C.c_void_p
        if comm is not None:
            _check(lib().cs_domain_attach_comm(C.c_void_p(self.handle), C.c_void_p(comm.handle)),
                   "cs_domain_attach_comm")

    def close(self):
        if getattr(self, "handle", None):
            lib().cs_domain_destroy(C.c_void_p(self.handle))
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def sync(self, x=None, y=None, z=None, h=None, keys=None):
        """x,y,z,h: torch tensors on the domain's device or in (pinned) host memory, or None to re-sync in place"""
        if x is None:
            args = [C.c_void_p(0)] * 5 + [C.c_size_t(0), C.c_int(0)]
        else:
            host = not x.is_cuda
            ptrs = [C.c_void_p(t.data_ptr()) for t in (x, y, z, h)]
            ptrs.append(C.c_void_p(keys.data_ptr()) if keys is not None else C.c_void_p(0))
            args = ptrs + [C.c_size_t(x.numel()), C.c_int(1 if host else 0)]
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_sync(C.c_void_p(self.handle), *args, _stream()), "cs_domain_sync")
        self._info()

    def reset(self):
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_reset(C.c_void_p(self.handle), _stream()), "cs_domain_reset")

    def _info(self):
        out = (C.c_uint64 * 8)()
        box = (C.c_double * 6)()
        _check(lib().cs_domain_info(C.c_void_p(self.handle), out, box), "cs_domain_info")
        (self.start_index, self.end_index, self.n_particles_with_halos, self.num_focus_leaves, self.num_focus_nodes,
         self.num_global_leaves, self.num_global_nodes, self.max_level) = [int(v) for v in out]
        self.box = tuple(box)

    def field(self, name):
        """torch view (no copy) of one of the arrays listed in FIELDS"""
        real_t = REAL_DTYPES[self.real]
        key_t = KEY_DTYPES[self.kt]
        n, nl, nn = self.n_particles_with_halos, self.num_focus_leaves, self.num_focus_nodes
        gl, gn = self.num_global_leaves, self.num_global_nodes
        spec = {
            "x": ((n,), real_t), "y": ((n,), real_t), "z": ((n,), real_t), "h": ((n,), real_t), "keys": ((n,), key_t),
            "focus_leaves": ((nl + 1,), key_t), "focus_leaf_counts": ((nl,), torch.uint32),
            "focus_node_counts": ((nn,), torch.uint32), "layout": ((nl + 1,), torch.uint32),
            "prefixes": ((nn,), key_t), "child_offsets": ((nn,), torch.int32),
            "parents": ((max(1, (nn - 1) // 8),), torch.int32), "level_range": ((self.max_level + 2,), torch.int32),
            "internal_to_leaf": ((nn,), torch.int32), "leaf_to_internal": ((nn,), torch.int32),
            "geo_centers": ((nn, 3), real_t), "geo_sizes": ((nn, 3), real_t), "halo_flags": ((nn,), torch.uint8),
            "global_leaves": ((gl + 1,), key_t), "global_counts": ((gl,), torch.uint32),
            "global_prefixes": ((gn,), key_t), "global_child_offsets": ((gn,), torch.int32),
        }[name]
        ptr = lib().cs_domain_ptr(C.c_void_p(self.handle), C.c_int(FIELDS.index(name)))
        shape, dtype = spec
        if 0 in shape or not ptr:
            return torch.empty(shape, dtype=dtype, device=self.device)
        return torch.as_tensor(_View(ptr, shape, dtype, self.device), device=self.device).view(dtype)

    def find_neighbors(self, ngmax, neighbors=None, counts=None):
        nloc = self.end_index - self.start_index
        if neighbors is None:
            neighbors = torch.zeros(nloc * ngmax, dtype=torch.uint32, device=self.device)
        if counts is None:
            counts = torch.zeros(nloc, dtype=torch.uint32, device=self.device)
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_find_neighbors(C.c_void_p(self.handle), C.c_uint32(ngmax), _ptr(neighbors),
                                                  _ptr(counts), _stream()), "cs_domain_find_neighbors")
        return neighbors[: nloc * ngmax].view(nloc, ngmax), counts[:nloc]

    def exchange_halos(self, *fields):
        """Domain::exchangeHalos: fields are device tensors with n_particles_with_halos rows; halo rows are filled in"""
        n = len(fields)
        for f in fields:
            assert f.is_cuda and f.is_contiguous() and f.shape[0] == self.n_particles_with_halos
        ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in fields])
        sizes = (C.c_int * n)(*[f.element_size() * (f.numel() // max(f.shape[0], 1)) for f in fields])
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_exchange_halos(C.c_void_p(self.handle), ptrs, sizes, C.c_int(n), _stream()),
                   "cs_domain_exchange_halos")

    def set_halo_factor(self, factor):
        """Domain::setHaloFactor"""
        _check(lib().cs_domain_set_halo_factor(C.c_void_p(self.handle), C.c_float(factor)), "cs_domain_set_halo_factor")

    def reapply_sync(self, *fields):
        """Domain::reapplySync: fields are device tensors in the particle order the last sync() consumed; returns new
        tensors with n_particles_with_halos rows whose assigned rows [start_index, end_index) hold the fields of the
        particles now assigned to this rank (halo rows are zero until exchange_halos)."""
        info = (C.c_uint64 * 4)()
        _check(lib().cs_domain_replay_info(C.c_void_p(self.handle), info), "cs_domain_replay_info")
        n = len(fields)
        outs = []
        for f in fields:
            assert f.is_cuda and f.is_contiguous() and f.shape[0] == info[0], (f.shape, info[0])
            outs.append(torch.zeros((int(info[1]),) + tuple(f.shape[1:]), dtype=f.dtype, device=f.device))
        src = (C.c_void_p * n)(*[f.data_ptr() for f in fields])
        dst = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        sizes = (C.c_int * n)(*[f.element_size() * (f.numel() // max(f.shape[0], 1)) for f in fields])
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_reapply_sync(C.c_void_p(self.handle), src, dst, sizes, C.c_int(n), _stream()),
                   "cs_domain_reapply_sync")
        return outs

    def download(self, x, y, z, h, keys):
        """asynchronous device -> (pinned) host copy of the synchronised arrays"""
        ptrs = [C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0) for t in (x, y, z, h, keys)]
        with torch.cuda.device(self.device):
            _check(lib().cs_domain_download(C.c_void_p(self.handle), *ptrs, _stream()), "cs_domain_download")
