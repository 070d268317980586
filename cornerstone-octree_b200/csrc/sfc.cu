/* SFC key generation for sm_100a: cs_compute_sfc_keys_* (replaces cstone::computeSfcKeys(Gpu,...),
 * reference sfc/sfc_gpu.cu:23-62, arithmetic of sfc/sfc.hpp:142-179, hilbert.hpp:43-94, morton.hpp:94-108).
 *
 * Design: the reference walks the Hilbert curve one level at a time (21 dependent iterations of ~20 integer ops);
 * here the curve is a finite state machine over (axis permutation, axis flips) advanced THREE levels per lookup
 * through a 512-entry-per-state table staged in shared memory, so a 64-bit key costs 7 LDS + ~100 integer ops and
 * the kernel stays HBM-bound (40 B/particle for u64/double).  x,y,z,key traffic uses 128-bit vector accesses.
 * The state machine (hilbert.cuh) is derived by composing the reference's per-level flip/rotate rules, so it encodes
 * exactly the same curve (checked bit-for-bit against the oracle in tests/).
 */
#include <mutex>
#include <vector>

#include "common.cuh"
#include "cstone_b200.h"
#include "hilbert.cuh"

namespace csb
{

namespace
{

/* ---------------------------------------------------------------- host: three-level table of the state machine */

struct HilbertLut
{
    int numStates{0};
    std::vector<uint16_t> table; // [state][cx<<6 | cy<<3 | cz] -> (nextState << 9) | 9 key bits
};

HilbertLut buildHilbertLut()
{
    // the one-level machine of hilbert.cuh, advanced three levels per table entry
    const HilbertFsm& fsm = hilbertFsm;
    HilbertLut lut;
    lut.numStates = fsm.numStates;
    lut.table.resize(size_t(lut.numStates) * 512);
    for (int s = 0; s < lut.numStates; ++s)
    {
        for (unsigned c = 0; c < 512; ++c)
        {
            unsigned cx = (c >> 6) & 7, cy = (c >> 3) & 7, cz = c & 7;
            int cur      = s;
            unsigned key = 0;
            for (int b = 2; b >= 0; --b)
            {
                unsigned octant = (((cx >> b) & 1) << 2) | (((cy >> b) & 1) << 1) | ((cz >> b) & 1);
                unsigned e      = fsm.enc[cur * 8 + octant];
                key             = (key << 3) | (e & 7u);
                cur             = int(e >> 3);
            }
            lut.table[size_t(s) * 512 + c] = uint16_t((cur << 9) | key);
        }
    }
    return lut;
}

struct DeviceLut
{
    uint16_t* d_table{nullptr};
    int numStates{0};
};

//! per-device LUT copy (one process may drive several GPUs)
int getDeviceLut(DeviceLut& out)
{
    static std::mutex mtx;
    static HilbertLut host;
    static DeviceLut perDevice[64];
    std::lock_guard<std::mutex> lk(mtx);
    if (host.numStates == 0) { host = buildHilbertLut(); }
    int dev = 0;
    CSB_CHECK(cudaGetDevice(&dev));
    CSB_REQUIRE(dev < 64, "device ordinal too large");
    CSB_REQUIRE(host.numStates <= hilbertMaxStates, "unexpected number of Hilbert states");
    if (!perDevice[dev].d_table)
    {
        CSB_CHECK(cudaMalloc(&perDevice[dev].d_table, host.table.size() * sizeof(uint16_t)));
        CSB_CHECK(cudaMemcpy(perDevice[dev].d_table, host.table.data(), host.table.size() * sizeof(uint16_t),
                             cudaMemcpyHostToDevice));
        perDevice[dev].numStates = host.numStates;
    }
    out = perDevice[dev];
    return 0;
}

/* ---------------------------------------------------------------- device */

template<class T>
struct KeyParams
{
    T m[3];     // 2^L * ilen
    T minm[3];  // lim_min * m
};

//! 3 levels per lookup; coordinates of 32-bit keys are treated as 12-level ones (two leading zero levels map the
//! state machine back to identity and contribute zero digits)
template<class K>
__device__ inline K hilbertLut(unsigned ix, unsigned iy, unsigned iz, const uint16_t* lut)
{
    constexpr int numChunks = (KeyTraits<K>::maxLevel + 2) / 3;
    K key                   = 0;
    unsigned state          = 0;
#pragma unroll
    for (int k = numChunks - 1; k >= 0; --k)
    {
        unsigned c = (((ix >> (3 * k)) & 7u) << 6) | (((iy >> (3 * k)) & 7u) << 3) | ((iz >> (3 * k)) & 7u);
        unsigned e = lut[state | c];
        key        = (key << 9) | K(e & 511u);
        state      = e & ~511u; // (nextState << 9)
    }
    return key;
}

__device__ inline uint32_t expandBits(uint32_t v)
{
    v &= 0x000003ffu;
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ inline uint64_t expandBits(uint64_t v)
{
    uint64_t x = v & 0x1fffffu;
    x          = (x | x << 32u) & 0x001f00000000ffffull;
    x          = (x | x << 16u) & 0x001f0000ff0000ffull;
    x          = (x | x << 8u) & 0x100f00f00f00f00full;
    x          = (x | x << 4u) & 0x10c30c30c30c30c3ull;
    x          = (x | x << 2u) & 0x1249249249249249ull;
    return x;
}

template<class K, class T, int KIND>
__device__ inline K encodeOne(T x, T y, T z, const KeyParams<T>& p, const uint16_t* lut)
{
    constexpr int mcoord = (1 << KeyTraits<K>::maxLevel) - 1;
    // int i = floor(x * m) - min * m : separate multiply, floor, subtract, then truncating conversion
    int ix = int(rfloor(x * p.m[0]) - p.minm[0]);
    int iy = int(rfloor(y * p.m[1]) - p.minm[1]);
    int iz = int(rfloor(z * p.m[2]) - p.minm[2]);
    ix     = min(ix, mcoord);
    iy     = min(iy, mcoord);
    iz     = min(iz, mcoord);
    if constexpr (KIND == 0) { return hilbertLut<K>(unsigned(ix), unsigned(iy), unsigned(iz), lut); }
    else { return expandBits(K(unsigned(ix))) * 4 + expandBits(K(unsigned(iy))) * 2 + expandBits(K(unsigned(iz))); }
}

template<class E, int N>
struct alignas(sizeof(E) * N) Pack
{
    E v[N];
};

/*! persistent grid-stride kernel. VEC: V = 16/sizeof(T) particles per thread and iteration through 128-bit accesses
 *  (requires 16-byte aligned pointers), otherwise scalar. Keys equal to removeKey (2^(3 maxLevel)) are preserved. */
template<class K, class T, int KIND, bool VEC>
__global__ void __launch_bounds__(256) sfcKeysKernel(const T* __restrict__ x,
                                                     const T* __restrict__ y,
                                                     const T* __restrict__ z,
                                                     K* __restrict__ keys,
                                                     size_t n,
                                                     KeyParams<T> p,
                                                     const uint16_t* __restrict__ lutGlobal,
                                                     int lutEntries)
{
    extern __shared__ uint16_t lut[];
    if constexpr (KIND == 0)
    {
        // table entries are 2 B; copy as 4 B words
        const uint32_t* src = reinterpret_cast<const uint32_t*>(lutGlobal);
        uint32_t* dst       = reinterpret_cast<uint32_t*>(lut);
        for (int i = threadIdx.x; i < lutEntries / 2; i += blockDim.x)
            dst[i] = src[i];
        __syncthreads();
    }

    constexpr K removeKey = nodeRange<K>(0);
    size_t tid            = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t nthreads       = size_t(gridDim.x) * blockDim.x;

    if constexpr (VEC)
    {
        constexpr int V = 16 / sizeof(T);
        size_t nvec     = n / V;
        auto* xv        = reinterpret_cast<const Pack<T, V>*>(x);
        auto* yv        = reinterpret_cast<const Pack<T, V>*>(y);
        auto* zv        = reinterpret_cast<const Pack<T, V>*>(z);
        // keys of one particle vector: KP packs of KV keys (16 bytes each, or one 8-byte pack for 32-bit keys of
        // double coordinates)
        constexpr int KV = (16 / sizeof(K)) < V ? int(16 / sizeof(K)) : V;
        constexpr int KP = V / KV;
        auto* kv         = reinterpret_cast<Pack<K, KV>*>(keys);
        for (size_t i = tid; i < nvec; i += nthreads)
        {
            Pack<T, V> px = xv[i], py = yv[i], pz = zv[i];
            Pack<K, KV> pk[KP];
#pragma unroll
            for (int q = 0; q < KP; ++q)
                pk[q] = kv[i * KP + q];
#pragma unroll
            for (int j = 0; j < V; ++j)
            {
                K old = pk[j / KV].v[j % KV];
                K enc = encodeOne<K, T, KIND>(px.v[j], py.v[j], pz.v[j], p, lut);
                pk[j / KV].v[j % KV] = (old != removeKey) ? enc : old;
            }
#pragma unroll
            for (int q = 0; q < KP; ++q)
                kv[i * KP + q] = pk[q];
        }
        // tail
        for (size_t i = nvec * V + tid; i < n; i += nthreads)
        {
            if (keys[i] != removeKey) { keys[i] = encodeOne<K, T, KIND>(x[i], y[i], z[i], p, lut); }
        }
    }
    else
    {
        for (size_t i = tid; i < n; i += nthreads)
        {
            if (keys[i] != removeKey) { keys[i] = encodeOne<K, T, KIND>(x[i], y[i], z[i], p, lut); }
        }
    }
}

template<class K, class T>
int computeSfcKeys(int kind,
                   const T* x,
                   const T* y,
                   const T* z,
                   K* keys,
                   size_t n,
                   const double* lim,
                   const int* bnd,
                   cudaStream_t stream)
{
    CSB_REQUIRE(kind == 0 || kind == 1, "sfc kind must be 0 (Hilbert) or 1 (Morton)");
    if (n == 0) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);
    KeyParams<T> p;
    constexpr unsigned cubeLength = 1u << KeyTraits<K>::maxLevel;
    for (int d = 0; d < 3; ++d)
    {
        p.m[d]    = cubeLength * box.ilen[d];
        p.minm[d] = box.lim[2 * d] * p.m[d];
    }

    DeviceLut lut;
    if (int e = getDeviceLut(lut)) { return e; }
    int lutEntries = lut.numStates * 512;
    size_t smem    = kind == 0 ? size_t(lutEntries) * sizeof(uint16_t) : 0;

    int dev = 0, numSm = 0;
    CSB_CHECK(cudaGetDevice(&dev));
    CSB_CHECK(cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev));

    auto aligned = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; };
    bool vec     = aligned(x) && aligned(y) && aligned(z) && aligned(keys);

    constexpr int V   = 16 / sizeof(T);
    size_t workItems  = vec ? (n + V - 1) / V : n;
    unsigned maxGrid  = unsigned(numSm) * 8;
    unsigned grid     = unsigned(std::min<size_t>(maxGrid, (workItems + 255) / 256));
    grid              = grid ? grid : 1;

#define CSB_LAUNCH_KEYS(KIND, VEC)                                                                                     \
    sfcKeysKernel<K, T, KIND, VEC><<<grid, 256, smem, stream>>>(x, y, z, keys, n, p, lut.d_table, lutEntries)
    if (kind == 0)
    {
        if (vec) { CSB_LAUNCH_KEYS(0, true); }
        else { CSB_LAUNCH_KEYS(0, false); }
    }
    else
    {
        if (vec) { CSB_LAUNCH_KEYS(1, true); }
        else { CSB_LAUNCH_KEYS(1, false); }
    }
#undef CSB_LAUNCH_KEYS
    CSB_LAUNCH_CHECK();
    return 0;
}

} // namespace

int hilbertNumStates()
{
    DeviceLut lut;
    if (getDeviceLut(lut)) { return -1; }
    return lut.numStates;
}

} // namespace csb

extern "C"
{

int cs_compute_sfc_keys_u32f(int kind, const float* x, const float* y, const float* z, uint32_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream)
{
    return csb::computeSfcKeys<uint32_t, float>(kind, x, y, z, keys, n, lim, bnd, cudaStream_t(stream));
}

int cs_compute_sfc_keys_u32d(int kind, const double* x, const double* y, const double* z, uint32_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream)
{
    return csb::computeSfcKeys<uint32_t, double>(kind, x, y, z, keys, n, lim, bnd, cudaStream_t(stream));
}

int cs_compute_sfc_keys_u64f(int kind, const float* x, const float* y, const float* z, uint64_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream)
{
    return csb::computeSfcKeys<uint64_t, float>(kind, x, y, z, keys, n, lim, bnd, cudaStream_t(stream));
}

int cs_compute_sfc_keys_u64d(int kind, const double* x, const double* y, const double* z, uint64_t* keys, size_t n,
                             const double* lim, const int* bnd, void* stream)
{
    return csb::computeSfcKeys<uint64_t, double>(kind, x, y, z, keys, n, lim, bnd, cudaStream_t(stream));
}

} // extern "C"
