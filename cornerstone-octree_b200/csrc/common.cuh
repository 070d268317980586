/* cornerstone-b200: shared device/host helpers for the sm_100a kernels.
 *
 * Key algebra follows the reference's conventions (cstone/sfc/common.hpp): 32-bit keys hold 10 octal digits
 * (2 unused bits), 64-bit keys 21 digits (1 unused bit); node prefixes use the Warren-Salmon placeholder bit.
 * All floating-point expressions that feed bit-exact results are written with separate mul/add and the whole
 * library is compiled with --fmad=false so nothing contracts to FMA (SURVEY.md hazard H1).
 */
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace csb
{

using TreeNodeIndex = int;
using LocalIndex    = unsigned;

/* ------------------------------------------------------------------------------------------------ errors */

void setLastError(const std::string& msg);

#define CSB_CHECK(call)                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t err__ = (call);                                                                                    \
        if (err__ != cudaSuccess)                                                                                      \
        {                                                                                                              \
            ::csb::setLastError(std::string(#call) + " failed: " + cudaGetErrorString(err__) + " at " + __FILE__ +     \
                                ":" + std::to_string(__LINE__));                                                       \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

#define CSB_REQUIRE(cond, msg)                                                                                         \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            ::csb::setLastError(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + std::to_string(__LINE__));    \
            return 2;                                                                                                  \
        }                                                                                                              \
    } while (0)

void countLaunch();

//! call once after every kernel launch: counts it (cs_kernel_launch_count) and surfaces launch errors
#define CSB_LAUNCH_CHECK()                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        ::csb::countLaunch();                                                                                          \
        CSB_CHECK(cudaGetLastError());                                                                                 \
    } while (0)

//! library-owned scratch, one growing buffer per (device, stream, slot); nullptr (+ last error) on failure (lib.cu)
void* scratch(cudaStream_t s, int slot, size_t bytes);
enum ScratchSlot : int
{
    SCRATCH_A = 0,
    SCRATCH_B = 1,
    SCRATCH_C = 2,
    SCRATCH_D = 3,
    SCRATCH_E = 4
};
#define CSB_SCRATCH(ptr, type, stream, slot, bytes)                                                                    \
    type ptr = static_cast<type>(::csb::scratch(stream, slot, bytes));                                                 \
    if (!ptr) { return 1; }

//! experiment knobs (cs_tuning_set; not part of the drop-in surface).  Read once per entry-point call.
enum TuningKnob : int
{
    TUNE_NB_GROUPS = 1, // findNeighbors target groups: 0 (default) = full groups over runs of sibling leaves,
                        // 1 = leaf aligned
    TUNE_NB_SEARCH = 2, // findNeighbors: 0 = pick by leaf occupancy, 1 = per-lane walks, 2 = group-steered search
    TUNE_COUNT     = 16
};
int tuning(int knob);

inline unsigned iceil(size_t a, size_t b) { return unsigned((a + b - 1) / b); }

/* ------------------------------------------------------------------------------------------------ key traits */

template<class K>
struct KeyTraits;

template<>
struct KeyTraits<uint32_t>
{
    static constexpr int bits       = 32;
    static constexpr int maxLevel   = 10;
    static constexpr int unusedBits = 2;
};

template<>
struct KeyTraits<uint64_t>
{
    static constexpr int bits       = 64;
    static constexpr int maxLevel   = 21;
    static constexpr int unusedBits = 1;
};

template<class K>
__host__ __device__ constexpr K nodeRange(unsigned level)
{
    return K(1) << (3u * (KeyTraits<K>::maxLevel - level));
}

__host__ __device__ inline int clz(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __clz(int(v));
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

__host__ __device__ inline int clz(uint64_t v)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

//! level of a node spanning @p range keys (range is a power of 8)
template<class K>
__host__ __device__ inline unsigned treeLevel(K range)
{
    return (clz(K(range - 1)) - KeyTraits<K>::unusedBits) / 3;
}

template<class K>
__host__ __device__ inline int commonPrefix(K a, K b)
{
    return clz(K(a ^ b)) - KeyTraits<K>::unusedBits;
}

template<class K>
__host__ __device__ inline K encodePlaceholderBit(K code, int prefixLength)
{
    int nShifts = 3 * KeyTraits<K>::maxLevel - prefixLength;
    return (K(1) << prefixLength) | (code >> nShifts);
}

template<class K>
__host__ __device__ inline unsigned decodePrefixLength(K code)
{
    return KeyTraits<K>::bits - 1 - clz(code);
}

template<class K>
__host__ __device__ inline K decodePlaceholderBit(K code)
{
    int prefixLength = decodePrefixLength(code);
    K ret            = code ^ (K(1) << prefixLength);
    return ret << (3 * KeyTraits<K>::maxLevel - prefixLength);
}

template<class K>
__host__ __device__ inline unsigned octalDigit(K code, unsigned position)
{
    return unsigned((code >> (3u * (KeyTraits<K>::maxLevel - position))) & 7u);
}

__host__ __device__ inline int digitWeight(int digit)
{
    int fourGeqMask = -int(digit >= 4);
    return ((7 - digit) & fourGeqMask) - (digit & ~fourGeqMask);
}

//! first index in [0,n) with a[idx] >= v
template<class K, class I>
__host__ __device__ inline I lowerBound(const K* a, I n, K v)
{
    I lo = 0, hi = n;
    while (lo < hi)
    {
        I m = lo + (hi - lo) / 2;
        if (a[m] < v) { lo = m + 1; }
        else { hi = m; }
    }
    return lo;
}

//! first index in [0,n) with a[idx] > v
template<class K, class I>
__host__ __device__ inline I upperBound(const K* a, I n, K v)
{
    I lo = 0, hi = n;
    while (lo < hi)
    {
        I m = lo + (hi - lo) / 2;
        if (!(v < a[m])) { lo = m + 1; }
        else { hi = m; }
    }
    return lo;
}

/* ------------------------------------------------------------------------------------------------ box */

//! coordinate bounding box, same derived quantities as cstone::Box<T> (sfc/box.hpp:100-122), all stored in T
template<class T>
struct Box
{
    T lim[6];  // xmin xmax ymin ymax zmin zmax
    T len[3];  // max - min
    T ilen[3]; // T(1) / (max - min)
    int bnd[3]; // 0 open, 1 periodic, 2 fixed, 3 cubic_open

    __host__ __device__ bool pbc(int d) const { return bnd[d] == 1; }
};

template<class T>
inline Box<T> makeBox(const double* lim, const int* bnd)
{
    Box<T> b;
    for (int i = 0; i < 6; ++i)
        b.lim[i] = T(lim[i]);
    for (int d = 0; d < 3; ++d)
    {
        b.len[d]  = b.lim[2 * d + 1] - b.lim[2 * d];
        b.ilen[d] = T(1) / (b.lim[2 * d + 1] - b.lim[2 * d]);
        b.bnd[d]  = bnd[d];
    }
    return b;
}

/* ------------------------------------------------------------------------------------------------ fp helpers */

__device__ inline float rfloor(float v) { return floorf(v); }
__device__ inline double rfloor(double v) { return floor(v); }
__device__ inline float rrint(float v) { return rintf(v); }
__device__ inline double rrint(double v) { return rint(v); }
__device__ inline float rabs(float v) { return fabsf(v); }
__device__ inline double rabs(double v) { return fabs(v); }

//! d - pbc * L * rint(d * iL), the fold of sfc/box.hpp:176-190, evaluated left to right without contraction
template<class T>
__device__ inline T pbcFold(T d, int dim, const Box<T>& box)
{
    T f = T(box.pbc(dim) ? 1 : 0);
    return d - f * box.len[dim] * rrint(d * box.ilen[dim]);
}

/* ------------------------------------------------------------------------------------------------ scans */

//! exclusive scan of u32/i32 values on the stream; in may equal out. tmp must hold scanTempBytes(n)
size_t scanTempBytes(size_t n);
int exclusiveScanU32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, cudaStream_t s);

} // namespace csb
