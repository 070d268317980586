/* Internal octree construction ("linkOctree"), count upsweep and geometric node centres for sm_100a.
 * Replaces the reference's tree/octree_gpu.cu (createUnsortedLayout :41-64, linkTree :78-113, getLevelRange :116-124,
 * invertOrder :127-136, buildOctreeGpu :139-168, upsweepSumKernel :205-236) and
 * focus/source_center_gpu.cu:212-237 (computeGeoCentersKernel); arithmetic as in tree/octree.hpp:40-196,
 * sfc/hilbert.hpp:130-173,259-275 and sfc/box.hpp:318-335.
 */
#include <algorithm>

#include "common.cuh"
#include "hilbert.cuh"
#include "cstone_b200.h"

namespace csb
{

int sortByKeySkipU64(uint64_t*, uint32_t*, size_t, uint64_t*, uint32_t*, void*, size_t, cudaStream_t);
int sortByKeySkipU32(uint32_t*, uint32_t*, size_t, uint32_t*, uint32_t*, void*, size_t, cudaStream_t);
size_t sortTempBytesU64(size_t n);
size_t sortTempBytesU32(size_t n);

namespace
{

inline int sortByKeyK(uint64_t* k, uint32_t* v, size_t n, uint64_t* kb, uint32_t* vb, void* t, size_t tb, cudaStream_t s)
{
    return sortByKeySkipU64(k, v, n, kb, vb, t, tb, s);
}
inline int sortByKeyK(uint32_t* k, uint32_t* v, size_t n, uint32_t* kb, uint32_t* vb, void* t, size_t tb, cudaStream_t s)
{
    return sortByKeySkipU32(k, v, n, kb, vb, t, tb, s);
}
template<class K>
size_t sortTempBytesK(size_t n)
{
    if constexpr (sizeof(K) == 8) { return sortTempBytesU64(n); }
    else { return sortTempBytesU32(n); }
}

template<class K>
__device__ inline int binaryKeyWeight(K key, unsigned level)
{
    int ret = 0;
    for (unsigned l = 1; l <= level + 1; ++l)
        ret += digitWeight(int(octalDigit(key, l)));
    return ret;
}

template<class K>
__global__ void unsortedLayoutKernel(const K* __restrict__ leaves,
                                     int numInternal,
                                     int numLeaves,
                                     K* __restrict__ prefixes,
                                     int* __restrict__ internalToLeaf)
{
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= numLeaves) { return; }
    K key                             = leaves[tid];
    K next                            = leaves[tid + 1];
    unsigned level                    = treeLevel(K(next - key));
    prefixes[tid + numInternal]       = encodePlaceholderBit(key, 3 * int(level));
    internalToLeaf[tid + numInternal] = tid + numInternal;

    unsigned prefixLength = unsigned(commonPrefix(key, next));
    if (prefixLength % 3 == 0 && tid < numLeaves - 1)
    {
        int octIndex             = (tid + binaryKeyWeight(key, prefixLength / 3)) / 7;
        prefixes[octIndex]       = encodePlaceholderBit(key, int(prefixLength));
        internalToLeaf[octIndex] = octIndex;
    }
}

__global__ void invertOrderKernel(int* __restrict__ internalToLeaf, int* __restrict__ leafToInternal, int numNodes,
                                  int numInternal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }
    int v             = internalToLeaf[i];
    leafToInternal[v] = i;
    internalToLeaf[i] = v - numInternal;
}

template<class K>
__global__ void levelRangeKernel(const K* __restrict__ prefixes, int numNodes, int* __restrict__ levelRange)
{
    constexpr int maxLevel = KeyTraits<K>::maxLevel;
    int level              = threadIdx.x;
    if (level <= maxLevel)
    {
        levelRange[level] = lowerBound(prefixes, numNodes, encodePlaceholderBit(K(0), 3 * level));
    }
    else if (level == maxLevel + 1) { levelRange[level] = numNodes; }
}

template<class K>
__global__ void linkTreeKernel(const K* __restrict__ prefixes,
                               int numInternal,
                               const int* __restrict__ leafToInternal,
                               const int* __restrict__ levelRange,
                               int* __restrict__ childOffsets,
                               int* __restrict__ parents)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numInternal) { return; }
    int idxA              = leafToInternal[i];
    K prefix              = prefixes[idxA];
    K nodeKey             = decodePlaceholderBit(prefix);
    unsigned prefixLength = decodePrefixLength(prefix);
    unsigned level        = prefixLength / 3;
    K childPrefix         = encodePlaceholderBit(nodeKey, int(prefixLength) + 3);

    int s0       = levelRange[level + 1];
    int s1       = levelRange[level + 2];
    int childIdx = s0 + lowerBound(prefixes + s0, s1 - s0, childPrefix);
    if (childIdx != s1 && childPrefix == prefixes[childIdx])
    {
        childOffsets[idxA]          = childIdx;
        parents[(childIdx - 1) / 8] = idxA;
    }
}

__global__ void upsweepSumKernel(int first, int last, const int* __restrict__ childOffsets, uint32_t* counts)
{
    int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= last) { return; }
    int c = childOffsets[i];
    if (c)
    {
        uint64_t sum = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o)
            sum += counts[c + o];
        counts[i] = uint32_t(sum < 0xFFFFFFFFull ? sum : 0xFFFFFFFFull);
    }
}

/* ---------------------------------------------------------------- geometric centres */

template<class K, class T>
__global__ void geoCentersKernel(int kind, const K* __restrict__ prefixes, int numNodes, T* __restrict__ centers,
                                 T* __restrict__ sizes, Box<T> box)
{
    __shared__ unsigned char hilbertTables[hilbertTableBytes];
    if (kind == 0)
    {
        stageHilbertTables(hilbertTables);
        __syncthreads();
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }
    constexpr int maxCoord = 1 << KeyTraits<K>::maxLevel;
    constexpr T uL         = T(1) / maxCoord;

    K prefix            = prefixes[i];
    K startKey          = decodePlaceholderBit(prefix);
    unsigned level      = decodePrefixLength(prefix) / 3;
    unsigned cubeLength = unsigned(maxCoord) >> level;
    unsigned mask       = ~(cubeLength - 1);
    unsigned ix, iy, iz;
    if (kind == 0) { hilbertDecode(startKey, ix, iy, iz, hilbertTables); }
    else { decodeMorton(startKey, ix, iy, iz); }
    int imin[3] = {int(ix & mask), int(iy & mask), int(iz & mask)};
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        int imax           = imin[d] + int(cubeLength);
        T halfUnit         = T(0.5) * uL * box.len[d];
        centers[3 * i + d] = box.lim[2 * d] + T(imax + imin[d]) * halfUnit;
        sizes[3 * i + d]   = T(imax - imin[d]) * halfUnit;
    }
}

inline void* align256(void* p)
{
    return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
}

template<class K>
size_t buildOctreeTempBytes(int numLeaves)
{
    size_t numNodes = size_t(numLeaves) + size_t(numLeaves - 1) / 7;
    return (numNodes * sizeof(K) + 256) + (numNodes * 4 + 256) + sortTempBytesK<K>(numNodes) + 256;
}

} // namespace

template<class K>
int buildOctree(const K* leaves, int numLeaves, K* prefixes, int* childOffsets, int* parents, int* levelRange,
                int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes, cudaStream_t s)
{
    CSB_REQUIRE(numLeaves >= 1, "empty leaf array");
    CSB_REQUIRE(tmpBytes >= buildOctreeTempBytes<K>(numLeaves), "build_octree temp storage too small");
    int numInternal = (numLeaves - 1) / 7;
    int numNodes    = numLeaves + numInternal;

    K* keyBuf          = static_cast<K*>(align256(tmp));
    uint32_t* valueBuf = static_cast<uint32_t*>(align256(keyBuf + numNodes));
    void* sortTmp      = align256(valueBuf + numNodes);
    size_t sortBytes   = tmpBytes - size_t(static_cast<char*>(sortTmp) - static_cast<char*>(tmp));

    unsortedLayoutKernel<K><<<iceil(numLeaves, 256), 256, 0, s>>>(leaves, numInternal, numLeaves, prefixes,
                                                                   internalToLeaf);
    CSB_LAUNCH_CHECK();
    if (int e = sortByKeyK(prefixes, reinterpret_cast<uint32_t*>(internalToLeaf), size_t(numNodes), keyBuf, valueBuf,
                           sortTmp, sortBytes, s))
    {
        return e;
    }
    invertOrderKernel<<<iceil(numNodes, 256), 256, 0, s>>>(internalToLeaf, leafToInternal, numNodes, numInternal);
    CSB_LAUNCH_CHECK();
    levelRangeKernel<K><<<1, 32, 0, s>>>(prefixes, numNodes, levelRange);
    CSB_LAUNCH_CHECK();
    CSB_CHECK(cudaMemsetAsync(childOffsets, 0, size_t(numNodes) * sizeof(int), s));
    if (numInternal)
    {
        linkTreeKernel<K><<<iceil(numInternal, 256), 256, 0, s>>>(prefixes, numInternal, leafToInternal, levelRange,
                                                                   childOffsets, parents);
        CSB_LAUNCH_CHECK();
    }
    return 0;
}

template<class K, class T>
int computeGeoCenters(int kind, const K* prefixes, int numNodes, T* centers, T* sizes, const double* lim,
                      const int* bnd, cudaStream_t s)
{
    CSB_REQUIRE(kind == 0 || kind == 1, "sfc kind must be 0 (Hilbert) or 1 (Morton)");
    if (numNodes <= 0) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);
    geoCentersKernel<K, T><<<iceil(numNodes, 256), 256, 0, s>>>(kind, prefixes, numNodes, centers, sizes, box);
    CSB_LAUNCH_CHECK();
    return 0;
}

int upsweepSum(int maxLevel, const int* levelRangeHost, const int* childOffsets, uint32_t* counts, cudaStream_t s)
{
    for (int level = maxLevel; level >= 0; --level)
    {
        int first = levelRangeHost[level], last = levelRangeHost[level + 1];
        if (last > first)
        {
            upsweepSumKernel<<<iceil(last - first, 128), 128, 0, s>>>(first, last, childOffsets, counts);
            CSB_LAUNCH_CHECK();
        }
    }
    return 0;
}

template int buildOctree<uint32_t>(const uint32_t*, int, uint32_t*, int*, int*, int*, int*, int*, void*, size_t,
                                   cudaStream_t);
template int buildOctree<uint64_t>(const uint64_t*, int, uint64_t*, int*, int*, int*, int*, int*, void*, size_t,
                                   cudaStream_t);
template int computeGeoCenters<uint32_t, float>(int, const uint32_t*, int, float*, float*, const double*, const int*,
                                                cudaStream_t);
template int computeGeoCenters<uint64_t, float>(int, const uint64_t*, int, float*, float*, const double*, const int*,
                                                cudaStream_t);
template int computeGeoCenters<uint64_t, double>(int, const uint64_t*, int, double*, double*, const double*,
                                                 const int*, cudaStream_t);

size_t buildOctreeTempBytesU32(int numLeaves) { return buildOctreeTempBytes<uint32_t>(numLeaves); }
size_t buildOctreeTempBytesU64(int numLeaves) { return buildOctreeTempBytes<uint64_t>(numLeaves); }

} // namespace csb

extern "C"
{

size_t cs_build_octree_temp_bytes_u32(int numLeaves) { return csb::buildOctreeTempBytesU32(numLeaves); }
size_t cs_build_octree_temp_bytes_u64(int numLeaves) { return csb::buildOctreeTempBytesU64(numLeaves); }

int cs_build_octree_u32(const uint32_t* leaves, int numLeaves, uint32_t* prefixes, int* childOffsets, int* parents,
                        int* levelRange, int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes,
                        void* stream)
{
    return csb::buildOctree<uint32_t>(leaves, numLeaves, prefixes, childOffsets, parents, levelRange, internalToLeaf,
                                      leafToInternal, tmp, tmpBytes, cudaStream_t(stream));
}
int cs_build_octree_u64(const uint64_t* leaves, int numLeaves, uint64_t* prefixes, int* childOffsets, int* parents,
                        int* levelRange, int* internalToLeaf, int* leafToInternal, void* tmp, size_t tmpBytes,
                        void* stream)
{
    return csb::buildOctree<uint64_t>(leaves, numLeaves, prefixes, childOffsets, parents, levelRange, internalToLeaf,
                                      leafToInternal, tmp, tmpBytes, cudaStream_t(stream));
}

int cs_upsweep_sum(int maxLevel, const int* levelRangeHost, const int* childOffsets, uint32_t* counts, void* stream)
{
    return csb::upsweepSum(maxLevel, levelRangeHost, childOffsets, counts, cudaStream_t(stream));
}

int cs_compute_geo_centers_u32f(int kind, const uint32_t* prefixes, int numNodes, float* centers, float* sizes,
                                const double* lim, const int* bnd, void* stream)
{
    return csb::computeGeoCenters<uint32_t, float>(kind, prefixes, numNodes, centers, sizes, lim, bnd,
                                                   cudaStream_t(stream));
}
int cs_compute_geo_centers_u64f(int kind, const uint64_t* prefixes, int numNodes, float* centers, float* sizes,
                                const double* lim, const int* bnd, void* stream)
{
    return csb::computeGeoCenters<uint64_t, float>(kind, prefixes, numNodes, centers, sizes, lim, bnd,
                                                   cudaStream_t(stream));
}
int cs_compute_geo_centers_u64d(int kind, const uint64_t* prefixes, int numNodes, double* centers, double* sizes,
                                const double* lim, const int* bnd, void* stream)
{
    return csb::computeGeoCenters<uint64_t, double>(kind, prefixes, numNodes, centers, sizes, lim, bnd,
                                                    cudaStream_t(stream));
}

} // extern "C"
