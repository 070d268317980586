/* Focus-tree (locally essential tree) maintenance kernels for sm_100a.
 * Replaces the reference's focus/rebalance_gpu.cu (rebalanceDecisionEssentialKernel :29-60, protectAncestorsKernel
 * :113-160, enforceKeysKernel :166-200) and tree/csarray_gpu.cu:238-270 (countSfcGaps/fillSfcGaps used by
 * focus/inject.hpp:50-84), with the arithmetic of focus/rebalance.hpp:31-252 and sfc/common.hpp:360-470.
 * Integer-only: bit-exact.
 */
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "cstone_b200.h"
#include "focus.cuh"

namespace csb
{

namespace
{

template<class K>
__host__ __device__ inline int lastNzPlace(K x)
{
    constexpr int maxLevel = KeyTraits<K>::maxLevel;
    if (x == 0) { return maxLevel; }
    int ctz = 0;
    while (((x >> ctz) & K(1)) == 0)
        ++ctz;
    return maxLevel - ctz / 3;
}

template<class K>
__host__ __device__ inline K makePrefix(K a)
{
    if (a == 0) { return 1; }
    return encodePlaceholderBit(a, 3 * lastNzPlace(a));
}

template<class K>
__host__ __device__ inline K octalPower(int pos)
{
    return K(1) << (3 * (KeyTraits<K>::maxLevel - pos));
}

//! smallest node that contains nodeKey (tree/octree.hpp:198-217)
template<class K>
__device__ inline int containingNode(K nodeKey, const K* __restrict__ prefixes, const int* __restrict__ childOffsets)
{
    int nodeLevel = int(decodePrefixLength(nodeKey) / 3);
    K key         = decodePlaceholderBit(nodeKey);
    int ret       = 0;
    for (int i = 1; i <= nodeLevel; ++i)
    {
        if (childOffsets[ret] == 0 || nodeKey == prefixes[ret]) { break; }
        ret = childOffsets[ret] + int(octalDigit(key, unsigned(i)));
    }
    return ret;
}

template<class K>
__global__ void essentialOpsKernel(const K* __restrict__ prefixes,
                                   const int* __restrict__ childOffsets,
                                   const int* __restrict__ parents,
                                   const uint32_t* __restrict__ counts,
                                   const uint8_t* __restrict__ macs,
                                   K focusStart,
                                   K focusEnd,
                                   uint32_t bucketSize,
                                   int* __restrict__ nodeOps,
                                   int numNodes)
{
    constexpr unsigned maxLevel = KeyTraits<K>::maxLevel;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }

    K nodeKey      = prefixes[i];
    unsigned level = decodePrefixLength(nodeKey) / 3;
    int op         = 1;
    bool merge     = false;
    if (i)
    {
        int parent      = parents[(i - 1) / 8];
        bool countMerge = counts[parent] <= bucketSize;
        bool macMerge   = macs[parent] == 0;
        K firstGroupKey = decodePlaceholderBit(prefixes[parent]);
        K lastGroupKey  = firstGroupKey + 8 * nodeRange<K>(level);
        bool inFringe   = lastGroupKey > focusStart && focusEnd > firstGroupKey; // overlapTwoRanges
        merge           = countMerge || (macMerge && !inFringe);
    }
    if (merge) { op = 0; }
    else
    {
        K nodeStart  = decodePlaceholderBit(nodeKey);
        bool isLeaf  = childOffsets[i] == 0;
        bool inFocus = nodeStart >= focusStart && nodeStart < focusEnd;
        uint32_t c   = counts[i];
        if (isLeaf && (macs[i] || inFocus))
        {
            if (level + 3 < maxLevel && c > 4096u * bucketSize) { op = 4096; }
            else if (level + 2 < maxLevel && c > 512u * bucketSize) { op = 512; }
            else if (level + 1 < maxLevel && c > 64u * bucketSize) { op = 64; }
            else if (level < maxLevel && c > bucketSize) { op = 8; }
        }
    }
    nodeOps[i] = op;
}

/*! focus/rebalance.hpp:183-252, one thread per mandatory key.  The reference loops over the keys serially; the
 *  operations commute (merges are only ever cancelled 0 -> 1, splits raised with max), so atomics give the same
 *  nodeOps and the same maximum status. */
template<class K>
__global__ void enforceKeysKernel(const K* __restrict__ keys,
                                  int numKeys,
                                  const K* __restrict__ prefixes,
                                  const int* __restrict__ childOffsets,
                                  const int* __restrict__ parents,
                                  int* nodeOps,
                                  int* status)
{
    constexpr int maxLevel = KeyTraits<K>::maxLevel;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numKeys) { return; }
    K key = keys[t];
    if (key == 0 || key == nodeRange<K>(0)) { return; }

    int st            = ENFORCE_CONVERGED;
    K nodeKeyWant     = makePrefix(key);
    int nodeIdx       = containingNode(nodeKeyWant, prefixes, childOffsets);
    K nodeKeyHave     = prefixes[nodeIdx];
    int nodeLevelHave = int(decodePrefixLength(nodeKeyHave) / 3);

    // nodeOps only ever move 0 -> 1 (cancelled merge) or up (atomicMax), so a plain read that already shows the wanted
    // state is final; the atomics are only issued when the read says work is left.  With millions of mandatory keys
    // (the global leaves) mapping to a few coarse focus nodes this removes nearly all of the contended atomics.
    volatile int* ops = nodeOps;
    bool trySplit     = nodeKeyHave != nodeKeyWant && nodeLevelHave < maxLevel;
    bool undoMerges   = ops[nodeIdx] == 0 || trySplit;
    if (undoMerges && nodeIdx > 0)
    {
        st         = ENFORCE_CANCEL_MERGE;
        int parent = nodeIdx;
        do
        {
            parent           = parents[(parent - 1) / 8];
            int firstSibling = childOffsets[parent];
            for (int i = firstSibling; i < firstSibling + 8; ++i)
                if (ops[i] == 0) { atomicCAS(&nodeOps[i], 0, 1); }
        } while (parent != 0);
    }
    if (trySplit)
    {
        int keyPos    = lastNzPlace(key);
        int levelDiff = keyPos - nodeLevelHave;
        st            = levelDiff > 1 ? ENFORCE_FAILED : ENFORCE_REBALANCE;
        levelDiff     = levelDiff < 1 ? levelDiff : 1;
        int want      = 1 << (3 * levelDiff);
        if (ops[nodeIdx] < want) { atomicMax(&nodeOps[nodeIdx], want); }
    }
    if (st != ENFORCE_CONVERGED && *(volatile int*)status < st) { atomicMax(status, st); }
}

//! focus/rebalance.hpp:91-116,156-169; in-place with the reference's benign race
template<class K>
__global__ void protectAncestorsKernel(const K* __restrict__ prefixes,
                                       const int* __restrict__ parents,
                                       int* nodeOps,
                                       int numNodes,
                                       int* numChanges)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }
    volatile int* ops = nodeOps;
    int decision;
    if (i == 0) { decision = ops[0]; }
    else
    {
        int a = i;
        while (ops[a] == 0)
            a = parents[(a - 1) / 8];
        decision = decodePlaceholderBit(prefixes[i]) == decodePlaceholderBit(prefixes[a]) ? ops[a] : 0;
    }
    if (decision != 1) { *numChanges = 1; }
    ops[i] = decision;
}

__global__ void gatherLeafOpsKernel(const int* __restrict__ leafToInternalLeaves,
                                    int numLeaves,
                                    const int* __restrict__ nodeOpsAll,
                                    int* __restrict__ leafOps,
                                    int* notAllOne)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numLeaves) { return; }
    if (i == numLeaves)
    {
        leafOps[i] = 0;
        return;
    }
    int op     = nodeOpsAll[leafToInternalLeaves[i]];
    leafOps[i] = op;
    if (op != 1) { *notAllOne = 1; }
}

/*! number of octree nodes needed to span [a,b) (sfc/common.hpp:409-470 spanSfcRange): walk up from a while its
 *  trailing digits can be completed to the next coarser node, then walk down to b */
template<class K>
__host__ __device__ inline int spanSfcRangeImpl(K a, K b, K* output)
{
    constexpr int unusedBits = KeyTraits<K>::unusedBits;
    if (a == b) { return 0; }
    int numValues      = 0;
    int firstDiffPos   = (clz(K(a ^ b)) + 3 - unusedBits) / 3;
    int aLastNzPos     = lastNzPlace(a);
    int bLastNzPos     = lastNzPlace(b);
    for (int pos = aLastNzPos; pos > firstDiffPos; --pos)
    {
        int numDigits = (8 - int(octalDigit(a, unsigned(pos)))) % 8;
        numValues += numDigits;
        while (numDigits--)
        {
            if (output) { *output++ = a; }
            a += octalPower<K>(pos);
        }
    }
    for (int pos = firstDiffPos; pos <= bLastNzPos; ++pos)
    {
        int numDigits = int(octalDigit(b, unsigned(pos))) - int(octalDigit(a, unsigned(pos)));
        numValues += numDigits;
        while (numDigits--)
        {
            if (output) { *output++ = a; }
            a += octalPower<K>(pos);
        }
    }
    return numValues;
}

template<class K>
__global__ void countGapsKernel(const K* __restrict__ keys, int numGaps, uint32_t* __restrict__ gapCounts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numGaps) { return; }
    gapCounts[i] = i < numGaps ? uint32_t(spanSfcRangeImpl<K>(keys[i], keys[i + 1], nullptr)) : 0u;
}

template<class K>
__global__ void fillGapsKernel(const K* __restrict__ keys, int numGaps, const uint32_t* __restrict__ offsets,
                               K* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numGaps) { return; }
    spanSfcRangeImpl<K>(keys[i], keys[i + 1], out + offsets[i]);
    if (i == numGaps - 1) { out[offsets[numGaps]] = keys[numGaps]; }
}

__global__ void scatterCountsKernel(const int* __restrict__ leafToInternalLeaves, int numLeaves,
                                    const uint32_t* __restrict__ leafCounts, uint32_t* __restrict__ nodeCounts)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < numLeaves) { nodeCounts[leafToInternalLeaves[i]] = leafCounts[i]; }
}

template<class T>
__global__ void gatherVec3Kernel(const int* __restrict__ map, int n, const T* __restrict__ src, T* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) { return; }
    int s          = map[i];
    dst[3 * i]     = src[3 * s];
    dst[3 * i + 1] = src[3 * s + 1];
    dst[3 * i + 2] = src[3 * s + 2];
}

__global__ void layoutCountsKernel(const uint32_t* __restrict__ leafCounts, const uint8_t* __restrict__ flags,
                                   const int* __restrict__ leafToInternalLeaves, int numLeaves, int ownStart,
                                   int ownEnd, uint32_t* __restrict__ layout)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numLeaves) { return; }
    if (i == numLeaves)
    {
        layout[i] = 0;
        return;
    }
    bool have = (ownStart <= i && i < ownEnd) || flags[leafToInternalLeaves[i]];
    layout[i] = have ? leafCounts[i] : 0u;
}

template<class T>
__global__ void minMaxPartialKernel(const T* __restrict__ a, size_t n, T* __restrict__ partial /* [2*gridDim] */)
{
    __shared__ T smin[32], smax[32];
    size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t nt  = size_t(gridDim.x) * blockDim.x;
    T mn = a[0], mx = a[0];
    for (size_t i = tid; i < n; i += nt)
    {
        T v = a[i];
        mn  = v < mn ? v : mn;
        mx  = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        T a1 = __shfl_xor_sync(0xffffffffu, mn, o);
        T b1 = __shfl_xor_sync(0xffffffffu, mx, o);
        mn   = a1 < mn ? a1 : mn;
        mx   = b1 > mx ? b1 : mx;
    }
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        smin[warp] = mn;
        smax[warp] = mx;
    }
    __syncthreads();
    if (warp == 0)
    {
        unsigned nw = blockDim.x >> 5;
        mn          = lane < nw ? smin[lane] : smin[0];
        mx          = lane < nw ? smax[lane] : smax[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            T a1 = __shfl_xor_sync(0xffffffffu, mn, o);
            T b1 = __shfl_xor_sync(0xffffffffu, mx, o);
            mn   = a1 < mn ? a1 : mn;
            mx   = b1 > mx ? b1 : mx;
        }
        if (lane == 0)
        {
            partial[2 * blockIdx.x]     = mn;
            partial[2 * blockIdx.x + 1] = mx;
        }
    }
}

__global__ void maxU32Kernel(const uint32_t* __restrict__ a, size_t n, uint32_t* result)
{
    size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t nt  = size_t(gridDim.x) * blockDim.x;
    uint32_t m = 0;
    for (size_t i = tid; i < n; i += nt)
        m = max(m, a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) { atomicMax(result, m); }
}

__global__ void maxIntoKernel(uint32_t* __restrict__ a, const uint32_t* __restrict__ b, size_t n)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = max(a[i], b[i]); }
}

template<class K>
__global__ void lowerBoundsKernel(const K* __restrict__ keys, size_t n, const K* __restrict__ targets, int numTargets,
                                  uint32_t* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < numTargets) { out[i] = uint32_t(lowerBound(keys, n, targets[i])); }
}

} // namespace

template<class K>
int essentialOps(const K* prefixes, const int* childOffsets, const int* parents, const uint32_t* counts,
                 const uint8_t* macs, K focusStart, K focusEnd, uint32_t bucketSize, int* nodeOps, int numNodes,
                 cudaStream_t s)
{
    essentialOpsKernel<K><<<iceil(numNodes, 256), 256, 0, s>>>(prefixes, childOffsets, parents, counts, macs,
                                                               focusStart, focusEnd, bucketSize, nodeOps, numNodes);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int enforceKeys(const K* keys, int numKeys, const K* prefixes, const int* childOffsets, const int* parents,
                int* nodeOps, int* statusDev, cudaStream_t s)
{
    if (numKeys == 0) { return 0; }
    enforceKeysKernel<K><<<iceil(numKeys, 128), 128, 0, s>>>(keys, numKeys, prefixes, childOffsets, parents, nodeOps,
                                                             statusDev);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int protectAncestors(const K* prefixes, const int* parents, int* nodeOps, int numNodes, int* changesDev, cudaStream_t s)
{
    protectAncestorsKernel<K><<<iceil(numNodes, 256), 256, 0, s>>>(prefixes, parents, nodeOps, numNodes, changesDev);
    CSB_LAUNCH_CHECK();
    return 0;
}

int gatherLeafOps(const int* leafToInternalLeaves, int numLeaves, const int* nodeOpsAll, int* leafOps,
                  int* notAllOneDev, cudaStream_t s)
{
    gatherLeafOpsKernel<<<iceil(numLeaves + 1, 256), 256, 0, s>>>(leafToInternalLeaves, numLeaves, nodeOpsAll, leafOps,
                                                                  notAllOneDev);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int countGaps(const K* keys, int numGaps, uint32_t* gapCounts, cudaStream_t s)
{
    countGapsKernel<K><<<iceil(numGaps + 1, 256), 256, 0, s>>>(keys, numGaps, gapCounts);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int fillGaps(const K* keys, int numGaps, const uint32_t* offsets, K* out, cudaStream_t s)
{
    fillGapsKernel<K><<<iceil(numGaps, 256), 256, 0, s>>>(keys, numGaps, offsets, out);
    CSB_LAUNCH_CHECK();
    return 0;
}

int scatterCounts(const int* leafToInternalLeaves, int numLeaves, const uint32_t* leafCounts, uint32_t* nodeCounts,
                  cudaStream_t s)
{
    scatterCountsKernel<<<iceil(numLeaves, 256), 256, 0, s>>>(leafToInternalLeaves, numLeaves, leafCounts, nodeCounts);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class T>
int gatherVec3(const int* map, int n, const T* src, T* dst, cudaStream_t s)
{
    if (n == 0) { return 0; }
    gatherVec3Kernel<T><<<iceil(n, 256), 256, 0, s>>>(map, n, src, dst);
    CSB_LAUNCH_CHECK();
    return 0;
}

int layoutCounts(const uint32_t* leafCounts, const uint8_t* flags, const int* leafToInternalLeaves, int numLeaves,
                 int ownStart, int ownEnd, uint32_t* layout, cudaStream_t s)
{
    layoutCountsKernel<<<iceil(numLeaves + 1, 256), 256, 0, s>>>(leafCounts, flags, leafToInternalLeaves, numLeaves,
                                                                 ownStart, ownEnd, layout);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class T>
int minMaxPartials(const T* a, size_t n, T* partial, int numBlocks, cudaStream_t s)
{
    minMaxPartialKernel<T><<<numBlocks, 512, 0, s>>>(a, n, partial);
    CSB_LAUNCH_CHECK();
    return 0;
}

int maxU32(const uint32_t* a, size_t n, uint32_t* resultDev, cudaStream_t s)
{
    CSB_CHECK(cudaMemsetAsync(resultDev, 0, sizeof(uint32_t), s));
    if (n == 0) { return 0; }
    unsigned grid = unsigned(std::min<size_t>(592, (n + 255) / 256));
    maxU32Kernel<<<grid, 256, 0, s>>>(a, n, resultDev);
    CSB_LAUNCH_CHECK();
    return 0;
}

int maxInto(uint32_t* a, const uint32_t* b, size_t n, cudaStream_t s)
{
    if (n == 0) { return 0; }
    maxIntoKernel<<<iceil(n, 256), 256, 0, s>>>(a, b, n);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int lowerBounds(const K* keys, size_t n, const K* targets, int numTargets, uint32_t* out, cudaStream_t s)
{
    lowerBoundsKernel<K><<<iceil(numTargets, 64), 64, 0, s>>>(keys, n, targets, numTargets, out);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int spanSfcRangeHost(K a, K b, K* output)
{
    return spanSfcRangeImpl<K>(a, b, output);
}

#define CSB_INST_K(K)                                                                                                  \
    template int essentialOps<K>(const K*, const int*, const int*, const uint32_t*, const uint8_t*, K, K, uint32_t,    \
                                 int*, int, cudaStream_t);                                                             \
    template int enforceKeys<K>(const K*, int, const K*, const int*, const int*, int*, int*, cudaStream_t);           \
    template int protectAncestors<K>(const K*, const int*, int*, int, int*, cudaStream_t);                             \
    template int countGaps<K>(const K*, int, uint32_t*, cudaStream_t);                                                 \
    template int fillGaps<K>(const K*, int, const uint32_t*, K*, cudaStream_t);                                        \
    template int lowerBounds<K>(const K*, size_t, const K*, int, uint32_t*, cudaStream_t);                             \
    template int spanSfcRangeHost<K>(K, K, K*);
CSB_INST_K(uint32_t)
CSB_INST_K(uint64_t)
#undef CSB_INST_K

template int gatherVec3<float>(const int*, int, const float*, float*, cudaStream_t);
template int gatherVec3<double>(const int*, int, const double*, double*, cudaStream_t);
template int minMaxPartials<float>(const float*, size_t, float*, int, cudaStream_t);
template int minMaxPartials<double>(const double*, size_t, double*, int, cudaStream_t);
template int minMaxPartials<uint32_t>(const uint32_t*, size_t, uint32_t*, int, cudaStream_t);

//! minMax (primitives/primitives_gpu.h:72-73): per-block partial results, folded on the host
template<class T>
int minMaxHost(const T* first, size_t n, T* minOut, T* maxOut, cudaStream_t s)
{
    CSB_REQUIRE(n > 0, "minMax of an empty range");
    const int blocks = int(std::min<size_t>(592, (n + 511) / 512));
    CSB_SCRATCH(partial, T*, s, SCRATCH_A, size_t(2) * blocks * sizeof(T));
    if (int e = minMaxPartials<T>(first, n, partial, blocks, s)) { return e; }
    std::vector<T> host(size_t(2) * blocks);
    CSB_CHECK(cudaMemcpyAsync(host.data(), partial, host.size() * sizeof(T), cudaMemcpyDeviceToHost, s));
    CSB_CHECK(cudaStreamSynchronize(s));
    T mn = host[0], mx = host[1];
    for (int b = 1; b < blocks; ++b)
    {
        mn = std::min(mn, host[2 * b]);
        mx = std::max(mx, host[2 * b + 1]);
    }
    *minOut = mn;
    *maxOut = mx;
    return 0;
}

} // namespace csb

extern "C"
{

/* countSfcGapsGpu / fillSfcGapsGpu (tree/csarray_gpu.h:78-82, csarray_gpu.cu:238-270): number of octree nodes that span
 * each gap [tree[i], tree[i+1]) of a sorted key sequence, and the nodes themselves at the scanned offsets */
int cs_count_sfc_gaps_u32(const uint32_t* tree, int numNodes, int* nodeOps, void* stream)
{
    return csb::countGaps<uint32_t>(tree, numNodes, reinterpret_cast<uint32_t*>(nodeOps), cudaStream_t(stream));
}
int cs_count_sfc_gaps_u64(const uint64_t* tree, int numNodes, int* nodeOps, void* stream)
{
    return csb::countGaps<uint64_t>(tree, numNodes, reinterpret_cast<uint32_t*>(nodeOps), cudaStream_t(stream));
}
int cs_fill_sfc_gaps_u32(const uint32_t* tree, int numNodes, const int* nodeOps, uint32_t* newTree, void* stream)
{
    return csb::fillGaps<uint32_t>(tree, numNodes, reinterpret_cast<const uint32_t*>(nodeOps), newTree,
                                   cudaStream_t(stream));
}
int cs_fill_sfc_gaps_u64(const uint64_t* tree, int numNodes, const int* nodeOps, uint64_t* newTree, void* stream)
{
    return csb::fillGaps<uint64_t>(tree, numNodes, reinterpret_cast<const uint32_t*>(nodeOps), newTree,
                                   cudaStream_t(stream));
}

/* rebalanceDecisionEssentialGpu / protectAncestorsGpu / enforceKeysGpu (focus/rebalance_gpu.h:27-79); the host-value
 * results the reference returns (converged flag, ResolutionStatus) come back through the last pointer argument, which
 * synchronises the stream exactly as the reference functions do */
#define CSB_FOCUS_ABI(SFX, K)                                                                                          \
    int cs_rebalance_decision_essential_##SFX(const K* prefixes, const int* childOffsets, const int* parents,         \
                                              const uint32_t* counts, const uint8_t* macs, K focusStart, K focusEnd,  \
                                              uint32_t bucketSize, int* nodeOps, int numNodes, void* stream)          \
    {                                                                                                                  \
        return csb::essentialOps<K>(prefixes, childOffsets, parents, counts, macs, focusStart, focusEnd, bucketSize,  \
                                    nodeOps, numNodes, cudaStream_t(stream));                                          \
    }                                                                                                                  \
    int cs_protect_ancestors_##SFX(const K* prefixes, const int* parents, int* nodeOps, int numNodes, int* converged, \
                                   void* stream)                                                                       \
    {                                                                                                                  \
        cudaStream_t s = cudaStream_t(stream);                                                                         \
        CSB_SCRATCH(flag, int*, s, csb::SCRATCH_A, sizeof(int));                                                       \
        CSB_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));                                                           \
        if (int e = csb::protectAncestors<K>(prefixes, parents, nodeOps, numNodes, flag, s)) { return e; }            \
        int changes = 0;                                                                                               \
        CSB_CHECK(cudaMemcpyAsync(&changes, flag, sizeof(int), cudaMemcpyDeviceToHost, s));                            \
        CSB_CHECK(cudaStreamSynchronize(s));                                                                           \
        *converged = changes == 0;                                                                                     \
        return 0;                                                                                                      \
    }                                                                                                                  \
    int cs_enforce_keys_##SFX(const K* keys, int numKeys, const K* prefixes, const int* childOffsets,                 \
                              const int* parents, int* nodeOps, int* status, void* stream)                            \
    {                                                                                                                  \
        cudaStream_t s = cudaStream_t(stream);                                                                         \
        CSB_SCRATCH(flag, int*, s, csb::SCRATCH_A, sizeof(int));                                                       \
        CSB_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));                                                           \
        if (int e = csb::enforceKeys<K>(keys, numKeys, prefixes, childOffsets, parents, nodeOps, flag, s)) { return e; } \
        CSB_CHECK(cudaMemcpyAsync(status, flag, sizeof(int), cudaMemcpyDeviceToHost, s));                              \
        CSB_CHECK(cudaStreamSynchronize(s));                                                                           \
        return 0;                                                                                                      \
    }
CSB_FOCUS_ABI(u32, uint32_t)
CSB_FOCUS_ABI(u64, uint64_t)
#undef CSB_FOCUS_ABI

/* minMax (primitives/primitives_gpu.h:72-73): smallest and largest element of a device array, returned on the host
 * (synchronises the stream like the reference) */
int cs_min_max_f(const float* first, size_t n, float* minOut, float* maxOut, void* stream)
{
    return csb::minMaxHost<float>(first, n, minOut, maxOut, cudaStream_t(stream));
}
int cs_min_max_d(const double* first, size_t n, double* minOut, double* maxOut, void* stream)
{
    return csb::minMaxHost<double>(first, n, minOut, maxOut, cudaStream_t(stream));
}
int cs_min_max_u32(const uint32_t* first, size_t n, uint32_t* minOut, uint32_t* maxOut, void* stream)
{
    return csb::minMaxHost<uint32_t>(first, n, minOut, maxOut, cudaStream_t(stream));
}

} // extern "C"
