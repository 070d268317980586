/* Kernels of the locally essential tree (LET) of a multi-rank domain:
 *   - min-distance MAC spheres and the MAC marking traversal (traversal/macs.hpp:40-57,118-260,
 *     traversal/collisions_gpu.cu:91-140, focus/octree_focus_mpi.hpp:422-499)
 *   - node counts of LET leaves taken from the replicated global tree (focus/rebalance.hpp:263-284)
 *   - index gathers/scatters of the peer count exchange (focus/exchange_focus.hpp:310-366)
 *   - range packing of the halo exchange (halos/gather_halos_gpu.cu:27-58)
 */
#include "common.cuh"
#include "focus.cuh"
#include "hilbert.cuh"

namespace csb
{

namespace
{

//! centres (x, y, z, mac^2) for the minimum distance MAC: computeMinMacR2 (traversal/macs.hpp:40-57)
template<class T>
__global__ void minMacKernel(const T* __restrict__ geoCenters, const T* __restrict__ geoSizes, int numNodes,
                             float invThetaEff, T* __restrict__ centers4)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }
    T sx = geoSizes[3 * i], sy = geoSizes[3 * i + 1], sz = geoSizes[3 * i + 2];
    T mx = sx > sy ? sx : sy;
    mx   = mx > sz ? mx : sz;
    T l   = T(2) * mx;
    T mac = l * invThetaEff;
    centers4[4 * i]     = geoCenters[3 * i];
    centers4[4 * i + 1] = geoCenters[3 * i + 1];
    centers4[4 * i + 2] = geoCenters[3 * i + 2];
    centers4[4 * i + 3] = mac * mac;
}

//! containedIn(codeStart, codeEnd, IBox) of traversal/boxoverlap.hpp:96-117 (Hilbert keys)
template<class K>
__device__ inline bool iboxContainedIn(K codeStart, K codeEnd, const int* lo, const int* hi)
{
    constexpr int pbcRange = 1 << KeyTraits<K>::maxLevel;
    int mn = min(min(lo[0], lo[1]), lo[2]);
    int mx = max(max(hi[0], hi[1]), hi[2]);
    if (mn < 0 || mx > pbcRange) { return codeStart == 0 && codeEnd == nodeRange<K>(0); }
    K lowCode      = iHilbertLoop<K>(unsigned(lo[0]), unsigned(lo[1]), unsigned(lo[2]));
    K highCode     = iHilbertLoop<K>(unsigned(hi[0] - 1), unsigned(hi[1] - 1), unsigned(hi[2] - 1));
    unsigned level = unsigned(commonPrefix(lowCode, highCode)) / 3;
    K nodeStart    = lowCode & ~(nodeRange<K>(level) - 1);
    K nodeEnd      = nodeStart + nodeRange<K>(level);
    return nodeStart >= codeStart && nodeEnd <= codeEnd;
}

/*! markMacs (traversal/macs.hpp:149-260): one thread per focus leaf whose extended box is not interior to the focus;
 *  marks every LET node outside the focus that fails the MAC against that leaf.  Stores race benignly (all write 1). */
template<class K, class T>
__global__ void __launch_bounds__(128) markMacsKernel(const K* __restrict__ prefixes,
                                                      const int* __restrict__ childOffsets,
                                                      const int* __restrict__ parents,
                                                      const T* __restrict__ centers4,
                                                      Box<T> box,
                                                      const K* __restrict__ focusNodes,
                                                      int numFocusNodes,
                                                      uint8_t* markings)
{
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= numFocusNodes) { return; }
    constexpr int maxCoord = 1 << KeyTraits<K>::maxLevel;
    constexpr T uL         = T(1) / maxCoord;

    K focusStart = focusNodes[0];
    K focusEnd   = focusNodes[numFocusNodes];
    K a = focusNodes[tid], b = focusNodes[tid + 1];

    unsigned level      = treeLevel<K>(b - a);
    unsigned cubeLength = unsigned(maxCoord) >> level;
    unsigned mask       = ~(cubeLength - 1);
    unsigned ix, iy, iz;
    decodeHilbert(a, ix, iy, iz);
    int lo[3] = {int(ix & mask), int(iy & mask), int(iz & mask)};
    int hi[3] = {lo[0] + int(cubeLength), lo[1] + int(cubeLength), lo[2] + int(cubeLength)};
    int elo[3] = {lo[0] - 1, lo[1] - 1, lo[2] - 1};
    int ehi[3] = {hi[0] + 1, hi[1] + 1, hi[2] + 1};
    if (iboxContainedIn<K>(focusStart, focusEnd, elo, ehi)) { return; }

    T tc[3], ts[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        T halfUnit = T(0.5) * uL * box.len[d];
        tc[d]      = box.lim[2 * d] + T(hi[d] + lo[d]) * halfUnit;
        ts[d]      = T(hi[d] - lo[d]) * halfUnit;
    }

    auto check = [&](int idx)
    {
        K nodePrefix         = prefixes[idx];
        unsigned sourceLevel = decodePrefixLength(nodePrefix) / 3;
        K nodeStart          = decodePlaceholderBit(nodePrefix);
        K nodeEnd            = nodeStart + nodeRange<K>(sourceLevel);
        if (!(nodeStart < focusStart || nodeEnd > focusEnd)) { return false; } // fully inside the focus
        // evaluateMacPbc (macs.hpp:118-130)
        T dx[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            T v   = tc[d] - centers4[4 * idx + d];
            v     = rabs(pbcFold(v, d, box));
            v -= ts[d];
            v += rabs(v);
            v *= T(0.5);
            dx[d] = v;
        }
        T R2        = dx[0] * dx[0] + (dx[1] * dx[1] + dx[2] * dx[2]);
        bool violates = R2 < rabs(centers4[4 * idx + 3]);
        if (violates && !markings[idx]) { markings[idx] = 1; }
        return violates;
    };

    // singleTraversal (traversal/traversal.hpp:26-69)
    if (!check(0)) { return; }
    int node = childOffsets[0];
    if (node == 0) { return; }
    bool backtrack = false;
    while (node != 0)
    {
        int child    = childOffsets[node];
        bool isLeaf  = child == 0;
        bool descend = !backtrack && check(node);
        if (!isLeaf && descend)
        {
            node      = child;
            backtrack = false;
        }
        else if (((node - 1) & 7) < 7)
        {
            ++node;
            backtrack = false;
        }
        else
        {
            node      = parents[(node - 1) >> 3];
            backtrack = true;
        }
    }
}

//! rangeCount (focus/rebalance.hpp:263-284) with the global counts given as an exclusive 64-bit scan
template<class K>
__global__ void rangeCountKernel(const K* __restrict__ gLeaves, int numGlobalLeaves,
                                 const uint64_t* __restrict__ gCountScan, const K* __restrict__ fLeaves,
                                 const int* __restrict__ idx, int numIdx, uint32_t* __restrict__ leafCounts)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numIdx) { return; }
    int leaf   = idx[t];
    K startKey = fLeaves[leaf];
    K endKey   = fLeaves[leaf + 1];
    // findNodeBelow = upper_bound - 1, findNodeAbove = lower_bound over the numGlobalLeaves + 1 keys
    int s = upperBound(gLeaves, numGlobalLeaves + 1, startKey) - 1;
    int e = lowerBound(gLeaves, numGlobalLeaves + 1, endKey);
    uint64_t c       = gCountScan[e] - gCountScan[s];
    leafCounts[leaf] = uint32_t(c < 0xFFFFFFFFull ? c : 0xFFFFFFFFull);
}

__global__ void gatherU32Kernel(const int* __restrict__ idx, int n, const uint32_t* __restrict__ src,
                                uint32_t* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { dst[i] = src[idx[i]]; }
}

__global__ void scatterU32Kernel(const int* __restrict__ idx, int n, const uint32_t* __restrict__ src,
                                 uint32_t* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { dst[idx[i]] = src[i]; }
}

/*! gatherRanges (halos/gather_halos_gpu.cu:27-58) for four arrays at once: element k of the packed message comes from
 *  range r = upper_bound(scan, k) - 1 at offset rangeStart[r] + k - scan[r] */
template<class E>
__global__ void gatherRanges4Kernel(const uint32_t* __restrict__ rangeScan, const uint32_t* __restrict__ rangeStart,
                                    int numRanges, uint32_t total, const E* __restrict__ a, const E* __restrict__ b,
                                    const E* __restrict__ c, const E* __restrict__ d, E* __restrict__ out,
                                    size_t blockElems)
{
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) { return; }
    int r        = int(upperBound(rangeScan, numRanges, k)) - 1;
    uint32_t src = rangeStart[r] + (k - rangeScan[r]);
    out[k]                  = a[src];
    out[blockElems + k]     = b[src];
    out[2 * blockElems + k] = c[src];
    out[3 * blockElems + k] = d[src];
}

} // namespace

template<class T>
int minMacCenters(const T* geoCenters, const T* geoSizes, int numNodes, float invThetaEff, T* centers4, cudaStream_t s)
{
    if (numNodes == 0) { return 0; }
    minMacKernel<T><<<iceil(numNodes, 256), 256, 0, s>>>(geoCenters, geoSizes, numNodes, invThetaEff, centers4);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K, class T>
int markMacs(const K* prefixes, const int* childOffsets, const int* parents, const T* centers4, const double* lim,
             const int* bnd, const K* focusNodes, int numFocusNodes, uint8_t* markings, cudaStream_t s)
{
    if (numFocusNodes <= 0) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);
    markMacsKernel<K, T><<<iceil(numFocusNodes, 128), 128, 0, s>>>(prefixes, childOffsets, parents, centers4, box,
                                                                   focusNodes, numFocusNodes, markings);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int rangeCount(const K* gLeaves, int numGlobalLeaves, const uint64_t* gCountScan, const K* fLeaves, const int* idx,
               int numIdx, uint32_t* leafCounts, cudaStream_t s)
{
    if (numIdx == 0) { return 0; }
    rangeCountKernel<K><<<iceil(numIdx, 256), 256, 0, s>>>(gLeaves, numGlobalLeaves, gCountScan, fLeaves, idx, numIdx,
                                                           leafCounts);
    CSB_LAUNCH_CHECK();
    return 0;
}

int gatherU32(const int* idx, int n, const uint32_t* src, uint32_t* dst, cudaStream_t s)
{
    if (n == 0) { return 0; }
    gatherU32Kernel<<<iceil(n, 256), 256, 0, s>>>(idx, n, src, dst);
    CSB_LAUNCH_CHECK();
    return 0;
}

int scatterU32(const int* idx, int n, const uint32_t* src, uint32_t* dst, cudaStream_t s)
{
    if (n == 0) { return 0; }
    scatterU32Kernel<<<iceil(n, 256), 256, 0, s>>>(idx, n, src, dst);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class E>
int gatherRanges4(const uint32_t* rangeScan, const uint32_t* rangeStart, int numRanges, uint32_t total, const E* a,
                  const E* b, const E* c, const E* d, E* out, size_t blockElems, cudaStream_t s)
{
    if (total == 0) { return 0; }
    gatherRanges4Kernel<E><<<iceil(total, 256), 256, 0, s>>>(rangeScan, rangeStart, numRanges, total, a, b, c, d, out,
                                                             blockElems);
    CSB_LAUNCH_CHECK();
    return 0;
}

template int minMacCenters<float>(const float*, const float*, int, float, float*, cudaStream_t);
template int minMacCenters<double>(const double*, const double*, int, float, double*, cudaStream_t);
template int markMacs<uint32_t, float>(const uint32_t*, const int*, const int*, const float*, const double*, const int*,
                                       const uint32_t*, int, uint8_t*, cudaStream_t);
template int markMacs<uint64_t, float>(const uint64_t*, const int*, const int*, const float*, const double*, const int*,
                                       const uint64_t*, int, uint8_t*, cudaStream_t);
template int markMacs<uint64_t, double>(const uint64_t*, const int*, const int*, const double*, const double*,
                                        const int*, const uint64_t*, int, uint8_t*, cudaStream_t);
template int rangeCount<uint32_t>(const uint32_t*, int, const uint64_t*, const uint32_t*, const int*, int, uint32_t*,
                                  cudaStream_t);
template int rangeCount<uint64_t>(const uint64_t*, int, const uint64_t*, const uint64_t*, const int*, int, uint32_t*,
                                  cudaStream_t);
template int gatherRanges4<float>(const uint32_t*, const uint32_t*, int, uint32_t, const float*, const float*,
                                  const float*, const float*, float*, size_t, cudaStream_t);
template int gatherRanges4<double>(const uint32_t*, const uint32_t*, int, uint32_t, const double*, const double*,
                                   const double*, const double*, double*, size_t, cudaStream_t);

} // namespace csb
