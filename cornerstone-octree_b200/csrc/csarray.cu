/* Cornerstone leaf-array ("csarray") kernels for sm_100a: node counts, rebalance decisions, rebalance, and the
 * converge-from-root driver.  Replaces the reference's tree/csarray_gpu.cu (computeNodeCountsGpu :86-122,
 * computeNodeOpsGpu :182-205, rebalanceTreeGpu :212-231) with the arithmetic of tree/csarray.hpp
 * (calculateNodeCount :68-79, siblingAndLevel :237-253, calculateNodeOp :267-293, processNode :339-370).
 * All integer work: results are bit-exact.
 */
#include <algorithm>

#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

/* ---------------------------------------------------------------- node counts */

/* counts[i] = lower_bound(keys, leaves[i+1]) - lower_bound(keys, leaves[i]) (calculateNodeCount, tree/csarray.hpp:68-79).
 * Leaves and keys are both sorted, so a block of NC_LEAVES consecutive leaves needs one contiguous window of keys:
 *   1. coarseBoundsKernel: lower bound of every NC_LEAVES-th leaf key by binary search over all keys (few searches,
 *      all in flight at once);
 *   2. nodeCountsKernel: each block streams its key window through shared memory with coalesced loads - every key is
 *      read from HBM once - and the threads search their leaf keys there.  Very long windows (coarse trees with few
 *      leaves) are searched in global memory instead, bounded by the window.
 */
constexpr int NC_LEAVES  = 64;   // leaves per block
constexpr int NC_THREADS = 256;  // threads per block: all of them stream keys, the first NC_LEAVES search
constexpr int NC_WINDOW  = 3072; // keys staged in shared memory per tile (24 KiB of 64-bit keys: 8 blocks per SM)
constexpr int NC_MAX_TILES = 8; // longer windows (coarse trees) are searched in global memory

template<class K>
__global__ void coarseBoundsKernel(const K* __restrict__ leaves, int numLeaves, const K* __restrict__ keys, size_t n,
                                   int numChunks, uint64_t* __restrict__ coarse)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > numChunks) { return; }
    int leaf  = min(j * NC_LEAVES, numLeaves);
    coarse[j] = lowerBound(keys, n, leaves[leaf]);
}

template<class K>
__global__ void __launch_bounds__(NC_THREADS) nodeCountsKernel(const K* __restrict__ leaves,
                                                              uint32_t* __restrict__ counts,
                                                              int numLeaves,
                                                              const K* __restrict__ keys,
                                                              const uint64_t* __restrict__ coarse,
                                                              uint32_t maxCount)
{
    extern __shared__ __align__(16) unsigned char ncSmem[];
    K* window = reinterpret_cast<K*>(ncSmem);
    __shared__ uint32_t bound[NC_LEAVES + 1];

    const int c0   = blockIdx.x * NC_LEAVES;
    const int nl   = min(NC_LEAVES, numLeaves - c0);
    const size_t A = coarse[blockIdx.x], B = coarse[blockIdx.x + 1];
    const size_t W = B - A;
    const int t    = threadIdx.x;
    const K mine   = leaves[c0 + min(t, nl)]; // thread nl would look for leaves[c0 + nl], whose bound is B

    if (W <= size_t(NC_WINDOW) * NC_MAX_TILES)
    {
        // stream the window through shared memory tile by tile; a thread's bound is the number of window keys smaller
        // than its leaf key: whole tiles below it count fully, the tile that straddles it is searched
        uint32_t below = 0;
        bool open      = t < nl;
        for (size_t base = 0; base < W; base += NC_WINDOW)
        {
            const uint32_t tw = uint32_t(min(size_t(NC_WINDOW), W - base));
            const K* src      = keys + A + base;
            if (base) { __syncthreads(); }
            // fixed trip count, fully unrolled: all 12 loads of a thread are in flight together (a `for (i < tw)`
            // loop issues them a few at a time and the block waits on one DRAM round trip after the other)
#pragma unroll
            for (int u = 0; u < NC_WINDOW / NC_THREADS; ++u)
            {
                uint32_t i = uint32_t(u) * NC_THREADS + t;
                if (i < tw) { window[i] = src[i]; }
            }
            __syncthreads();
            if (open)
            {
                if (window[tw - 1] < mine) { below += tw; }
                else
                {
                    below += lowerBound(window, tw, mine);
                    open = false;
                }
            }
        }
        if (t < nl) { bound[t] = below; }
    }
    else if (t < nl) { bound[t] = uint32_t(lowerBound(keys + A, W, mine)); } // W < 2^32: fewer than 2^32 keys per call
    if (t == 0) { bound[nl] = uint32_t(W); }
    __syncthreads();
    if (t < nl)
    {
        uint32_t c     = bound[t + 1] - bound[t];
        counts[c0 + t] = c < maxCount ? c : maxCount;
    }
}

/* ---------------------------------------------------------------- rebalance decision */

template<class K>
__device__ inline int nodeOp(const K* __restrict__ tree, int nodeIdx, const uint32_t* __restrict__ counts,
                             uint32_t bucketSize)
{
    constexpr unsigned maxLevel = KeyTraits<K>::maxLevel;
    K thisNode                  = tree[nodeIdx];
    K range                     = tree[nodeIdx + 1] - thisNode;
    unsigned level              = treeLevel(range);

    int siblingIdx = -1;
    if (level > 0)
    {
        siblingIdx    = int(octalDigit(thisNode, level));
        bool siblings = tree[nodeIdx - siblingIdx + 8] == tree[nodeIdx - siblingIdx] + nodeRange<K>(level - 1);
        if (!siblings) { siblingIdx = -1; }
    }

    if (siblingIdx > 0)
    {
        const uint32_t* g  = counts + nodeIdx - siblingIdx;
        uint64_t parentCnt = uint64_t(g[0]) + g[1] + g[2] + g[3] + g[4] + g[5] + g[6] + g[7];
        if (parentCnt <= uint64_t(bucketSize)) { return 0; }
    }

    uint32_t c = counts[nodeIdx];
    // the products wrap in 32-bit arithmetic exactly like the reference's `bucketSize * 512` (unsigned)
    if (c > bucketSize * 512u && level + 3 < maxLevel) { return 4096; }
    if (c > bucketSize * 64u && level + 2 < maxLevel) { return 512; }
    if (c > bucketSize * 8u && level + 1 < maxLevel) { return 64; }
    if (c > bucketSize && level < maxLevel) { return 8; }
    return 1;
}

template<class K>
__global__ void nodeOpsKernel(const K* __restrict__ tree,
                              int numLeaves,
                              const uint32_t* __restrict__ counts,
                              uint32_t bucketSize,
                              int* __restrict__ nodeOps,
                              uint32_t* changed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numLeaves) { return; }
    if (i == numLeaves)
    {
        nodeOps[i] = 0;
        return;
    }
    int op     = nodeOp(tree, i, counts, bucketSize);
    nodeOps[i] = op;
    if (op != 1) { *changed = 1u; } // benign race: every writer stores the same value
}

/* ---------------------------------------------------------------- rebalance */

template<class K>
__global__ void rebalanceKernel(const K* __restrict__ oldTree,
                                int numLeaves,
                                int newNumLeaves,
                                const int* __restrict__ nodeOps,
                                K* __restrict__ newTree)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { newTree[newNumLeaves] = oldTree[numLeaves]; }
    if (i >= numLeaves) { return; }

    K thisNode     = oldTree[i];
    unsigned level = treeLevel(K(oldTree[i + 1] - thisNode));
    int at         = nodeOps[i];
    int opCode     = nodeOps[i + 1] - at;
    if (opCode == 1) { newTree[at] = thisNode; }
    else if (opCode >= 8)
    {
        unsigned levelDiff = opCode == 8 ? 1 : opCode == 64 ? 2 : opCode == 512 ? 3 : 4;
        K childRange       = nodeRange<K>(level + levelDiff);
        for (int s = 0; s < opCode; ++s)
            newTree[at + s] = thisNode + K(s) * childRange;
    }
}

inline void* align256(void* p)
{
    return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
}

} // namespace

template<class K>
int computeNodeCounts(const K* leaves, uint32_t* counts, int numLeaves, const K* keys, size_t n, uint32_t maxCount,
                      cudaStream_t s)
{
    if (numLeaves <= 0) { return 0; }
    CSB_REQUIRE(n < (size_t(1) << 32), "computeNodeCounts supports fewer than 2^32 keys");
    const int numChunks = int(iceil(numLeaves, NC_LEAVES));
    CSB_SCRATCH(coarse, uint64_t*, s, SCRATCH_E, (size_t(numChunks) + 1) * sizeof(uint64_t));
    coarseBoundsKernel<K><<<iceil(numChunks + 1, 128), 128, 0, s>>>(leaves, numLeaves, keys, n, numChunks, coarse);
    CSB_LAUNCH_CHECK();
    constexpr size_t smem = size_t(NC_WINDOW) * sizeof(K);
    CSB_CHECK(cudaFuncSetAttribute(nodeCountsKernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    nodeCountsKernel<K><<<numChunks, NC_THREADS, smem, s>>>(leaves, counts, numLeaves, keys, coarse, maxCount);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int computeNodeOps(const K* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize, int* nodeOps, void* tmp,
                   int* newNumLeaves, int* converged, cudaStream_t s)
{
    CSB_REQUIRE(numLeaves > 0, "empty leaf array");
    uint32_t* flag = static_cast<uint32_t*>(align256(tmp));
    void* scanTmp  = flag + 64;
    CSB_CHECK(cudaMemsetAsync(flag, 0, sizeof(uint32_t), s));
    nodeOpsKernel<K><<<iceil(numLeaves + 1, 256), 256, 0, s>>>(leaves, numLeaves, counts, bucketSize, nodeOps, flag);
    CSB_LAUNCH_CHECK();
    if (int e = exclusiveScanU32(reinterpret_cast<uint32_t*>(nodeOps), reinterpret_cast<uint32_t*>(nodeOps),
                                 size_t(numLeaves) + 1, scanTmp, s))
    {
        return e;
    }
    uint32_t changed = 0;
    CSB_CHECK(cudaMemcpyAsync(newNumLeaves, nodeOps + numLeaves, sizeof(int), cudaMemcpyDeviceToHost, s));
    CSB_CHECK(cudaMemcpyAsync(&changed, flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CSB_CHECK(cudaStreamSynchronize(s));
    *converged = changed ? 0 : 1;
    return 0;
}

template<class K>
int rebalanceTree(const K* leaves, int numLeaves, int newNumLeaves, const int* nodeOps, K* newLeaves, cudaStream_t s)
{
    rebalanceKernel<K><<<iceil(numLeaves, 256), 256, 0, s>>>(leaves, numLeaves, newNumLeaves, nodeOps, newLeaves);
    CSB_LAUNCH_CHECK();
    return 0;
}

size_t nodeOpsTempBytes(size_t numLeaves) { return scanTempBytes(numLeaves + 1) + 1024; }

template<class K>
int computeOctree(const K* keys, size_t n, uint32_t bucketSize, K* leaves, uint32_t* counts, int capacity,
                  int* numLeavesOut, cudaStream_t s)
{
    CSB_REQUIRE(capacity >= 1, "leaf capacity must be positive");
    CSB_SCRATCH(alt, K*, s, SCRATCH_A, (size_t(capacity) + 1) * sizeof(K));
    CSB_SCRATCH(nodeOps, int*, s, SCRATCH_B, (size_t(capacity) + 1) * sizeof(int));
    CSB_SCRATCH(tmp, void*, s, SCRATCH_C, nodeOpsTempBytes(capacity));

    K root[2]      = {0, nodeRange<K>(0)};
    uint32_t cnt0  = uint32_t(std::min<size_t>(n, 0xFFFFFFFFu));
    K* cur         = leaves;
    K* nxt         = alt;
    int status     = 0;
    int numLeaves  = 1;
    CSB_CHECK(cudaMemcpyAsync(cur, root, sizeof(root), cudaMemcpyHostToDevice, s));
    CSB_CHECK(cudaMemcpyAsync(counts, &cnt0, sizeof(cnt0), cudaMemcpyHostToDevice, s));
    CSB_CHECK(cudaStreamSynchronize(s)); // root/cnt0 live on the stack

    int converged = 0;
    while (!converged && status == 0)
    {
        int newNumLeaves = 0;
        status = computeNodeOps<K>(cur, numLeaves, counts, bucketSize, nodeOps, tmp, &newNumLeaves, &converged, s);
        if (status) { break; }
        if (newNumLeaves > capacity)
        {
            *numLeavesOut = newNumLeaves;
            setLastError("cs_compute_octree: leaf capacity too small");
            status = 3;
            break;
        }
        status = rebalanceTree<K>(cur, numLeaves, newNumLeaves, nodeOps, nxt, s);
        if (status) { break; }
        std::swap(cur, nxt);
        numLeaves = newNumLeaves;
        status    = computeNodeCounts<K>(cur, counts, numLeaves, keys, n, 0xFFFFFFFFu, s);
    }
    if (status == 0)
    {
        if (cur != leaves)
        {
            CSB_CHECK(cudaMemcpyAsync(leaves, cur, (size_t(numLeaves) + 1) * sizeof(K), cudaMemcpyDeviceToDevice, s));
        }
        *numLeavesOut = numLeaves;
    }
    CSB_CHECK(cudaStreamSynchronize(s));
    return status;
}

template int computeNodeCounts<uint32_t>(const uint32_t*, uint32_t*, int, const uint32_t*, size_t, uint32_t,
                                         cudaStream_t);
template int computeNodeCounts<uint64_t>(const uint64_t*, uint32_t*, int, const uint64_t*, size_t, uint32_t,
                                         cudaStream_t);
template int computeNodeOps<uint32_t>(const uint32_t*, int, const uint32_t*, uint32_t, int*, void*, int*, int*,
                                      cudaStream_t);
template int computeNodeOps<uint64_t>(const uint64_t*, int, const uint32_t*, uint32_t, int*, void*, int*, int*,
                                      cudaStream_t);
template int rebalanceTree<uint32_t>(const uint32_t*, int, int, const int*, uint32_t*, cudaStream_t);
template int rebalanceTree<uint64_t>(const uint64_t*, int, int, const int*, uint64_t*, cudaStream_t);

} // namespace csb

extern "C"
{

int cs_compute_node_counts_u32(const uint32_t* leaves, uint32_t* counts, int numLeaves, const uint32_t* keys, size_t n,
                               uint32_t maxCount, void* stream)
{
    return csb::computeNodeCounts<uint32_t>(leaves, counts, numLeaves, keys, n, maxCount, cudaStream_t(stream));
}
int cs_compute_node_counts_u64(const uint64_t* leaves, uint32_t* counts, int numLeaves, const uint64_t* keys, size_t n,
                               uint32_t maxCount, void* stream)
{
    return csb::computeNodeCounts<uint64_t>(leaves, counts, numLeaves, keys, n, maxCount, cudaStream_t(stream));
}

size_t cs_node_ops_temp_bytes(int numLeaves) { return csb::nodeOpsTempBytes(size_t(numLeaves)); }

int cs_compute_node_ops_u32(const uint32_t* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize,
                            int* nodeOps, void* tmp, int* newNumLeaves, int* converged, void* stream)
{
    return csb::computeNodeOps<uint32_t>(leaves, numLeaves, counts, bucketSize, nodeOps, tmp, newNumLeaves, converged,
                                         cudaStream_t(stream));
}
int cs_compute_node_ops_u64(const uint64_t* leaves, int numLeaves, const uint32_t* counts, uint32_t bucketSize,
                            int* nodeOps, void* tmp, int* newNumLeaves, int* converged, void* stream)
{
    return csb::computeNodeOps<uint64_t>(leaves, numLeaves, counts, bucketSize, nodeOps, tmp, newNumLeaves, converged,
                                         cudaStream_t(stream));
}

int cs_rebalance_tree_u32(const uint32_t* leaves, int numLeaves, int newNumLeaves, const int* nodeOps,
                          uint32_t* newLeaves, void* stream)
{
    return csb::rebalanceTree<uint32_t>(leaves, numLeaves, newNumLeaves, nodeOps, newLeaves, cudaStream_t(stream));
}
int cs_rebalance_tree_u64(const uint64_t* leaves, int numLeaves, int newNumLeaves, const int* nodeOps,
                          uint64_t* newLeaves, void* stream)
{
    return csb::rebalanceTree<uint64_t>(leaves, numLeaves, newNumLeaves, nodeOps, newLeaves, cudaStream_t(stream));
}

int cs_compute_octree_u32(const uint32_t* keys, size_t n, uint32_t bucketSize, uint32_t* leaves, uint32_t* counts,
                          int capacity, int* numLeaves, void* stream)
{
    return csb::computeOctree<uint32_t>(keys, n, bucketSize, leaves, counts, capacity, numLeaves, cudaStream_t(stream));
}
int cs_compute_octree_u64(const uint64_t* keys, size_t n, uint32_t bucketSize, uint64_t* leaves, uint32_t* counts,
                          int capacity, int* numLeaves, void* stream)
{
    return csb::computeOctree<uint64_t>(keys, n, bucketSize, leaves, counts, capacity, numLeaves, cudaStream_t(stream));
}

} // extern "C"
