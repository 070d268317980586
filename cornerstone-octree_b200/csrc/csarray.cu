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
 * Leaves and keys are both sorted, so a chunk of consecutive leaves needs one contiguous window of keys:
 *   1. coarseBoundsKernel: lower bound of the first leaf key of every chunk among all keys (a warp-cooperative 32-ary
 *      search per chunk);
 *   2. nodeCountsPipelinedKernel: the key windows are streamed through shared memory - every key is read from HBM
 *      once - and the threads search their leaf keys there.  Windows that do not fit (coarse trees with few leaves) are
 *      searched in global memory instead, bounded by the window.
 */
/*! lower bound of every NC_LEAVES-th leaf key among all keys: one WARP per search, 32 probes per round (a 32-ary
 *  search: 6 dependent rounds for 64 Mi keys instead of the 26 of a binary search - this latency chain was 12 % of the
 *  stage, profiles/r1_notes.md) */
template<class K>
__global__ void coarseBoundsKernel(const K* __restrict__ leaves, int numLeaves, const K* __restrict__ keys, size_t n,
                                   int numChunks, int NC_LEAVES, uint64_t* __restrict__ coarse)
{
    const int j         = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (j > numChunks) { return; }
    const K v = leaves[min(j * NC_LEAVES, numLeaves)];
    // invariant: keys before lo are < v, keys from lo + len on are >= v
    size_t lo = 0, len = n;
    while (len > 0)
    {
        const size_t step = (len + 31) / 32;
        const size_t q    = lo + (lane + 1) * step - 1; // probes are ascending: the lanes that see a smaller key form a prefix
        const bool valid  = q < lo + len;
        const bool less   = valid && keys[q] < v;
        const unsigned c  = __popc(__ballot_sync(0xffffffffu, less));
        const size_t end  = lo + len;
        lo += c * step;
        // probe c (if it exists) holds a key >= v: the answer is at most its position
        const size_t qc = lo + step - 1;
        len             = qc < end ? step - 1 : end - lo;
    }
    if (lane == 0) { coarse[j] = lo; }
}

/* ---- persistent blocks, key windows brought in by the bulk-copy engine ----
 * A block that loads one window, searches it and exits exposes the DRAM latency of its loads once per window (round 1:
 * 0.154 ms, 0.56 of the HBM peak).  Here each block walks its share of the chunks with the key windows of the next
 * STAGES - 1 chunks in flight (0.137 ms, 0.63 of the peak; shapes tried: profiles/r2_notes.md): thread 0 issues one `cp.async.bulk` (global -> shared, 1D, completion counted in bytes on an mbarrier) per
 * window, the threads only search.  The bulk engine needs 16-byte aligned addresses and sizes, so the copied range is
 * the window rounded outwards to 16 bytes (clipped to the key array; clipped-off head / tail keys - at most 3, only in
 * the first and last window of an array - are loaded normally). */
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra WAIT_DONE;\n bra WAIT_LOOP;\n WAIT_DONE:\n}" ::"r"(smemAddr(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void* dstShared, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(dstShared)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

template<int LEAVES, int WINDOW, int STAGES>
struct NcPipeShape
{
    static constexpr int leaves = LEAVES, window = WINDOW, stages = STAGES;
};

struct NcStageMeta
{
    unsigned long long first; // index of the first key of the window (coarse[chunk])
    uint32_t width;           // number of keys in the window
    uint32_t shift;           // window[i] = buffer[shift + i]
    uint32_t headEnd;         // window keys [0, headEnd) and
    uint32_t tailBegin;       // [tailBegin, width) were clipped off the bulk copy: loaded normally
    uint32_t staged;          // 0: window too long for shared memory, searched in global memory
};

template<class K, class Shape>
__global__ void __launch_bounds__(Shape::leaves) nodeCountsPipelinedKernel(const K* __restrict__ leaves,
                                                                          uint32_t* __restrict__ counts,
                                                                          int numLeaves,
                                                                          const K* __restrict__ keys,
                                                                          size_t n,
                                                                          const uint64_t* __restrict__ coarse,
                                                                          int numChunks,
                                                                          uint32_t maxCount)
{
    constexpr int LEAVES = Shape::leaves, WINDOW = Shape::window, STAGES = Shape::stages;
    constexpr int BUF = WINDOW + 48 / int(sizeof(K)); // keys per buffer: the window plus alignment slack on both sides
    extern __shared__ __align__(16) unsigned char ncSmem[];
    K* buffers = reinterpret_cast<K*>(ncSmem);
    __shared__ __align__(8) uint64_t bar[STAGES];
    __shared__ NcStageMeta meta[STAGES];
    __shared__ uint32_t bound[LEAVES + 1];

    const int t = threadIdx.x;
    if (t == 0)
    {
        for (int b = 0; b < STAGES; ++b)
            mbarInit(&bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    //! thread 0: describe the window of `chunk` in stage b and start its bulk copy
    auto issue = [&](int chunk, int b)
    {
        const unsigned long long A = coarse[chunk], B = coarse[chunk + 1];
        const uint32_t W           = uint32_t(B - A);
        NcStageMeta m;
        m.first  = A;
        m.width  = W;
        m.staged = W <= uint32_t(WINDOW);
        m.shift = 0, m.headEnd = 0, m.tailBegin = W;
        uint32_t bytes = 0;
        K* dst         = nullptr;
        uintptr_t lo   = 0;
        if (m.staged && W)
        {
            const uintptr_t base = reinterpret_cast<uintptr_t>(keys), end = base + n * sizeof(K);
            const uintptr_t a = base + A * sizeof(K), e = base + B * sizeof(K);
            lo                = a & ~uintptr_t(15);          // rounded outwards ...
            uintptr_t hi      = (e + 15) & ~uintptr_t(15);
            if (lo < base) { lo += 16; }                     // ... and clipped to the array
            if (hi > end) { hi -= 16; }
            // buffer position of key A: 16 bytes of slack in front keep the copy destination aligned like its source
            m.shift = uint32_t((16 + (a & 15)) / sizeof(K));
            if (hi > lo)
            {
                bytes       = uint32_t(hi - lo);
                m.headEnd   = lo > a ? uint32_t((lo - a) / sizeof(K)) : 0u;
                m.tailBegin = hi < e ? uint32_t((hi - a) / sizeof(K)) : W;
                dst         = buffers + size_t(b) * BUF + m.shift + (long long)(lo - a) / (long long)sizeof(K);
            }
            else { m.headEnd = W; } // the whole (tiny) window is loaded normally
        }
        /* the description of the stage is written BEFORE the barrier is armed: the arrival below releases it to the
         * threads that wait on the barrier.  (Written after the copy was started, a short copy could complete the phase
         * before the description was in place - found by compute-sanitizer, which slows this thread down.) */
        meta[b] = m;
        mbarExpectTx(&bar[b], bytes);
        if (bytes) { bulkCopyG2S(dst, reinterpret_cast<const void*>(lo), bytes, &bar[b]); }
    };

    // prologue: the first STAGES - 1 windows of this block
    if (t == 0)
    {
        for (int k = 0; k < STAGES - 1; ++k)
        {
            const int chunk = blockIdx.x + k * gridDim.x;
            if (chunk < numChunks) { issue(chunk, k); }
        }
    }
    int chunk = blockIdx.x;
    K mine    = K(0);
    if (chunk < numChunks) { mine = leaves[min(chunk * LEAVES + t, numLeaves)]; }

    for (int it = 0; chunk < numChunks; ++it, chunk += gridDim.x)
    {
        const int b = it % STAGES;
        // keep STAGES - 1 windows in flight: the stage that was consumed in the previous iteration is free again
        const int ahead = chunk + (STAGES - 1) * gridDim.x;
        if (t == 0 && ahead < numChunks) { issue(ahead, (it + STAGES - 1) % STAGES); }
        // this thread's leaf key of the next chunk, one iteration ahead of its use
        const int nextChunk = chunk + gridDim.x;
        K mineNext          = K(0);
        if (nextChunk < numChunks) { mineNext = leaves[min(nextChunk * LEAVES + t, numLeaves)]; }

        mbarWait(&bar[b], uint32_t(it / STAGES) & 1u);
        const NcStageMeta m = meta[b];
        const int c0        = chunk * LEAVES;
        const int nl        = min(LEAVES, numLeaves - c0);
        K* window           = buffers + size_t(b) * BUF + m.shift;
        if (m.staged)
        {
            if (m.headEnd || m.tailBegin < m.width) // first / last window of the key array only
            {
                for (uint32_t i = t; i < m.headEnd; i += LEAVES)
                    window[i] = keys[m.first + i];
                for (uint32_t i = m.tailBegin + t; i < m.width; i += LEAVES)
                    window[i] = keys[m.first + i];
                __syncthreads();
            }
            if (t < nl) { bound[t] = lowerBound(window, m.width, mine); }
        }
        else if (t < nl) { bound[t] = uint32_t(lowerBound(keys + m.first, size_t(m.width), mine)); }
        if (t == 0) { bound[nl] = m.width; }
        __syncthreads();
        if (t < nl)
        {
            const uint32_t c = bound[t + 1] - bound[t];
            counts[c0 + t]   = c < maxCount ? c : maxCount;
        }
        __syncthreads(); // stage b and bound[] are free
        mine = mineNext;
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
    using Shape         = NcPipeShape<128, 6144, 2>; // 2 x 48 KiB of 64-bit keys per block, two blocks per SM
    const int numChunks = int(iceil(numLeaves, Shape::leaves));
    CSB_SCRATCH(coarse, uint64_t*, s, SCRATCH_E, (size_t(numChunks) + 1) * sizeof(uint64_t));
    coarseBoundsKernel<K><<<iceil((size_t(numChunks) + 1) * 32, 128), 128, 0, s>>>(leaves, numLeaves, keys, n, numChunks,
                                                                                    Shape::leaves, coarse);
    CSB_LAUNCH_CHECK();
    int dev = 0, numSm = 0;
    CSB_CHECK(cudaGetDevice(&dev));
    CSB_CHECK(cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev));
    constexpr size_t smem = size_t(Shape::stages) * (Shape::window + 48 / sizeof(K)) * sizeof(K);
    CSB_CHECK(cudaFuncSetAttribute(nodeCountsPipelinedKernel<K, Shape>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   int(smem)));
    const int grid = std::min(numChunks, numSm * 2);
    nodeCountsPipelinedKernel<K, Shape><<<grid, Shape::leaves, smem, s>>>(leaves, counts, numLeaves, keys, n, coarse,
                                                                         numChunks, maxCount);
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
