/* Stable merge of sorted runs of (key, index) pairs for sm_100a.
 *
 * After exchangeParticles the assigned particles of a rank are P sorted runs: the present particles (a piece of the
 * first sort) and one block per source rank, each in its sender's key order.  The reference sorts the whole buffer
 * again (domain/assignment.hpp:197-201, sortByKey); a stable merge of the runs in buffer order gives the same
 * permutation - ties between equal keys go to the earlier run, exactly like the stable sort - with log2(P) streaming
 * passes of 2 (K + 4) bytes per element instead of 8 radix passes.
 *
 * Pairwise rounds.  Per round: mergePartitionKernel finds, for every output tile boundary, how many elements of the
 * first run precede it (merge path: binary search along the cross diagonal), mergeTilesKernel loads the two pieces of
 * a tile into shared memory with coalesced loads, every thread locates its own diagonal there, merges MG_VT elements
 * serially and the tile leaves through shared memory with coalesced stores.
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

/* tile shape: 256 threads x 4 elements.  Measured on 8 runs of 8 Mi (u64, u32) pairs (tools/exp_merge.py): 2.79 ms with
 * 256 x 16, 1.92 with 256 x 8, 1.62 with 256 x 4, 1.70 with 128 x 4, 1.90 with 256 x 2 - short serial merges and many
 * resident tiles hide the shared-memory latency of the merge steps better than long ones */
constexpr int MG_THREADS  = 256;
constexpr int MG_VT       = 4;
constexpr int MG_TILE     = MG_THREADS * MG_VT;
constexpr int MG_MAXPAIRS = 64;

struct MergePair
{
    unsigned long long aStart, aLen, bStart, bLen, outStart;
    unsigned firstTile, numTiles;
};

struct MergeTable
{
    MergePair pair[MG_MAXPAIRS];
    int numPairs;
};

//! number of elements taken from a[0, lenA) among the first diag outputs of the stable merge of a and b (a wins ties)
template<class K, class I>
__device__ inline I mergePath(const K* a, I lenA, const K* b, I lenB, I diag)
{
    I lo = diag > lenB ? diag - lenB : 0;
    I hi = diag < lenA ? diag : lenA;
    while (lo < hi)
    {
        I mid = lo + (hi - lo) / 2;
        if (!(b[diag - 1 - mid] < a[mid])) { lo = mid + 1; }
        else { hi = mid; }
    }
    return lo;
}

__device__ inline int pairOfTile(const MergeTable& t, unsigned tile)
{
    int p = 0;
    while (p + 1 < t.numPairs && tile >= t.pair[p + 1].firstTile)
        ++p;
    return p;
}

//! splits[tile + pairIndex] for every tile boundary of every pair (numTiles + 1 boundaries per pair)
template<class K>
__global__ void mergePartitionKernel(const K* __restrict__ keys, MergeTable table, unsigned totalBoundaries,
                                     unsigned long long* __restrict__ splits)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= totalBoundaries) { return; }
    // boundary i belongs to pair p if firstTile[p] + p <= i <= firstTile[p] + p + numTiles[p]
    int p = 0;
    while (p + 1 < table.numPairs && i >= table.pair[p + 1].firstTile + unsigned(p + 1))
        ++p;
    const MergePair& mp     = table.pair[p];
    unsigned long long diag = (unsigned long long)(i - mp.firstTile - unsigned(p)) * MG_TILE;
    diag                    = diag < mp.aLen + mp.bLen ? diag : mp.aLen + mp.bLen;
    splits[i] = mergePath<K, unsigned long long>(keys + mp.aStart, mp.aLen, keys + mp.bStart, mp.bLen, diag);
}

template<class K>
__global__ void __launch_bounds__(MG_THREADS) mergeTilesKernel(const K* __restrict__ keysIn,
                                                               const uint32_t* __restrict__ valsIn,
                                                               K* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                               MergeTable table,
                                                               const unsigned long long* __restrict__ splits)
{
    __shared__ K keysS[MG_TILE];
    __shared__ uint32_t valsS[MG_TILE];

    const unsigned tile = blockIdx.x;
    const int p         = pairOfTile(table, tile);
    const MergePair& mp = table.pair[p];
    const unsigned lt   = tile - mp.firstTile;
    const unsigned long long total = mp.aLen + mp.bLen;
    const unsigned long long diag0 = (unsigned long long)lt * MG_TILE;
    const unsigned long long diag1 = diag0 + MG_TILE < total ? diag0 + MG_TILE : total;
    const unsigned long long a0 = splits[tile + p], a1 = splits[tile + p + 1];
    const unsigned long long b0 = diag0 - a0, b1 = diag1 - a1;
    const int aCount = int(a1 - a0), bCount = int(b1 - b0), count = aCount + bCount;

    const K* aKeys        = keysIn + mp.aStart + a0;
    const K* bKeys        = keysIn + mp.bStart + b0;
    const uint32_t* aVals = valsIn + mp.aStart + a0;
    const uint32_t* bVals = valsIn + mp.bStart + b0;
    for (int i = threadIdx.x; i < count; i += MG_THREADS)
    {
        keysS[i] = i < aCount ? aKeys[i] : bKeys[i - aCount];
        valsS[i] = i < aCount ? aVals[i] : bVals[i - aCount];
    }
    __syncthreads();

    // this thread's MG_VT outputs start at diagonal d of the tile
    const int d   = min(int(threadIdx.x) * MG_VT, count);
    int ai        = mergePath<K, int>(keysS, aCount, keysS + aCount, bCount, d);
    int bi        = aCount + (d - ai);
    const int bEnd = count;
    K outK[MG_VT];
    uint32_t outV[MG_VT];
    K aKey = ai < aCount ? keysS[ai] : K(0);
    K bKey = bi < bEnd ? keysS[bi] : K(0);
#pragma unroll
    for (int i = 0; i < MG_VT; ++i)
    {
        const bool takeA = bi >= bEnd || (ai < aCount && !(bKey < aKey));
        const int src    = takeA ? ai : bi;
        outK[i]          = takeA ? aKey : bKey;
        outV[i]          = d + i < count ? valsS[min(src, count - 1)] : 0u;
        if (takeA)
        {
            ++ai;
            aKey = ai < aCount ? keysS[ai] : K(0);
        }
        else
        {
            ++bi;
            bKey = bi < bEnd ? keysS[bi] : K(0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MG_VT; ++i)
    {
        if (d + i < count)
        {
            keysS[d + i] = outK[i];
            valsS[d + i] = outV[i];
        }
    }
    __syncthreads();
    K* ko        = keysOut + mp.outStart + diag0;
    uint32_t* vo = valsOut + mp.outStart + diag0;
    for (int i = threadIdx.x; i < count; i += MG_THREADS)
    {
        ko[i] = keysS[i];
        vo[i] = valsS[i];
    }
}

} // namespace

/*! keys/vals hold numRuns sorted runs, run r = [runOffsets[r], runOffsets[r+1]); on return the whole range is sorted,
 *  equal keys in run order (and in their order inside a run).  keyBuf/valBuf: double buffers of the same length.
 *  resultInBuffers (optional): if given and the runs start at element 0, a result that ends up in the double buffers
 *  after an odd number of rounds is left there and the flag is set (the caller swaps its buffers) instead of being
 *  copied back. */
template<class K>
int mergeSortedRuns(K* keys, uint32_t* vals, const size_t* runOffsets, int numRuns, K* keyBuf, uint32_t* valBuf,
                    cudaStream_t s, bool* resultInBuffers)
{
    if (resultInBuffers) { *resultInBuffers = false; }
    // run boundaries without the empty runs
    std::vector<size_t> runs{runOffsets[0]};
    for (int r = 1; r <= numRuns; ++r)
        if (runOffsets[r] != runs.back()) { runs.push_back(runOffsets[r]); }
    if (runs.size() <= 2) { return 0; } // zero or one non-empty run: sorted already

    K* kin         = keys;
    K* kout        = keyBuf;
    uint32_t* vin  = vals;
    uint32_t* vout = valBuf;
    const size_t base = runs.front(), total = runs.back() - runs.front();
    while (runs.size() > 2)
    {
        const size_t nr = runs.size() - 1;
        std::vector<size_t> next{runs[0]};
        size_t r = 0;
        while (r + 1 < nr)
        {
            MergeTable table{};
            unsigned tiles = 0;
            while (r + 1 < nr && table.numPairs < MG_MAXPAIRS)
            {
                MergePair& mp = table.pair[table.numPairs++];
                mp.aStart     = runs[r];
                mp.aLen       = runs[r + 1] - runs[r];
                mp.bStart     = runs[r + 1];
                mp.bLen       = runs[r + 2] - runs[r + 1];
                mp.outStart   = runs[r];
                mp.firstTile  = tiles;
                mp.numTiles   = unsigned((mp.aLen + mp.bLen + MG_TILE - 1) / MG_TILE);
                tiles += mp.numTiles;
                next.push_back(runs[r + 2]);
                r += 2;
            }
            unsigned boundaries = tiles + unsigned(table.numPairs);
            CSB_SCRATCH(splits, unsigned long long*, s, SCRATCH_E, size_t(boundaries) * sizeof(unsigned long long));
            mergePartitionKernel<K><<<iceil(boundaries, 128), 128, 0, s>>>(kin, table, boundaries, splits);
            CSB_LAUNCH_CHECK();
            mergeTilesKernel<K><<<tiles, MG_THREADS, 0, s>>>(kin, vin, kout, vout, table, splits);
            CSB_LAUNCH_CHECK();
        }
        if (r < nr)
        {
            // odd run out: carried to the next round unchanged
            size_t len = runs[r + 1] - runs[r];
            CSB_CHECK(cudaMemcpyAsync(kout + runs[r], kin + runs[r], len * sizeof(K), cudaMemcpyDeviceToDevice, s));
            CSB_CHECK(cudaMemcpyAsync(vout + runs[r], vin + runs[r], len * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                      s));
            next.push_back(runs[r + 1]);
        }
        runs.swap(next);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    if (kin != keys && resultInBuffers && base == 0) { *resultInBuffers = true; } // the caller swaps its buffers
    else if (kin != keys)
    {
        CSB_CHECK(cudaMemcpyAsync(keys + base, kin + base, total * sizeof(K), cudaMemcpyDeviceToDevice, s));
        CSB_CHECK(cudaMemcpyAsync(vals + base, vin + base, total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

template int mergeSortedRuns<uint32_t>(uint32_t*, uint32_t*, const size_t*, int, uint32_t*, uint32_t*, cudaStream_t,
                                       bool*);
template int mergeSortedRuns<uint64_t>(uint64_t*, uint32_t*, const size_t*, int, uint64_t*, uint32_t*, cudaStream_t,
                                       bool*);

} // namespace csb

extern "C"
{

/* stable merge of numRuns sorted runs (run r = [runOffsets[r], runOffsets[r+1]), host array of numRuns + 1 offsets) of
 * keys with their 32-bit values; keyBuf / valueBuf are double buffers of runOffsets[numRuns] elements */
int cs_merge_sorted_runs_u32(uint32_t* keys, uint32_t* values, const size_t* runOffsets, int numRuns, uint32_t* keyBuf,
                             uint32_t* valueBuf, void* stream)
{
    return csb::mergeSortedRuns<uint32_t>(keys, values, runOffsets, numRuns, keyBuf, valueBuf, cudaStream_t(stream),
                                          nullptr);
}

int cs_merge_sorted_runs_u64(uint64_t* keys, uint32_t* values, const size_t* runOffsets, int numRuns, uint64_t* keyBuf,
                             uint32_t* valueBuf, void* stream)
{
    return csb::mergeSortedRuns<uint64_t>(keys, values, runOffsets, numRuns, keyBuf, valueBuf, cudaStream_t(stream),
                                          nullptr);
}

} // extern "C"

