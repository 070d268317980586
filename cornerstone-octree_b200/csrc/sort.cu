/* Stable key+index LSD radix sort for sm_100a ("onesweep": one histogram pass, then one fused
 * rank+look-back+scatter pass per 8-bit digit), plus sequence/gather helpers.
 *
 * Replaces cstone::sortByKey(Gpu,...) = cub::DeviceRadixSort::SortPairs (reference primitives/primitives_gpu.cu:310-356),
 * sequence (:56-63) and the gather kernels (:74-90).  Results are fully specified (stable ascending order), so
 * parity is bit-exact against the CPU std::stable_sort path (primitives/gather.hpp:42-67).
 *
 * Per digit pass, a 384-thread CTA owns one tile of keys (dynamic tile ids, so predecessors are always resident):
 *   warp-striped load -> per-warp digit ranks from 8 ballots (stable: rank order == input order)
 *   -> CTA digit histogram -> publish AGGREGATE -> decoupled look-back over predecessor tiles -> publish INCLUSIVE
 *   -> keys/values staged in shared memory in locally sorted order -> run-coalesced stores.
 * Algorithmic traffic: K (histogram read) + passes * 2 * (K + 4) bytes per element (200 B for u64 keys + u32 index).
 */
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

constexpr int RADIX_BITS   = 8;
constexpr int RADIX        = 1 << RADIX_BITS;
constexpr int LB_WINDOW = 8; // predecessor tiles read per look-back round trip
constexpr int MIN_TILE = 2048; // smallest tile of any kernel variant (sizes the look-back state)

constexpr uint32_t FLAG_AGG   = 1u << 30;
constexpr uint32_t FLAG_INCL  = 2u << 30;
constexpr uint32_t FLAG_MASK  = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

template<class K>
struct SortCfg
{
    static constexpr int passes = sizeof(K); // 8-bit digits over all key bits
};

template<class K>
constexpr size_t sortSmemBytes(int threads, int ipt)
{
    return size_t(threads) * ipt * sizeof(K) + size_t(threads) * ipt * 4 + size_t(threads / 32) * RADIX * 4 +
           RADIX * 4 + 64 * 4;
}

//! kernel variant used by sortByKey: 0 = 512x12 (default, fastest in tools/exp_sort.py with the windowed look-back),
//! 1 = 256x12, 2 = 256x15, 3 = 384x12, 4 = 256x9 (threads x keys/thread for 64-bit keys; 32-bit keys take 4/3 as many)
int g_sortVariant = 0;
int g_sortDebugNoLookback = 0; // experiments only: wrong results, isolates the cost of the look-back chain

/* ---------------------------------------------------------------- histogram of all digit places */

template<class K>
__global__ void __launch_bounds__(512) radixHistogramKernel(const K* __restrict__ keys, size_t n, uint32_t* globalHist)
{
    constexpr int P   = SortCfg<K>::passes;
    constexpr int VEC = 16 / sizeof(K); // keys per 128-bit load
    constexpr int ILP = 4;              // independent loads in flight per thread
    __shared__ uint32_t hist[P * RADIX];
    for (int i = threadIdx.x; i < P * RADIX; i += blockDim.x)
        hist[i] = 0;
    __syncthreads();

    auto add = [&](K key)
    {
#pragma unroll
        for (int p = 0; p < P; ++p)
            atomicAdd(&hist[p * RADIX + unsigned((key >> (RADIX_BITS * p)) & (RADIX - 1))], 1u);
    };

    const size_t tid      = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t nthreads = size_t(gridDim.x) * blockDim.x;
    // head up to the first 16-byte boundary, vector body, tail
    size_t head = (reinterpret_cast<uintptr_t>(keys) & 15) ? (16 - (reinterpret_cast<uintptr_t>(keys) & 15)) / sizeof(K) : 0;
    head        = head < n ? head : n;
    const size_t numVec = (n - head) / VEC;
    const uint4* vkeys  = reinterpret_cast<const uint4*>(keys + head);
    for (size_t v = tid; v < numVec; v += nthreads * ILP)
    {
        uint4 q[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            size_t idx = v + size_t(u) * nthreads;
            q[u]       = idx < numVec ? vkeys[idx] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u)
        {
            if (v + size_t(u) * nthreads >= numVec) { break; }
            if constexpr (sizeof(K) == 8)
            {
                add(K(q[u].x) | (K(q[u].y) << 32));
                add(K(q[u].z) | (K(q[u].w) << 32));
            }
            else
            {
                add(K(q[u].x));
                add(K(q[u].y));
                add(K(q[u].z));
                add(K(q[u].w));
            }
        }
    }
    for (size_t i = tid; i < head; i += nthreads)
        add(keys[i]);
    for (size_t i = head + numVec * VEC + tid; i < n; i += nthreads)
        add(keys[i]);

    __syncthreads();
    for (int i = threadIdx.x; i < P * RADIX; i += blockDim.x)
    {
        uint32_t c = hist[i];
        if (c) { atomicAdd(&globalHist[i], c); }
    }
}

//! one 256-thread block per digit place: in-place exclusive scan of the 256 bin counts
__global__ void __launch_bounds__(RADIX) scanHistogramKernel(uint32_t* globalHist)
{
    __shared__ uint32_t warpSums[RADIX / 32];
    uint32_t* h   = globalHist + blockIdx.x * RADIX;
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v    = h[threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= unsigned(o)) { incl += t; }
    }
    if (lane == 31) { warpSums[warp] = incl; }
    __syncthreads();
    uint32_t off = 0;
    for (unsigned w = 0; w < warp; ++w)
        off += warpSums[w];
    h[threadIdx.x] = off + incl - v;
}

/* ---------------------------------------------------------------- one digit pass */

/*! peers &= lanes whose digit agrees with mine in bit B.  Bit set: keep voters; bit clear: keep non-voters.  Spelled in
 *  PTX so that it stays at 4 instructions per bit (test, vote, select, and-xor); the C++ forms compile to 6. */
template<int B>
__device__ __forceinline__ void ballotStep(unsigned d, unsigned& peers)
{
    asm volatile("{\n"
                 " .reg .pred p;\n"
                 " .reg .b32 t, m;\n"
                 " and.b32 t, %1, %2;\n"
                 " setp.ne.b32 p, t, 0;\n"
                 " vote.sync.ballot.b32 t, p, 0xffffffff;\n"
                 " selp.b32 m, 0, 0xffffffff, p;\n"
                 " lop3.b32 %0, %0, t, m, 0x60;\n"
                 "}"
                 : "+r"(peers)
                 : "r"(d), "n"(1u << B));
}

template<class K, bool HAS_VALUES, int THREADS, int IPT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) onesweepKernel(const K* __restrict__ keysIn,
                                                                  K* __restrict__ keysOut,
                                                                  const uint32_t* __restrict__ valsIn,
                                                                  uint32_t* __restrict__ valsOut,
                                                                  size_t n,
                                                                  int shift,
                                                                  const uint32_t* __restrict__ digitBase,
                                                                  volatile uint32_t* tileStates,
                                                                  uint32_t* tileCounter,
                                                                  int debugNoLookback,
                                                                  long long iotaStart)
{
    constexpr int TILE         = THREADS * IPT;
    constexpr int SORT_THREADS = THREADS;
    constexpr int SORT_WARPS   = THREADS / 32;
    static_assert(THREADS >= RADIX, "one thread per digit bin in the scan phases");

    extern __shared__ __align__(16) unsigned char smemRaw[];
    K* keysS           = reinterpret_cast<K*>(smemRaw);
    uint32_t* valsS    = reinterpret_cast<uint32_t*>(keysS + TILE);
    uint32_t* warpHist = valsS + TILE;                  // [SORT_WARPS][RADIX]
    uint32_t* binBase  = warpHist + SORT_WARPS * RADIX; // [RADIX] global base minus local offset
    uint32_t* scratch  = binBase + RADIX;               // [64]

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) { scratch[32] = atomicAdd(tileCounter, 1u); }
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        warpHist[i] = 0;
    __syncthreads();
    const uint32_t tileIdx = scratch[32];
    const size_t tileBase  = size_t(tileIdx) * TILE;
    const uint32_t tileCount = uint32_t(min(size_t(TILE), n - tileBase));

    /* ---- load (warp-striped: element order == (warp, item, lane) order); 32-bit offsets inside the tile */
    K key[IPT];
    const uint32_t warpOff = warp * 32 * IPT + lane;
    const K* tileKeys      = keysIn + tileBase;
    if (tileCount == TILE)
    {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
            key[i] = tileKeys[warpOff + i * 32];
    }
    else
    {
#pragma unroll
        for (int i = 0; i < IPT; ++i)
        {
            uint32_t off = warpOff + i * 32;
            key[i]       = off < tileCount ? tileKeys[off] : K(~K(0));
        }
    }

    /* ---- stable ranks within the warp per digit */
    uint32_t rank[IPT];
    uint32_t* wh = warpHist + warp * RADIX;
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
        unsigned d      = unsigned((key[i] >> shift) & (RADIX - 1));
        // lanes holding the same digit, from 8 ballots: MATCH.ANY runs on the ADU pipe at ~1 warp instruction per
        // 64 cycles per SM on sm_100 and was the limiter of this kernel (profiles/r1_first_path_summary.txt)
        unsigned peers = 0xffffffffu;
        ballotStep<0>(d, peers);
        ballotStep<1>(d, peers);
        ballotStep<2>(d, peers);
        ballotStep<3>(d, peers);
        ballotStep<4>(d, peers);
        ballotStep<5>(d, peers);
        ballotStep<6>(d, peers);
        ballotStep<7>(d, peers);
        static_assert(RADIX_BITS == 8);
        unsigned leader = __ffs(peers) - 1;
        unsigned below  = __popc(peers & ((1u << lane) - 1u));
        uint32_t pre    = 0;
        if (lane == leader)
        {
            pre   = wh[d];
            wh[d] = pre + __popc(peers);
        }
        __syncwarp();
        pre     = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + below;
    }
    __syncthreads();

    /* ---- CTA histogram: exclusive scan over warps per digit, then exclusive scan over digits */
    uint32_t binCount = 0, incl = 0;
    if (tid < RADIX)
    {
        uint32_t running = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
        {
            uint32_t t              = warpHist[w * RADIX + tid];
            warpHist[w * RADIX + tid] = running;
            running += t;
        }
        binCount = running;
        incl     = binCount;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= unsigned(o)) { incl += t; }
        }
        if (lane == 31) { scratch[warp] = incl; }
    }
    __syncthreads();
    uint32_t binOffset = 0;
    if (tid < RADIX)
    {
        uint32_t off = 0;
        for (unsigned w = 0; w < warp; ++w)
            off += scratch[w];
        binOffset = off + incl - binCount; // position of this digit's run in the locally sorted tile
        // publish this tile's digit counts (tile-major: one coalesced 1 KiB record per tile)
        volatile uint32_t* myState = tileStates + size_t(tileIdx) * RADIX + tid;
        *myState                   = (tileIdx == 0 ? FLAG_INCL : FLAG_AGG) | binCount;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            warpHist[w * RADIX + tid] += binOffset;
    }
    __syncthreads(); // warpHist is final: the warps without a bin go on to stage their keys during the look-back
    if (tid < RADIX)
    {
        volatile uint32_t* myState = tileStates + size_t(tileIdx) * RADIX + tid;
        /* ---- decoupled look-back with a window of LB_WINDOW predecessors per round trip.  Consecutive tiles start
         *      ~40 ns apart, so a walk that pays one L2 round trip per predecessor keeps finding predecessors that
         *      have not resolved their own prefix yet (measured: 20 % of the pass); reading a window of states with
         *      independent loads resolves the same chain in one or two round trips. */
        uint32_t exclusive = 0;
        if (tileIdx != 0 && !debugNoLookback)
        {
            long long t = (long long)tileIdx - 1;
            bool done   = false;
            while (!done)
            {
                uint32_t sv[LB_WINDOW];
#pragma unroll
                for (int k = 0; k < LB_WINDOW; ++k)
                    sv[k] = (t - k >= 0) ? uint32_t(tileStates[size_t(t - k) * RADIX + tid]) : uint32_t(FLAG_INCL);
#pragma unroll
                for (int k = 0; k < LB_WINDOW; ++k)
                {
                    if (done) { break; }
                    if ((sv[k] & FLAG_MASK) == 0)
                    {
                        t -= k; // not published yet: poll again from here
                        break;
                    }
                    exclusive += sv[k] & VALUE_MASK;
                    if (sv[k] & FLAG_INCL) { done = true; }
                    else if (k == LB_WINDOW - 1) { t -= LB_WINDOW; }
                }
            }
            *myState = FLAG_INCL | (exclusive + binCount);
        }
        binBase[tid] = digitBase[tid] + exclusive - binOffset; // read after the barrier that follows the staging
    }

    /* ---- stage keys in locally sorted order */
#pragma unroll
    for (int i = 0; i < IPT; ++i)
    {
        unsigned d = unsigned((key[i] >> shift) & (RADIX - 1));
        rank[i] += wh[d];
        keysS[rank[i]] = key[i];
    }
    // the value loads are issued here so that their latency is covered by the key stores below (the registers of
    // key[] are free from this point on)
    uint32_t val[HAS_VALUES ? IPT : 1];
    if constexpr (HAS_VALUES)
    {
        const uint32_t* tileVals = valsIn + tileBase;
        if (iotaStart >= 0)
        {
            // first pass of an index sort: the values are the sequence iotaStart, iotaStart + 1, ... (the reference
            // fills them with sequence() first, primitives_gpu.cu:56-63); generated here instead of read
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                val[i] = uint32_t(iotaStart) + uint32_t(tileBase) + warpOff + i * 32;
        }
        else if (tileCount == TILE)
        {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                val[i] = tileVals[warpOff + i * 32];
        }
        else
        {
#pragma unroll
            for (int i = 0; i < IPT; ++i)
            {
                uint32_t off = warpOff + i * 32;
                val[i]       = off < tileCount ? tileVals[off] : 0u;
            }
        }
    }
    __syncthreads();

    /* ---- run-coalesced key stores */
    for (uint32_t j = tid; j < tileCount; j += SORT_THREADS)
    {
        K k        = keysS[j];
        unsigned d = unsigned((k >> shift) & (RADIX - 1));
        keysOut[uint32_t(binBase[d] + j)] = k; // 32-bit wrap-around is intended
    }

    if constexpr (HAS_VALUES)
    {
        // padding elements of a partial tile carry the largest key, so their ranks lie beyond tileCount
#pragma unroll
        for (int i = 0; i < IPT; ++i)
            valsS[rank[i]] = val[i];
        __syncthreads();
        for (uint32_t j = tid; j < tileCount; j += SORT_THREADS)
        {
            unsigned d = unsigned((keysS[j] >> shift) & (RADIX - 1));
            valsOut[uint32_t(binBase[d] + j)] = valsS[j];
        }
    }
}

template<class K>
size_t sortTempBytes(size_t n)
{
    using Cfg       = SortCfg<K>;
    size_t numTiles = (n + MIN_TILE - 1) / MIN_TILE;
    size_t words    = size_t(Cfg::passes) * RADIX      // digit histograms
                   + 64                                // tile counters (one per pass)
                   + size_t(Cfg::passes) * numTiles * RADIX; // look-back states
    return words * sizeof(uint32_t) + 256;
}

template<class K>
int sortByKey(K* keys, uint32_t* values, size_t n, K* keyBuf, uint32_t* valueBuf, void* tmp, size_t tmpBytes,
              cudaStream_t stream, long long iotaStart = -1, bool skipTrivialPasses = false)
{
    using Cfg = SortCfg<K>;
    if (n < 2) { return 0; }
    CSB_REQUIRE(n < (size_t(1) << 30), "sort_by_key supports fewer than 2^30 elements");
    CSB_REQUIRE(tmpBytes >= sortTempBytes<K>(n), "sort_by_key temp storage too small");
    CSB_REQUIRE(keyBuf != nullptr && (values == nullptr || valueBuf != nullptr), "sort_by_key needs double buffers");

    int dev = 0, numSm = 0;
    CSB_CHECK(cudaGetDevice(&dev));
    CSB_CHECK(cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev));

    uint32_t* hist     = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(tmp) + 255) & ~uintptr_t(255));
    uint32_t* counters = hist + Cfg::passes * RADIX;
    uint32_t* states   = counters + 64;

    unsigned histGrid = unsigned(std::min<size_t>(size_t(numSm) * 4, (n + 511) / 512));
    int status        = 0;
    auto run = [&](auto threadsC, auto iptC, auto minbC)
    {
        constexpr int THREADS = decltype(threadsC)::value;
        constexpr int IPT     = decltype(iptC)::value * (sizeof(K) == 8 ? 3 : 4) / 3; // u32 keys: 4/3 more per thread
        constexpr int MINB    = decltype(minbC)::value;
        constexpr int TILE    = THREADS * IPT;
        constexpr size_t smem = sortSmemBytes<K>(THREADS, IPT);
        auto kv               = onesweepKernel<K, true, THREADS, IPT, MINB>;
        auto ko               = onesweepKernel<K, false, THREADS, IPT, MINB>;
        if (cudaFuncSetAttribute(kv, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess ||
            cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
        {
            setLastError("sort_by_key: cannot configure shared memory");
            status = 1;
            return;
        }
        size_t numTiles  = (n + TILE - 1) / TILE;
        size_t zeroBytes = (size_t(Cfg::passes) * RADIX + 64 + size_t(Cfg::passes) * numTiles * RADIX) * 4;
        if (cudaMemsetAsync(hist, 0, zeroBytes, stream) != cudaSuccess)
        {
            setLastError("sort_by_key: memset failed");
            status = 1;
            return;
        }
        radixHistogramKernel<K><<<histGrid, 512, 0, stream>>>(keys, n, hist);
        countLaunch();
        scanHistogramKernel<<<Cfg::passes, RADIX, 0, stream>>>(hist);
        countLaunch();

        // a digit place in which all keys agree is an identity pass.  Keys with a small range (octree prefixes) have
        // several; finding them costs one read-back of the scanned histograms, so only callers that synchronise anyway
        // ask for it
        bool trivial[Cfg::passes] = {};
        if (skipTrivialPasses)
        {
            uint32_t base[Cfg::passes * RADIX];
            if (cudaMemcpyAsync(base, hist, sizeof(base), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
                cudaStreamSynchronize(stream) != cudaSuccess)
            {
                setLastError("sort_by_key: histogram read-back failed");
                status = 1;
                return;
            }
            for (int p = 0; p < Cfg::passes; ++p)
                for (int b = 0; b < RADIX; ++b)
                {
                    uint64_t next = b + 1 < RADIX ? base[p * RADIX + b + 1] : n;
                    if (next - base[p * RADIX + b] == n) { trivial[p] = true; }
                }
        }

        K* kin         = keys;
        K* kout        = keyBuf;
        uint32_t* vin  = values;
        uint32_t* vout = valueBuf;
        bool iotaDone  = iotaStart < 0;
        for (int p = 0; p < Cfg::passes; ++p)
        {
            if (trivial[p]) { continue; }
            const long long iota = iotaDone ? -1LL : iotaStart;
            iotaDone             = true;
            uint32_t* st = states + size_t(p) * numTiles * RADIX;
            if (values)
            {
                kv<<<unsigned(numTiles), THREADS, smem, stream>>>(kin, kout, vin, vout, n, p * RADIX_BITS,
                                                                  hist + p * RADIX, st, counters + p,
                                                                  g_sortDebugNoLookback, iota);
            }
            else
            {
                ko<<<unsigned(numTiles), THREADS, smem, stream>>>(kin, kout, nullptr, nullptr, n, p * RADIX_BITS,
                                                                  hist + p * RADIX, st, counters + p,
                                                                  g_sortDebugNoLookback, -1LL);
            }
            countLaunch();
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        if (!iotaDone && values) // every pass was trivial: the keys are constant, the permutation is the identity
        {
            if (cs_sequence_u32(uint32_t(iotaStart), n, values, stream) != 0) { status = 1; }
        }
        if (kin != keys) // an odd number of passes ran: the result sits in the double buffers
        {
            bool ok = cudaMemcpyAsync(keys, kin, n * sizeof(K), cudaMemcpyDeviceToDevice, stream) == cudaSuccess;
            if (values)
            {
                ok = ok && cudaMemcpyAsync(values, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) ==
                               cudaSuccess;
            }
            if (!ok)
            {
                setLastError("sort_by_key: copy back failed");
                status = 1;
            }
        }
    };
    using std::integral_constant;
    switch (g_sortVariant)
    {
        case 1: run(integral_constant<int, 256>{}, integral_constant<int, 12>{}, integral_constant<int, 4>{}); break;
        case 2: run(integral_constant<int, 256>{}, integral_constant<int, 15>{}, integral_constant<int, 3>{}); break;
        case 3: run(integral_constant<int, 384>{}, integral_constant<int, 12>{}, integral_constant<int, 3>{}); break;
        case 4: run(integral_constant<int, 256>{}, integral_constant<int, 9>{}, integral_constant<int, 5>{}); break;
        case 5: run(integral_constant<int, 512>{}, integral_constant<int, 15>{}, integral_constant<int, 2>{}); break;
        case 6: run(integral_constant<int, 384>{}, integral_constant<int, 18>{}, integral_constant<int, 2>{}); break;
        default: run(integral_constant<int, 512>{}, integral_constant<int, 12>{}, integral_constant<int, 2>{}); break;
    }
    if (status) { return status; }
    CSB_CHECK(cudaGetLastError());
    // passes is even for both key widths, so the sorted data is back in keys/values (skipped passes: copied back)
    static_assert(Cfg::passes % 2 == 0);
    return 0;
}

/* ---------------------------------------------------------------- sequence / gather */

__global__ void sequenceKernel(uint32_t start, size_t n, uint32_t* out)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = start + uint32_t(i); }
}

template<class E>
__global__ void gatherKernel(const uint32_t* __restrict__ ord, size_t n, const E* __restrict__ src, E* __restrict__ dst)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { dst[i] = src[ord[i]]; }
}

template<class E>
struct Ptr4
{
    const E* src[4];
    E* dst[4];
};

template<class E>
__global__ void gather4Kernel(const uint32_t* __restrict__ ord, size_t n, Ptr4<E> p)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
    {
        uint32_t o = ord[i];
        E a = p.src[0][o], b = p.src[1][o], c = p.src[2][o], d = p.src[3][o];
        p.dst[0][i] = a;
        p.dst[1][i] = b;
        p.dst[2][i] = c;
        p.dst[3][i] = d;
    }
}

/* Random gathers of 8-byte elements cost a 128-byte DRAM fetch each on B200 (measured: 520 B read per particle for four
 * arrays, independent of cudaLimitMaxL2FetchGranularity).  gatherArrays therefore first interleaves the four source
 * arrays into one array of 4-element records with fully coalesced traffic and then gathers whole records: one
 * 16/32-byte sector-aligned random read per particle instead of four scattered ones. */
template<class E>
struct alignas(4 * sizeof(E)) Rec4
{
    E v[4];
};

template<class E>
__global__ void packRec4Kernel(size_t n, const E* __restrict__ a, const E* __restrict__ b, const E* __restrict__ c,
                               const E* __restrict__ d, Rec4<E>* __restrict__ rec)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
    {
        Rec4<E> r;
        r.v[0] = a[i];
        r.v[1] = b[i];
        r.v[2] = c[i];
        r.v[3] = d[i];
        rec[i] = r;
    }
}

//! record load with the L2 fetch-size hint 64 B: a missing sector then costs 64 B of DRAM traffic instead of the whole
//! 128-byte line that the default policy brings in (LDG.LTC64B; measured on the random gather: see profiles/)
template<class E>
__device__ __forceinline__ Rec4<E> loadRecord64B(const Rec4<E>* p)
{
    Rec4<E> r;
    if constexpr (sizeof(E) == 8)
    {
        uint64_t a, b, c, d;
        asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2+16];" : "=l"(c), "=l"(d) : "l"(p));
        uint64_t w[4] = {a, b, c, d};
        memcpy(&r, w, sizeof(r)); // E is whatever 8-byte type the caller moves; keep the bits
    }
    else
    {
        uint32_t a, b, c, d;
        asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
        uint32_t w[4] = {a, b, c, d};
        memcpy(&r, w, sizeof(r));
    }
    return r;
}

int g_gatherVariant = 1; // 0: plain 256-bit record loads, 1: L2::64B fetch hint

template<class E, int VARIANT>
__global__ void gatherRec4Kernel(const uint32_t* __restrict__ ord, size_t n, const Rec4<E>* __restrict__ rec, E* __restrict__ a,
                                 E* __restrict__ b, E* __restrict__ c, E* __restrict__ d)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
    {
        Rec4<E> r;
        if constexpr (VARIANT == 1) { r = loadRecord64B(rec + ord[i]); }
        else { r = rec[ord[i]]; }
        a[i] = r.v[0];
        b[i] = r.v[1];
        c[i] = r.v[2];
        d[i] = r.v[3];
    }
}

template<class E>
int gatherArrays4(const uint32_t* ordering, size_t n, size_t srcCount, const void* const* src4, void* const* dst4,
                  cudaStream_t s)
{
    CSB_SCRATCH(rec, Rec4<E>*, s, SCRATCH_D, srcCount * sizeof(Rec4<E>));
    packRec4Kernel<E><<<iceil(srcCount, 256), 256, 0, s>>>(srcCount, static_cast<const E*>(src4[0]),
                                                           static_cast<const E*>(src4[1]),
                                                           static_cast<const E*>(src4[2]),
                                                           static_cast<const E*>(src4[3]), rec);
    CSB_LAUNCH_CHECK();
    auto kernel = g_gatherVariant == 1 ? gatherRec4Kernel<E, 1> : gatherRec4Kernel<E, 0>;
    kernel<<<iceil(n, 256), 256, 0, s>>>(ordering, n, rec, static_cast<E*>(dst4[0]), static_cast<E*>(dst4[1]),
                                         static_cast<E*>(dst4[2]), static_cast<E*>(dst4[3]));
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class E>
__global__ void packRec4RangeKernel(size_t first, size_t n, const E* __restrict__ a, const E* __restrict__ b,
                                    const E* __restrict__ c, const E* __restrict__ d, Rec4<E>* __restrict__ rec)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
    {
        Rec4<E> r;
        r.v[0]         = a[first + i];
        r.v[1]         = b[first + i];
        r.v[2]         = c[first + i];
        r.v[3]         = d[first + i];
        rec[first + i] = r;
    }
}

} // namespace

/*! records (x,y,z,h) of the elements [first, first + n) of four arrays, stored at the same positions of `rec`: the
 *  particle exchange and the gather that follows read whole 32-byte records through the ordering instead of four
 *  scattered elements (a random 8-byte read costs a 128-byte DRAM fetch, see gatherArrays4) */
int packRecords4(const void* const* src4, size_t first, size_t n, void* rec, int elemBytes, cudaStream_t s)
{
    if (n == 0) { return 0; }
    if (elemBytes == 8)
    {
        packRec4RangeKernel<uint64_t><<<iceil(n, 256), 256, 0, s>>>(
            first, n, static_cast<const uint64_t*>(src4[0]), static_cast<const uint64_t*>(src4[1]),
            static_cast<const uint64_t*>(src4[2]), static_cast<const uint64_t*>(src4[3]),
            static_cast<Rec4<uint64_t>*>(rec));
    }
    else
    {
        packRec4RangeKernel<uint32_t><<<iceil(n, 256), 256, 0, s>>>(
            first, n, static_cast<const uint32_t*>(src4[0]), static_cast<const uint32_t*>(src4[1]),
            static_cast<const uint32_t*>(src4[2]), static_cast<const uint32_t*>(src4[3]),
            static_cast<Rec4<uint32_t>*>(rec));
    }
    CSB_LAUNCH_CHECK();
    return 0;
}

//! dst4[k][i] = rec[ordering[i]].v[k]; the destination arrays may be peer memory
int gatherFromRecords4(const uint32_t* ordering, size_t n, const void* rec, void* const* dst4, int elemBytes,
                       cudaStream_t s)
{
    if (n == 0) { return 0; }
    if (elemBytes == 8)
    {
        gatherRec4Kernel<uint64_t, 1><<<iceil(n, 256), 256, 0, s>>>(
            ordering, n, static_cast<const Rec4<uint64_t>*>(rec), static_cast<uint64_t*>(dst4[0]),
            static_cast<uint64_t*>(dst4[1]), static_cast<uint64_t*>(dst4[2]), static_cast<uint64_t*>(dst4[3]));
    }
    else
    {
        gatherRec4Kernel<uint32_t, 1><<<iceil(n, 256), 256, 0, s>>>(
            ordering, n, static_cast<const Rec4<uint32_t>*>(rec), static_cast<uint32_t*>(dst4[0]),
            static_cast<uint32_t*>(dst4[1]), static_cast<uint32_t*>(dst4[2]), static_cast<uint32_t*>(dst4[3]));
    }
    CSB_LAUNCH_CHECK();
    return 0;
}

int sortByKeyU64(uint64_t* keys, uint32_t* values, size_t n, uint64_t* keyBuf, uint32_t* valueBuf, void* tmp,
                 size_t tmpBytes, cudaStream_t stream)
{
    return sortByKey<uint64_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream);
}

int sortByKeyU32(uint32_t* keys, uint32_t* values, size_t n, uint32_t* keyBuf, uint32_t* valueBuf, void* tmp,
                 size_t tmpBytes, cudaStream_t stream)
{
    return sortByKey<uint32_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream);
}

//! sort keys and produce the sorting permutation of the sequence first, first + 1, ... (values need no initialisation)
int sortByKeyIotaU64(uint64_t* keys, uint32_t* values, uint32_t first, size_t n, uint64_t* keyBuf, uint32_t* valueBuf,
                     void* tmp, size_t tmpBytes, cudaStream_t stream)
{
    if (n == 1) { return cs_sequence_u32(first, 1, values, stream); }
    return sortByKey<uint64_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream, (long long)first);
}

int sortByKeyIotaU32(uint32_t* keys, uint32_t* values, uint32_t first, size_t n, uint32_t* keyBuf, uint32_t* valueBuf,
                     void* tmp, size_t tmpBytes, cudaStream_t stream)
{
    if (n == 1) { return cs_sequence_u32(first, 1, values, stream); }
    return sortByKey<uint32_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream, (long long)first);
}

//! sort for callers that synchronise the stream anyway: identity passes (a digit place shared by all keys) are skipped
int sortByKeySkipU64(uint64_t* keys, uint32_t* values, size_t n, uint64_t* keyBuf, uint32_t* valueBuf, void* tmp,
                     size_t tmpBytes, cudaStream_t stream)
{
    return sortByKey<uint64_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream, -1, true);
}

int sortByKeySkipU32(uint32_t* keys, uint32_t* values, size_t n, uint32_t* keyBuf, uint32_t* valueBuf, void* tmp,
                     size_t tmpBytes, cudaStream_t stream)
{
    return sortByKey<uint32_t>(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, stream, -1, true);
}

void setSortVariant(int v)
{
    g_sortVariant          = v % 100;
    g_sortDebugNoLookback = v >= 100;
}
size_t sortTempBytesU64(size_t n) { return sortTempBytes<uint64_t>(n); }
size_t sortTempBytesU32(size_t n) { return sortTempBytes<uint32_t>(n); }

} // namespace csb

extern "C"
{

/* tuning hook (not part of the drop-in surface): select the onesweep kernel variant; 1000 + v selects the record
 * load of gatherArrays (0: plain, 1: L2 64-byte fetch hint) */
int cs_sort_set_variant(int variant)
{
    if (variant >= 1000) { csb::g_gatherVariant = variant - 1000; }
    else { csb::setSortVariant(variant); }
    return 0;
}

size_t cs_sort_by_key_temp_bytes_u32(size_t n) { return csb::sortTempBytesU32(n); }
size_t cs_sort_by_key_temp_bytes_u64(size_t n) { return csb::sortTempBytesU64(n); }

int cs_sort_by_key_u32(uint32_t* keys, uint32_t* values, size_t n, uint32_t* keyBuf, uint32_t* valueBuf, void* tmp,
                       size_t tmpBytes, void* stream)
{
    return csb::sortByKeyU32(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, cudaStream_t(stream));
}

int cs_sort_by_key_u64(uint64_t* keys, uint32_t* values, size_t n, uint64_t* keyBuf, uint32_t* valueBuf, void* tmp,
                       size_t tmpBytes, void* stream)
{
    return csb::sortByKeyU64(keys, values, n, keyBuf, valueBuf, tmp, tmpBytes, cudaStream_t(stream));
}

int cs_sequence_u32(uint32_t start, size_t n, uint32_t* out, void* stream)
{
    if (n == 0) { return 0; }
    csb::sequenceKernel<<<csb::iceil(n, 256), 256, 0, cudaStream_t(stream)>>>(start, n, out);
    CSB_LAUNCH_CHECK();
    return 0;
}

int cs_gather(const uint32_t* ordering, size_t n, const void* src, void* dst, int elemBytes, void* stream)
{
    CSB_REQUIRE(elemBytes == 4 || elemBytes == 8, "gather supports 4- and 8-byte elements");
    if (n == 0) { return 0; }
    if (elemBytes == 4)
    {
        csb::gatherKernel<uint32_t><<<csb::iceil(n, 256), 256, 0, cudaStream_t(stream)>>>(
            ordering, n, static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst));
    }
    else
    {
        csb::gatherKernel<uint64_t><<<csb::iceil(n, 256), 256, 0, cudaStream_t(stream)>>>(
            ordering, n, static_cast<const uint64_t*>(src), static_cast<uint64_t*>(dst));
    }
    CSB_LAUNCH_CHECK();
    return 0;
}

int cs_gather4(const uint32_t* ordering, size_t n, const void* const* src4, void* const* dst4, int elemBytes,
               void* stream)
{
    CSB_REQUIRE(elemBytes == 4 || elemBytes == 8, "gather supports 4- and 8-byte elements");
    if (n == 0) { return 0; }
    if (elemBytes == 4)
    {
        csb::Ptr4<uint32_t> p;
        for (int i = 0; i < 4; ++i)
        {
            p.src[i] = static_cast<const uint32_t*>(src4[i]);
            p.dst[i] = static_cast<uint32_t*>(dst4[i]);
        }
        csb::gather4Kernel<uint32_t><<<csb::iceil(n, 256), 256, 0, cudaStream_t(stream)>>>(ordering, n, p);
    }
    else
    {
        csb::Ptr4<uint64_t> p;
        for (int i = 0; i < 4; ++i)
        {
            p.src[i] = static_cast<const uint64_t*>(src4[i]);
            p.dst[i] = static_cast<uint64_t*>(dst4[i]);
        }
        csb::gather4Kernel<uint64_t><<<csb::iceil(n, 256), 256, 0, cudaStream_t(stream)>>>(ordering, n, p);
    }
    CSB_LAUNCH_CHECK();
    return 0;
}

int cs_gather_arrays4(const uint32_t* ordering, size_t n, size_t srcCount, const void* const* src4, void* const* dst4,
                      int elemBytes, void* stream)
{
    CSB_REQUIRE(elemBytes == 4 || elemBytes == 8, "gather supports 4- and 8-byte elements");
    if (n == 0) { return 0; }
    // small inputs are launch-latency bound: one direct pass
    if (srcCount < (size_t(1) << 16)) { return cs_gather4(ordering, n, src4, dst4, elemBytes, stream); }
    if (elemBytes == 4) { return csb::gatherArrays4<uint32_t>(ordering, n, srcCount, src4, dst4, cudaStream_t(stream)); }
    return csb::gatherArrays4<uint64_t>(ordering, n, srcCount, src4, dst4, cudaStream_t(stream));
}

} // extern "C"
