/* Exclusive prefix sum over uint32 arrays (replaces thrust::exclusive_scan in the reference,
 * primitives/primitives_gpu.cu:125-140).  Three phases: per-CTA sums, scan of the CTA sums, per-CTA rescan with
 * offset.  Arrays on this path are leaf/node sized (<= a few million), far from HBM-bound. */
#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_IPT     = 8;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_IPT;

//! CTA-wide inclusive scan of one value per thread; returns inclusive value, total in *total
__device__ inline uint32_t blockInclusiveScan(uint32_t v, uint32_t* smem /* >= 33 words */, uint32_t* total)
{
    unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= unsigned(o)) { incl += t; }
    }
    __syncthreads(); // protect smem reuse across calls
    if (lane == 31) { smem[warp] = incl; }
    __syncthreads();
    unsigned numWarps = blockDim.x >> 5;
    if (warp == 0)
    {
        uint32_t w  = lane < numWarps ? smem[lane] : 0;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= unsigned(o)) { wi += t; }
        }
        smem[lane] = wi - w; // exclusive warp offsets
        if (lane == 31) { smem[32] = wi; }
    }
    __syncthreads();
    *total = smem[32];
    return incl + smem[warp];
}

__global__ void __launch_bounds__(SCAN_THREADS) scanBlockSumsKernel(const uint32_t* __restrict__ in, size_t n,
                                                                    uint32_t* blockSums)
{
    __shared__ uint32_t smem[33];
    size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_IPT;
    uint32_t s  = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i)
        if (base + i < n) { s += in[base + i]; }
    uint32_t total;
    blockInclusiveScan(s, smem, &total);
    if (threadIdx.x == 0) { blockSums[blockIdx.x] = total; }
}

__global__ void __launch_bounds__(SCAN_THREADS) scanSumsKernel(uint32_t* blockSums, unsigned numBlocks)
{
    __shared__ uint32_t smem[33];
    uint32_t carry = 0;
    for (unsigned base = 0; base < numBlocks; base += SCAN_THREADS)
    {
        unsigned i   = base + threadIdx.x;
        uint32_t v   = i < numBlocks ? blockSums[i] : 0;
        uint32_t total;
        uint32_t incl = blockInclusiveScan(v, smem, &total);
        if (i < numBlocks) { blockSums[i] = carry + incl - v; }
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scanDownsweepKernel(const uint32_t* in, uint32_t* out, size_t n,
                                                                    const uint32_t* __restrict__ blockSums)
{
    __shared__ uint32_t smem[33];
    size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_IPT;
    uint32_t v[SCAN_IPT];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i)
    {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t incl = blockInclusiveScan(s, smem, &total);
    uint32_t run  = blockSums[blockIdx.x] + incl - s;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i)
    {
        if (base + i < n) { out[base + i] = run; }
        run += v[i];
    }
}

} // namespace

size_t scanTempBytes(size_t n) { return (size_t(iceil(n, SCAN_TILE)) + 1) * sizeof(uint32_t) + 256; }

int exclusiveScanU32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, cudaStream_t s)
{
    if (n == 0) { return 0; }
    uint32_t* blockSums = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(tmp) + 255) & ~uintptr_t(255));
    unsigned numBlocks  = iceil(n, SCAN_TILE);
    scanBlockSumsKernel<<<numBlocks, SCAN_THREADS, 0, s>>>(in, n, blockSums);
    CSB_LAUNCH_CHECK();
    scanSumsKernel<<<1, SCAN_THREADS, 0, s>>>(blockSums, numBlocks);
    CSB_LAUNCH_CHECK();
    scanDownsweepKernel<<<numBlocks, SCAN_THREADS, 0, s>>>(in, out, n, blockSums);
    CSB_LAUNCH_CHECK();
    return 0;
}

} // namespace csb

extern "C"
{

size_t cs_scan_temp_bytes(size_t n) { return csb::scanTempBytes(n); }

int cs_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, void* stream)
{
    return csb::exclusiveScanU32(in, out, n, tmp, cudaStream_t(stream));
}

} // extern "C"
