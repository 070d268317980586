/* Halo discovery for sm_100a: per-leaf search boxes and the collision traversal that flags foreign tree nodes.
 * Replaces computeBoundingBoxGpu (reference focus/source_center_gpu.cu:23-92, arithmetic focus/source_center.hpp:28-43)
 * and findHalosGpu (traversal/collisions_gpu.cu:23-88, arithmetic traversal/collisions.hpp:25-94 and
 * traversal/boxoverlap.hpp:127-152,280-291).
 */
#include "common.cuh"
#include "hilbert.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

/* ---------------------------------------------------------------- bounding boxes: 8 lanes per leaf */

template<class T>
__device__ inline T shflXor(T v, int m)
{
    return __shfl_xor_sync(0xffffffffu, v, m);
}

//! T: coordinate type, Th: type of the smoothing lengths and of `scale` (the reference instantiates (double, double),
//! (double, float) and (float, float), focus/source_center_gpu.cu:90-92): r = h * scale is formed in Th and promoted
template<class T, class Th>
__global__ void __launch_bounds__(256) boundingBoxKernel(const T* __restrict__ x,
                                                         const T* __restrict__ y,
                                                         const T* __restrict__ z,
                                                         const Th* __restrict__ h,
                                                         const uint32_t* __restrict__ layout,
                                                         int firstLeaf,
                                                         int lastLeaf,
                                                         Th scale,
                                                         T* __restrict__ sc,
                                                         T* __restrict__ ss)
{
    constexpr int G = 8;
    size_t tid      = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    int leaf        = firstLeaf + int(tid / G);
    int sub         = int(tid % G);
    bool active     = leaf < lastLeaf;
    int l           = active ? leaf : lastLeaf - 1;

    T mn[3] = {sc[3 * l], sc[3 * l + 1], sc[3 * l + 2]};
    T mx[3] = {mn[0], mn[1], mn[2]};
    uint32_t jb = layout[l], je = layout[l + 1];
    for (uint32_t j = jb + sub; j < je; j += G)
    {
        const Th rh = h[j] * scale;
        const T r   = T(rh);
        T p[3]      = {x[j], y[j], z[j]};
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            T lo = p[d] - r, hi = p[d] + r;
            mn[d] = lo < mn[d] ? lo : mn[d];
            mx[d] = hi > mx[d] ? hi : mx[d];
        }
    }
#pragma unroll
    for (int m = 1; m < G; m <<= 1)
    {
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            T a   = shflXor(mn[d], m);
            T b   = shflXor(mx[d], m);
            mn[d] = a < mn[d] ? a : mn[d];
            mx[d] = b > mx[d] ? b : mx[d];
        }
    }
    if (active && sub == 0)
    {
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            sc[3 * l + d] = (mx[d] + mn[d]) * T(0.5);
            ss[3 * l + d] = (mx[d] - mn[d]) * T(0.5);
        }
    }
}

/* ---------------------------------------------------------------- collisions */

template<class K, class T>
__device__ inline K sfc3DHilbert(T x, T y, T z, const Box<T>& box, const unsigned char* hilbertTables)
{
    constexpr unsigned cubeLength = 1u << KeyTraits<K>::maxLevel;
    constexpr int mcoord          = int(cubeLength - 1);
    T mx = cubeLength * box.ilen[0], my = cubeLength * box.ilen[1], mz = cubeLength * box.ilen[2];
    int ix = int(rfloor(x * mx) - box.lim[0] * mx);
    int iy = int(rfloor(y * my) - box.lim[2] * my);
    int iz = int(rfloor(z * mz) - box.lim[4] * mz);
    ix     = min(ix, mcoord);
    iy     = min(iy, mcoord);
    iz     = min(iz, mcoord);
    return hilbertEncode<K>(unsigned(ix), unsigned(iy), unsigned(iz), hilbertTables);
}

//! traversal/boxoverlap.hpp:127-152
template<class K, class T>
__device__ inline bool containedIn(K codeStart, K codeEnd, const T* c, const T* s, const Box<T>& box,
                                   const unsigned char* hilbertTables)
{
    T bmin[3], bmax[3];
    T dFromMin = 0, dFromMax = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        bmin[d] = c[d] - s[d];
        bmax[d] = c[d] + s[d];
        T a     = bmin[d] - box.lim[2 * d];
        T b     = bmax[d] - box.lim[2 * d + 1];
        dFromMin = (d == 0 || a < dFromMin) ? a : dFromMin;
        dFromMax = (d == 0 || b > dFromMax) ? b : dFromMax;
    }
    if (dFromMin < T(0) || dFromMax > T(0)) { return codeStart == 0 && codeEnd == nodeRange<K>(0); }

    constexpr int gridDim_ = 1 << KeyTraits<K>::maxLevel;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        bmax[d] += box.len[d] * (T(1) / gridDim_);

    K lowCode      = sfc3DHilbert<K>(bmin[0], bmin[1], bmin[2], box, hilbertTables);
    K highCode     = sfc3DHilbert<K>(bmax[0], bmax[1], bmax[2], box, hilbertTables);
    unsigned level = unsigned(commonPrefix(lowCode, highCode)) / 3;
    K nodeStart    = lowCode & ~(nodeRange<K>(level) - 1);
    K nodeEnd      = nodeStart + nodeRange<K>(level);
    return nodeStart >= codeStart && nodeEnd <= codeEnd;
}

//! traversal/boxoverlap.hpp:280-291
template<class T>
__device__ inline bool boxOverlap(const T* ac, const T* as, const T* bc, const T* bs, const Box<T>& box)
{
    bool ret = true;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        T dx = rabs(pbcFold(bc[d] - ac[d], d, box));
        dx -= as[d];
        dx -= bs[d];
        ret = ret && (dx < T(0));
    }
    return ret;
}

/*! one thread per own leaf (only leaves whose search box leaves the own SFC range do any work), stackless walk.
 *  Flag stores race benignly: every writer stores 1. */
template<class K, class T>
__global__ void __launch_bounds__(128) findHalosKernel(const K* __restrict__ prefixes,
                                                       const int* __restrict__ childOffsets,
                                                       const int* __restrict__ parents,
                                                       const T* __restrict__ centers,
                                                       const T* __restrict__ sizes,
                                                       const K* __restrict__ leaves,
                                                       const T* __restrict__ searchCenters,
                                                       const T* __restrict__ searchSizes,
                                                       Box<T> box,
                                                       int firstLeaf,
                                                       int lastLeaf,
                                                       uint8_t* flags)
{
    __shared__ unsigned char hilbertTables[hilbertTableBytes];
    stageHilbertTables(hilbertTables);
    __syncthreads();
    int leaf = firstLeaf + blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= lastLeaf) { return; }

    K lowestKey  = leaves[firstLeaf];
    K highestKey = leaves[lastLeaf];
    T tc[3]      = {searchCenters[3 * leaf], searchCenters[3 * leaf + 1], searchCenters[3 * leaf + 2]};
    T ts[3]      = {searchSizes[3 * leaf], searchSizes[3 * leaf + 1], searchSizes[3 * leaf + 2]};

    if (containedIn(lowestKey, highestKey, tc, ts, box, hilbertTables)) { return; }

    auto overlaps = [&](int idx)
    {
        K prefix           = prefixes[idx];
        unsigned prefixLen = decodePrefixLength(prefix);
        K nk1              = decodePlaceholderBit(prefix);
        K nk2              = nk1 + (K(1) << (3 * KeyTraits<K>::maxLevel - prefixLen));
        bool contained     = !(nk1 < lowestKey || nk2 > highestKey);
        if (contained) { return false; }
        T nc[3] = {centers[3 * idx], centers[3 * idx + 1], centers[3 * idx + 2]};
        T ns[3] = {sizes[3 * idx], sizes[3 * idx + 1], sizes[3 * idx + 2]};
        bool ov = boxOverlap(nc, ns, tc, ts, box);
        if (ov) { flags[idx] = 1; }
        return ov;
    };

    if (!overlaps(0)) { return; }
    int node = childOffsets[0];
    if (node == 0) { return; }
    bool backtrack = false;
    while (node != 0)
    {
        int child    = childOffsets[node];
        bool isLeaf  = child == 0;
        bool descend = !backtrack && overlaps(node);
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

} // namespace

template<class T, class Th>
int computeBoundingBoxes(const T* x, const T* y, const T* z, const Th* h, const uint32_t* layout, int firstLeaf,
                         int lastLeaf, Th scale, T* sc, T* ss, cudaStream_t s)
{
    if (lastLeaf <= firstLeaf) { return 0; }
    size_t threads = size_t(lastLeaf - firstLeaf) * 8;
    boundingBoxKernel<T, Th><<<iceil(threads, 256), 256, 0, s>>>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, sc, ss);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K, class T>
int findHalos(const K* prefixes, const int* childOffsets, const int* parents, const T* centers, const T* sizes,
              const K* leaves, const T* searchCenters, const T* searchSizes, const double* lim, const int* bnd,
              int firstLeaf, int lastLeaf, uint8_t* flags, cudaStream_t s)
{
    if (lastLeaf <= firstLeaf) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);
    findHalosKernel<K, T><<<iceil(lastLeaf - firstLeaf, 128), 128, 0, s>>>(
        prefixes, childOffsets, parents, centers, sizes, leaves, searchCenters, searchSizes, box, firstLeaf, lastLeaf,
        flags);
    CSB_LAUNCH_CHECK();
    return 0;
}

template int computeBoundingBoxes<float, float>(const float*, const float*, const float*, const float*, const uint32_t*,
                                                int, int, float, float*, float*, cudaStream_t);
template int computeBoundingBoxes<double, double>(const double*, const double*, const double*, const double*,
                                                  const uint32_t*, int, int, double, double*, double*, cudaStream_t);
template int computeBoundingBoxes<double, float>(const double*, const double*, const double*, const float*,
                                                 const uint32_t*, int, int, float, double*, double*, cudaStream_t);
template int findHalos<uint32_t, float>(const uint32_t*, const int*, const int*, const float*, const float*,
                                        const uint32_t*, const float*, const float*, const double*, const int*, int,
                                        int, uint8_t*, cudaStream_t);
template int findHalos<uint64_t, float>(const uint64_t*, const int*, const int*, const float*, const float*,
                                        const uint64_t*, const float*, const float*, const double*, const int*, int,
                                        int, uint8_t*, cudaStream_t);
template int findHalos<uint64_t, double>(const uint64_t*, const int*, const int*, const double*, const double*,
                                         const uint64_t*, const double*, const double*, const double*, const int*, int,
                                         int, uint8_t*, cudaStream_t);

} // namespace csb

extern "C"
{

int cs_compute_bounding_boxes_f(const float* x, const float* y, const float* z, const float* h, const uint32_t* layout,
                                int firstLeaf, int lastLeaf, float scale, float* searchCenters, float* searchSizes,
                                void* stream)
{
    return csb::computeBoundingBoxes<float, float>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, searchCenters, searchSizes,
                                            cudaStream_t(stream));
}
int cs_compute_bounding_boxes_d(const double* x, const double* y, const double* z, const double* h,
                                const uint32_t* layout, int firstLeaf, int lastLeaf, double scale,
                                double* searchCenters, double* searchSizes, void* stream)
{
    return csb::computeBoundingBoxes<double, double>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, searchCenters,
                                             searchSizes, cudaStream_t(stream));
}

int cs_compute_bounding_boxes_df(const double* x, const double* y, const double* z, const float* h, const uint32_t* layout,
                                 int firstLeaf, int lastLeaf, float scale, double* searchCenters, double* searchSizes,
                                 void* stream)
{
    return csb::computeBoundingBoxes<double, float>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, searchCenters,
                                                    searchSizes, cudaStream_t(stream));
}

int cs_find_halos_u32f(const uint32_t* prefixes, const int* childOffsets, const int* parents, const float* centers,
                       const float* sizes, const uint32_t* leaves, const float* searchCenters, const float* searchSizes,
                       const double* lim, const int* bnd, int firstLeaf, int lastLeaf, uint8_t* flags, void* stream)
{
    return csb::findHalos<uint32_t, float>(prefixes, childOffsets, parents, centers, sizes, leaves, searchCenters,
                                           searchSizes, lim, bnd, firstLeaf, lastLeaf, flags, cudaStream_t(stream));
}
int cs_find_halos_u64f(const uint64_t* prefixes, const int* childOffsets, const int* parents, const float* centers,
                       const float* sizes, const uint64_t* leaves, const float* searchCenters, const float* searchSizes,
                       const double* lim, const int* bnd, int firstLeaf, int lastLeaf, uint8_t* flags, void* stream)
{
    return csb::findHalos<uint64_t, float>(prefixes, childOffsets, parents, centers, sizes, leaves, searchCenters,
                                           searchSizes, lim, bnd, firstLeaf, lastLeaf, flags, cudaStream_t(stream));
}
int cs_find_halos_u64d(const uint64_t* prefixes, const int* childOffsets, const int* parents, const double* centers,
                       const double* sizes, const uint64_t* leaves, const double* searchCenters,
                       const double* searchSizes, const double* lim, const int* bnd, int firstLeaf, int lastLeaf,
                       uint8_t* flags, void* stream)
{
    return csb::findHalos<uint64_t, double>(prefixes, childOffsets, parents, centers, sizes, leaves, searchCenters,
                                            searchSizes, lim, bnd, firstLeaf, lastLeaf, flags, cudaStream_t(stream));
}

} // extern "C"
