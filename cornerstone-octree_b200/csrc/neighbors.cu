/* Radius neighbour search for sm_100a producing the reference's CPU list layout:
 * cstone::findNeighbors (findneighbors.hpp:77-177): neighbors[(i-first)*ngmax + k], neighborsCount[i-first].
 *
 * Design
 *  - Target groups are LEAF-ALIGNED: every tree leaf is cut into ceil(count/32) equal groups of at most 32 consecutive
 *    particles (the role of the reference's GroupView / computeFixedGroups, traversal/groups.hpp:28-64, but aligned to
 *    leaves).  All targets of a group sit in one leaf cell, so the union of the tree cells their search spheres touch
 *    is close to what a single target touches (27 cells instead of ~85 for arbitrary 32-particle SFC slices).
 *  - One warp owns one group (lanes = targets) and walks the octree ONCE for all of them with a warp-uniform, stackless
 *    depth-first traversal (child / next sibling / parent links as in traversal/traversal.hpp:26-69).  Each lane keeps
 *    the exact pruning state of the reference's per-particle walk: a bit per tree depth says whether this lane's own
 *    continuation test (point-to-cell min distance < (2h)^2, boxoverlap.hpp:229-250) passed on the current root path.
 *    The warp descends while any lane passes; at a leaf only lanes whose own path passed accept candidates, which are
 *    read with warp-uniform (broadcast) loads.
 *  - Leaves are reached in SFC order, so every lane appends neighbours in ascending particle index exactly like the CPU
 *    walk — truncation at ngmax keeps the same entries — and the distance arithmetic is the reference's, operation by
 *    operation (no FMA contraction; norm2 is the right fold x*x + (y*y + z*z), util/array.hpp:236-240; distanceSq is
 *    (x*x + y*y) + z*z, findneighbors.hpp:33-60).  Self exclusion is by index (j != i) as on the CPU (hazard H2).
 */
#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

constexpr int NB_THREADS = 128;

/* ---------------------------------------------------------------- leaf-aligned target groups */

__device__ inline void leafTargets(const uint32_t* __restrict__ layout, int leaf, uint32_t first, uint32_t last,
                                   uint32_t& s, uint32_t& e)
{
    s = max(layout[leaf], first);
    e = min(layout[leaf + 1], last);
    if (e < s) { e = s; }
}

__global__ void groupCountKernel(const uint32_t* __restrict__ layout, int numLeaves, uint32_t first, uint32_t last,
                                 uint32_t* __restrict__ groupCounts)
{
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf > numLeaves) { return; }
    uint32_t c = 0;
    if (leaf < numLeaves)
    {
        uint32_t s, e;
        leafTargets(layout, leaf, first, last, s, e);
        c = (e - s + 31) / 32;
    }
    groupCounts[leaf] = c;
}

__global__ void groupFillKernel(const uint32_t* __restrict__ layout, int numLeaves, uint32_t first, uint32_t last,
                                const uint32_t* __restrict__ groupOffsets, uint2* __restrict__ groups)
{
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= numLeaves) { return; }
    uint32_t s, e;
    leafTargets(layout, leaf, first, last, s, e);
    uint32_t c = e - s;
    if (c == 0) { return; }
    uint32_t ng   = (c + 31) / 32;
    uint32_t size = (c + ng - 1) / ng; // balanced split, <= 32
    uint32_t off  = groupOffsets[leaf];
    for (uint32_t k = 0; k < ng; ++k)
        groups[off + k] = make_uint2(s + k * size, min(s + (k + 1) * size, e));
}

/* ---------------------------------------------------------------- traversal */

template<class T>
struct Target
{
    T x, y, z;
    T radiusSq;
    bool usePbc;
};

//! one staged candidate; 4 x T so that (x,y) and (z,pad) are each one vector LDS
template<class T>
struct alignas(4 * sizeof(T)) Staged
{
    T x, y, z, pad;
};

//! continuation test of findneighbors.hpp:108-112
template<bool PBC, class T>
__device__ inline bool cellOverlap(const Target<T>& t, const T* __restrict__ centers, const T* __restrict__ sizes,
                                   int node, const Box<T>& box)
{
    T cx = centers[3 * node], cy = centers[3 * node + 1], cz = centers[3 * node + 2];
    T sx = sizes[3 * node], sy = sizes[3 * node + 1], sz = sizes[3 * node + 2];
    T dx, dy, dz;
    if (PBC && t.usePbc)
    {
        dx = rabs(pbcFold(cx - t.x, 0, box)) - sx;
        dy = rabs(pbcFold(cy - t.y, 1, box)) - sy;
        dz = rabs(pbcFold(cz - t.z, 2, box)) - sz;
    }
    else
    {
        dx = rabs(cx - t.x) - sx;
        dy = rabs(cy - t.y) - sy;
        dz = rabs(cz - t.z) - sz;
    }
    dx += rabs(dx);
    dy += rabs(dy);
    dz += rabs(dz);
    dx *= T(0.5);
    dy *= T(0.5);
    dz *= T(0.5);
    T n2 = dx * dx + (dy * dy + dz * dz);
    return n2 < t.radiusSq; // cellRadiusSq == radiusSq for searchExtFactor == 1
}

/*! PBC = false: the box has no periodic dimension, the fold code is not even compiled in.  PBC = true: whether the
 *  fold is needed is decided per warp (any lane whose search sphere leaves the box), so interior warps run the plain
 *  loop; lanes that do not need the fold select the unfolded difference, exactly as the reference picks per particle
 *  (findneighbors.hpp:104-106,150-151). */
template<class T, bool PBC>
__global__ void __launch_bounds__(NB_THREADS) findNeighborsKernel(const T* __restrict__ x,
                                                                  const T* __restrict__ y,
                                                                  const T* __restrict__ z,
                                                                  const T* __restrict__ h,
                                                                  uint32_t first,
                                                                  const uint2* __restrict__ groups,
                                                                  const uint32_t* __restrict__ numGroupsPtr,
                                                                  Box<T> box,
                                                                  const int* __restrict__ childOffsets,
                                                                  const int* __restrict__ parents,
                                                                  const int* __restrict__ internalToLeaf,
                                                                  const uint32_t* __restrict__ layout,
                                                                  const T* __restrict__ centers,
                                                                  const T* __restrict__ sizes,
                                                                  uint32_t ngmax,
                                                                  uint32_t* __restrict__ neighbors,
                                                                  uint32_t* __restrict__ neighborsCount)
{
    __shared__ Staged<T> stageAll[NB_THREADS / 32][32];

    const unsigned lane = threadIdx.x & 31;
    const size_t warpId = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    if (warpId >= size_t(*numGroupsPtr)) { return; }
    Staged<T>* stage = stageAll[threadIdx.x >> 5];

    const uint2 grp  = groups[warpId];
    const bool valid = grp.x + lane < grp.y;
    const uint32_t i = valid ? grp.x + lane : grp.y - 1;

    Target<T> t;
    t.x        = x[i];
    t.y        = y[i];
    t.z        = z[i];
    const T hi = h[i];
    t.radiusSq = T(4.0) * hi * hi;
    {
        bool anyPbc = box.pbc(0) || box.pbc(1) || box.pbc(2);
        T s         = T(2) * hi;
        bool inside = (t.x - s >= box.lim[0]) && (t.y - s >= box.lim[2]) && (t.z - s >= box.lim[4]) &&
                      (t.x + s <= box.lim[1]) && (t.y + s <= box.lim[3]) && (t.z + s <= box.lim[5]);
        t.usePbc    = PBC && anyPbc && !inside;
    }
    const bool warpPbc = PBC && __any_sync(0xffffffffu, t.usePbc);

    uint32_t* row     = neighbors + size_t(i - first) * size_t(ngmax);
    uint32_t numFound = 0;

    auto scanLeaf = [&](int node, bool mine)
    {
        int leafIdx = internalToLeaf[node];
        uint32_t jb = layout[leafIdx];
        uint32_t je = layout[leafIdx + 1];
        // candidates are staged 32 at a time in shared memory with coalesced loads, then broadcast to all lanes
        // (2 LDS per candidate instead of 3 uniform global loads: the loop was LSU-issue bound, profiles/)
        for (uint32_t base = jb; base < je; base += 32)
        {
            const uint32_t cnt = min(32u, je - base);
            __syncwarp();
            if (lane < cnt)
            {
                stage[lane].x = x[base + lane];
                stage[lane].y = y[base + lane];
                stage[lane].z = z[base + lane];
            }
            __syncwarp();
            if (warpPbc)
            {
                for (uint32_t k = 0; k < cnt; ++k)
                {
                    const uint32_t j = base + k;
                    T dx = stage[k].x - t.x;
                    T dy = stage[k].y - t.y;
                    T dz = stage[k].z - t.z;
                    T fx = pbcFold(dx, 0, box);
                    T fy = pbcFold(dy, 1, box);
                    T fz = pbcFold(dz, 2, box);
                    dx   = t.usePbc ? fx : dx;
                    dy   = t.usePbc ? fy : dy;
                    dz   = t.usePbc ? fz : dz;
                    T d2 = dx * dx + dy * dy + dz * dz;
                    if (mine && j != i && d2 < t.radiusSq)
                    {
                        if (numFound < ngmax) { row[numFound] = j; }
                        ++numFound;
                    }
                }
            }
            else
            {
#pragma unroll 4
                for (uint32_t k = 0; k < cnt; ++k)
                {
                    const uint32_t j = base + k;
                    T dx = stage[k].x - t.x;
                    T dy = stage[k].y - t.y;
                    T dz = stage[k].z - t.z;
                    T d2 = dx * dx + dy * dy + dz * dz;
                    if (mine && j != i && d2 < t.radiusSq)
                    {
                        if (numFound < ngmax) { row[numFound] = j; }
                        ++numFound;
                    }
                }
            }
        }
    };

    // bit l of `path` : this lane's own walk reached (passed the test at) the current ancestor of depth l
    uint32_t path = (valid && cellOverlap<PBC>(t, centers, sizes, 0, box)) ? 1u : 0u;
    if (__any_sync(0xffffffffu, path))
    {
        int rootChild = childOffsets[0];
        if (rootChild == 0) { scanLeaf(0, path & 1u); }
        else
        {
            int node       = rootChild;
            int depth      = 1;
            bool backtrack = false;
            while (node != 0)
            {
                int child    = childOffsets[node];
                bool isLeaf  = child == 0;
                bool mine    = false;
                bool descend = false;
                if (!backtrack)
                {
                    mine = ((path >> (depth - 1)) & 1u) && cellOverlap<PBC>(t, centers, sizes, node, box);
                    path = (path & ~(1u << depth)) | (uint32_t(mine) << depth);
                    descend = __any_sync(0xffffffffu, mine);
                }
                if (isLeaf && descend) { scanLeaf(node, mine); }

                if (!isLeaf && descend)
                {
                    node = child;
                    ++depth;
                    backtrack = false;
                }
                else if (((node - 1) & 7) < 7)
                {
                    ++node;
                    backtrack = false;
                }
                else
                {
                    node = parents[(node - 1) >> 3];
                    --depth;
                    backtrack = true;
                }
            }
        }
    }

    if (valid) { neighborsCount[i - first] = numFound; }
}

} // namespace

template<class T>
int findNeighbors(const T* x, const T* y, const T* z, const T* h, uint32_t first, uint32_t last, const double* lim,
                  const int* bnd, int numLeaves, const int* childOffsets, const int* parents, const int* internalToLeaf,
                  const uint32_t* layout, const T* centers, const T* sizes, uint32_t ngmax, uint32_t* neighbors,
                  uint32_t* neighborsCount, cudaStream_t s)
{
    CSB_REQUIRE(last >= first, "invalid particle range");
    CSB_REQUIRE(numLeaves >= 1, "empty tree");
    if (last == first) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);

    // leaf-aligned groups: counts -> exclusive scan -> fill.  Upper bound on the number of groups is known on the host,
    // the exact number stays on the device (no synchronisation)
    size_t maxGroups = size_t(numLeaves) + (size_t(last) - first) / 32 + 1;
    uint32_t* groupOffsets = nullptr;
    uint2* groups          = nullptr;
    void* scanTmp          = nullptr;
    CSB_CHECK(cudaMallocAsync(&groupOffsets, (size_t(numLeaves) + 1) * sizeof(uint32_t), s));
    CSB_CHECK(cudaMallocAsync(&groups, maxGroups * sizeof(uint2), s));
    CSB_CHECK(cudaMallocAsync(&scanTmp, scanTempBytes(size_t(numLeaves) + 1), s));

    groupCountKernel<<<iceil(numLeaves + 1, 256), 256, 0, s>>>(layout, numLeaves, first, last, groupOffsets);
    CSB_LAUNCH_CHECK();
    if (int e = exclusiveScanU32(groupOffsets, groupOffsets, size_t(numLeaves) + 1, scanTmp, s)) { return e; }
    groupFillKernel<<<iceil(numLeaves, 256), 256, 0, s>>>(layout, numLeaves, first, last, groupOffsets, groups);
    CSB_LAUNCH_CHECK();

    unsigned grid = iceil(maxGroups * 32, NB_THREADS);
    if (box.pbc(0) || box.pbc(1) || box.pbc(2))
    {
        findNeighborsKernel<T, true><<<grid, NB_THREADS, 0, s>>>(x, y, z, h, first, groups, groupOffsets + numLeaves,
                                                                 box, childOffsets, parents, internalToLeaf, layout,
                                                                 centers, sizes, ngmax, neighbors, neighborsCount);
    }
    else
    {
        findNeighborsKernel<T, false><<<grid, NB_THREADS, 0, s>>>(x, y, z, h, first, groups, groupOffsets + numLeaves,
                                                                  box, childOffsets, parents, internalToLeaf, layout,
                                                                  centers, sizes, ngmax, neighbors, neighborsCount);
    }
    CSB_LAUNCH_CHECK();
    CSB_CHECK(cudaFreeAsync(groupOffsets, s));
    CSB_CHECK(cudaFreeAsync(groups, s));
    CSB_CHECK(cudaFreeAsync(scanTmp, s));
    return 0;
}

template int findNeighbors<float>(const float*, const float*, const float*, const float*, uint32_t, uint32_t,
                                  const double*, const int*, int, const int*, const int*, const int*, const uint32_t*,
                                  const float*, const float*, uint32_t, uint32_t*, uint32_t*, cudaStream_t);
template int findNeighbors<double>(const double*, const double*, const double*, const double*, uint32_t, uint32_t,
                                   const double*, const int*, int, const int*, const int*, const int*, const uint32_t*,
                                   const double*, const double*, uint32_t, uint32_t*, uint32_t*, cudaStream_t);

} // namespace csb

extern "C"
{

int cs_find_neighbors_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                        const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount, void* stream)
{
    return csb::findNeighbors<float>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                     internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                     cudaStream_t(stream));
}

int cs_find_neighbors_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                        const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                        void* stream)
{
    return csb::findNeighbors<double>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                      internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                      cudaStream_t(stream));
}

} // extern "C"
