/* Radius neighbour search for sm_100a producing the reference's CPU list layout:
 * cstone::findNeighbors (findneighbors.hpp:77-177): neighbors[(i-first)*ngmax + k], neighborsCount[i-first].
 *
 * Design: one warp owns 32 consecutive (SFC-adjacent) target particles and walks the octree ONCE for all of them with a
 * warp-uniform, stackless depth-first traversal (child / next sibling / parent links as in traversal/traversal.hpp:26-69).
 * Each lane keeps the exact per-particle pruning state of the reference's per-particle walk: a bit per tree depth says
 * whether this lane's own continuation test (point-to-cell min distance < (2h)^2, boxoverlap.hpp:229-250) passed on the
 * current root path.  The warp descends while any lane passes; at a leaf only lanes whose own path passed scan the leaf's
 * particles.  Leaves are reached in SFC order, so every lane appends neighbours in ascending particle index exactly like
 * the CPU walk — truncation at ngmax keeps the same entries — and the distance arithmetic is the reference's,
 * operation by operation (no FMA contraction; norm2 is the right fold x*x + (y*y + z*z), util/array.hpp:236-240).
 * Self exclusion is by index (j != i) as on the CPU (SURVEY.md hazard H2).
 */
#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

constexpr int NB_THREADS = 128;

template<class T>
struct Target
{
    T x, y, z;
    T radiusSq;
    bool usePbc;
};

//! continuation test of findneighbors.hpp:108-112
template<class T>
__device__ inline bool cellOverlap(const Target<T>& t, const T* __restrict__ centers, const T* __restrict__ sizes,
                                   int node, const Box<T>& box)
{
    T cx = centers[3 * node], cy = centers[3 * node + 1], cz = centers[3 * node + 2];
    T sx = sizes[3 * node], sy = sizes[3 * node + 1], sz = sizes[3 * node + 2];
    T dx, dy, dz;
    if (t.usePbc)
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

template<class T>
__global__ void __launch_bounds__(NB_THREADS) findNeighborsKernel(const T* __restrict__ x,
                                                                  const T* __restrict__ y,
                                                                  const T* __restrict__ z,
                                                                  const T* __restrict__ h,
                                                                  uint32_t first,
                                                                  uint32_t last,
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
    const unsigned lane   = threadIdx.x & 31;
    const size_t warpId   = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    const size_t firstTgt = size_t(first) + warpId * 32;
    if (firstTgt >= last) { return; }

    const size_t iLong = firstTgt + lane;
    const bool valid   = iLong < last;
    const uint32_t i   = valid ? uint32_t(iLong) : uint32_t(last - 1);

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
        t.usePbc    = anyPbc && !inside;
    }

    uint32_t* row     = neighbors + (iLong - first) * size_t(ngmax);
    uint32_t numFound = 0;

    auto scanLeaf = [&](int node, bool mine)
    {
        int leafIdx  = internalToLeaf[node];
        uint32_t jb  = layout[leafIdx];
        uint32_t je  = layout[leafIdx + 1];
        for (uint32_t j = jb; j < je; ++j)
        {
            T dx = x[j] - t.x;
            T dy = y[j] - t.y;
            T dz = z[j] - t.z;
            if (t.usePbc)
            {
                dx = pbcFold(dx, 0, box);
                dy = pbcFold(dy, 1, box);
                dz = pbcFold(dz, 2, box);
            }
            T d2 = dx * dx + dy * dy + dz * dz;
            if (mine && j != i && d2 < t.radiusSq)
            {
                if (numFound < ngmax) { row[numFound] = j; }
                ++numFound;
            }
        }
    };

    // bit l of `path` : this lane's own walk reached (passed the test at) the current ancestor of depth l
    uint32_t path = (valid && cellOverlap(t, centers, sizes, 0, box)) ? 1u : 0u;
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
                    mine = ((path >> (depth - 1)) & 1u) && cellOverlap(t, centers, sizes, node, box);
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

    if (valid) { neighborsCount[iLong - first] = numFound; }
}

} // namespace

template<class T>
int findNeighbors(const T* x, const T* y, const T* z, const T* h, uint32_t first, uint32_t last, const double* lim,
                  const int* bnd, const int* childOffsets, const int* parents, const int* internalToLeaf,
                  const uint32_t* layout, const T* centers, const T* sizes, uint32_t ngmax, uint32_t* neighbors,
                  uint32_t* neighborsCount, cudaStream_t s)
{
    CSB_REQUIRE(last >= first, "invalid particle range");
    if (last == first) { return 0; }
    Box<T> box        = makeBox<T>(lim, bnd);
    size_t numTargets = size_t(last) - first;
    size_t numWarps   = (numTargets + 31) / 32;
    unsigned grid     = iceil(numWarps * 32, NB_THREADS);
    findNeighborsKernel<T><<<grid, NB_THREADS, 0, s>>>(x, y, z, h, first, last, box, childOffsets, parents,
                                                       internalToLeaf, layout, centers, sizes, ngmax, neighbors,
                                                       neighborsCount);
    CSB_LAUNCH_CHECK();
    return 0;
}

template int findNeighbors<float>(const float*, const float*, const float*, const float*, uint32_t, uint32_t,
                                  const double*, const int*, const int*, const int*, const int*, const uint32_t*,
                                  const float*, const float*, uint32_t, uint32_t*, uint32_t*, cudaStream_t);
template int findNeighbors<double>(const double*, const double*, const double*, const double*, uint32_t, uint32_t,
                                   const double*, const int*, const int*, const int*, const int*, const uint32_t*,
                                   const double*, const double*, uint32_t, uint32_t*, uint32_t*, cudaStream_t);

} // namespace csb

extern "C"
{

int cs_find_neighbors_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                        const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount, void* stream)
{
    return csb::findNeighbors<float>(x, y, z, h, firstId, lastId, lim, bnd, childOffsets, parents, internalToLeaf,
                                     layout, centers, sizes, ngmax, neighbors, neighborsCount, cudaStream_t(stream));
}

int cs_find_neighbors_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                        const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                        void* stream)
{
    return csb::findNeighbors<double>(x, y, z, h, firstId, lastId, lim, bnd, childOffsets, parents, internalToLeaf,
                                      layout, centers, sizes, ngmax, neighbors, neighborsCount, cudaStream_t(stream));
}

} // extern "C"
