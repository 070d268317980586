/* Target particle groups for sm_100a: computeFixedGroups and computeGroupSplits (traversal/groups_gpu.h:33-78,
 * traversal/groups_gpu.cu, traversal/groups_gpu.cuh).
 *
 * computeGroupSplits starts from fixed groups of 32 or 64 consecutive particles (one warp each) and cuts a group
 * wherever two consecutive particles are further apart than min(distCrit, 2 h / minExtent) in unit-box coordinates,
 * distCrit = tolFactor * cbrt(volume of the group's smallest leaf cell in the unit box).
 *
 * The reference stores the split bit masks, turns them into lengths (makeSplits) and scans twice.  Here the first
 * kernel produces the masks and the number of sub-groups per fixed group, and after one scan the second kernel writes
 * the group boundaries directly: a set bit p of fixed group g is the boundary first + g * groupSize + p + 1.  The
 * result array is the same: ascending boundaries from `first` to `last`.
 *
 * Arithmetic follows the reference expression by expression (including its use of the leaf of the FIRST 32-particle
 * segment of a lane for every segment, groups_gpu.cuh:178-186); like the rest of this library it is compiled without
 * FMA contraction.
 */
#include <algorithm>

#include "common.cuh"
#include "cstone_b200.h"
#include "focus.cuh"

namespace csb
{

namespace
{

__global__ void fixedGroupsKernel(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t numGroups,
                                  uint32_t* __restrict__ groups)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < numGroups) { groups[g] = first + g * groupSize; }
    if (g == numGroups) { groups[g] = last; }
}

template<class T>
__device__ inline T warpMinReal(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        T w = __shfl_xor_sync(0xffffffffu, v, o);
        v   = w < v ? w : v;
    }
    return v;
}

/*! split masks of the fixed groups: bit l of masks[g * NWT + k] = the particles k * 32 + l and k * 32 + l + 1 of group g
 *  are cut apart; counts[g] = 1 + number of set bits */
template<class Tc, class Th, int NWT>
__global__ void groupSplitMasksKernel(uint32_t first, uint32_t last, const Tc* __restrict__ x, const Tc* __restrict__ y,
                                      const Tc* __restrict__ z, const Th* __restrict__ h,
                                      const uint64_t* __restrict__ leaves, int numLeaves,
                                      const uint32_t* __restrict__ layout, Box<Tc> box, float tolFactor,
                                      uint32_t numFixedGroups, uint32_t* __restrict__ masks,
                                      uint32_t* __restrict__ counts)
{
    constexpr uint32_t groupSize = NWT * 32;
    const uint32_t g             = uint32_t((size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const unsigned lane          = threadIdx.x & 31;
    if (g >= numFixedGroups) { return; }

    uint32_t body[NWT];
#pragma unroll
    for (int k = 0; k < NWT; ++k)
        body[k] = min(first + g * groupSize + uint32_t(k) * 32u + lane, last - 1);

    // volume (in the unit box) of the leaf cell that holds this lane's particle of segment 0; the smallest one of the
    // warp sets the distance criterion
    const int leaf       = int(upperBound(layout, numLeaves, body[0])) - 1;
    const uint64_t range = leaves[leaf + 1] - leaves[leaf];
    const unsigned level = treeLevel(range);
    constexpr int maxCoord = 1 << KeyTraits<uint64_t>::maxLevel;
    constexpr Th uL        = Th(1.) / maxCoord;
    const int cube         = maxCoord >> level;
    const Th halfUnit      = Th(0.5) * uL * Th(1);
    const Th size          = Th(cube) * halfUnit;
    Th nodeVolume          = Th(1);
    {
        const Th vol = Th(8) * size * size * size;
        nodeVolume   = vol < nodeVolume ? vol : nodeVolume;
    }
    nodeVolume        = warpMinReal(nodeVolume);
    Th root;
    if constexpr (sizeof(Th) == 4) { root = cbrtf(nodeVolume); }
    else { root = cbrt(nodeVolume); }
    const Tc distCrit = root * tolFactor; // std::cbrt(nodeVolume) * tolFactor, groups_gpu.cuh:188
    const Tc critSq   = distCrit * distCrit;

    const Tc minExtent = min(min(box.len[0], box.len[1]), box.len[2]);
    Tc px[NWT], py[NWT], pz[NWT], pr[NWT];
#pragma unroll
    for (int k = 0; k < NWT; ++k)
    {
        px[k] = x[body[k]] * box.ilen[0];
        py[k] = y[body[k]] * box.ilen[1];
        pz[k] = z[body[k]] * box.ilen[2];
        pr[k] = h ? Tc(2) * h[body[k]] / minExtent : Tc(1);
    }

    uint32_t total = 1;
#pragma unroll
    for (int k = 0; k < NWT; ++k)
    {
        // the next particle: the next lane, across the segment boundary the first lane of the next segment; the last
        // particle of the group meets itself (difference 0, never a split)
        Tc nx = __shfl_down_sync(0xffffffffu, px[k], 1);
        Tc ny = __shfl_down_sync(0xffffffffu, py[k], 1);
        Tc nz = __shfl_down_sync(0xffffffffu, pz[k], 1);
        if (k + 1 < NWT)
        {
            const Tc sx = __shfl_sync(0xffffffffu, px[k + 1 < NWT ? k + 1 : k], 0);
            const Tc sy = __shfl_sync(0xffffffffu, py[k + 1 < NWT ? k + 1 : k], 0);
            const Tc sz = __shfl_sync(0xffffffffu, pz[k + 1 < NWT ? k + 1 : k], 0);
            if (lane == 31) { nx = sx, ny = sy, nz = sz; }
        }
        const Tc dx = nx - px[k], dy = ny - py[k], dz = nz - pz[k];
        const Tc distSq = dx * dx + (dy * dy + dz * dz); // norm2: right fold (util/array.hpp)
        const Tc rr     = pr[k] * pr[k];
        const bool split = distSq > (rr < critSq ? rr : critSq); // stl::min(distCritSq, r * r)
        const uint32_t m = __ballot_sync(0xffffffffu, split);
        if (lane == 0) { masks[size_t(g) * NWT + k] = m; }
        total += __popc(m);
    }
    if (lane == 0) { counts[g] = total; }
}

//! groups[offsets[g] + j] = j-th boundary of fixed group g (its start, then one per set mask bit); the final entry = last
template<int NWT>
__global__ void groupSplitFillKernel(uint32_t first, uint32_t last, uint32_t numFixedGroups,
                                     const uint32_t* __restrict__ masks, const uint32_t* __restrict__ offsets,
                                     uint32_t* __restrict__ groups)
{
    constexpr uint32_t groupSize = NWT * 32;
    const uint32_t g             = uint32_t((size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const unsigned lane          = threadIdx.x & 31;
    if (g >= numFixedGroups) { return; }
    uint32_t pos = offsets[g];
    if (lane == 0) { groups[pos] = first + g * groupSize; }
    ++pos;
#pragma unroll
    for (int k = 0; k < NWT; ++k)
    {
        const uint32_t m = masks[size_t(g) * NWT + k];
        if ((m >> lane) & 1u)
        {
            groups[pos + __popc(m & ((1u << lane) - 1u))] = first + g * groupSize + uint32_t(k) * 32u + lane + 1u;
        }
        pos += __popc(m);
    }
    if (g + 1 == numFixedGroups && lane == 0) { groups[offsets[numFixedGroups]] = last; }
}

} // namespace

int computeFixedGroups(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, cudaStream_t s)
{
    CSB_REQUIRE(last >= first && groupSize > 0, "computeFixedGroups: invalid range or group size");
    const uint32_t numGroups = uint32_t(iceil(size_t(last - first), groupSize));
    fixedGroupsKernel<<<iceil(size_t(numGroups) + 1, 256), 256, 0, s>>>(first, last, groupSize, numGroups, groups);
    CSB_LAUNCH_CHECK();
    return 0;
}

/*! first half of computeGroupSplits: masks and sub-group counts of the fixed groups, scanned; the total number of groups
 *  comes back on the host (the reference reads it back at the same point, groups_gpu.cu:86-88).  The masks and offsets
 *  stay in the library's scratch memory of this (thread, stream) for groupSplitsFinish. */
template<class Tc, class Th>
int groupSplitsBegin(uint32_t first, uint32_t last, const Tc* x, const Tc* y, const Tc* z, const Th* h,
                     const uint64_t* leaves, int numLeaves, const uint32_t* layout, const double* lim, const int* bnd,
                     uint32_t groupSize, float tolFactor, uint32_t* numGroupsOut, cudaStream_t s)
{
    CSB_REQUIRE(groupSize == 32 || groupSize == 64, "Unsupported spatial group size");
    CSB_REQUIRE(last >= first && numLeaves >= 1, "computeGroupSplits: invalid particle range or empty tree");
    *numGroupsOut = 0;
    const uint32_t numFixed = uint32_t(iceil(size_t(last - first), groupSize));
    if (numFixed == 0) { return 0; }
    const int nwt = int(groupSize / 32);
    CSB_SCRATCH(masks, uint32_t*, s, SCRATCH_A, size_t(numFixed) * nwt * sizeof(uint32_t));
    CSB_SCRATCH(counts, uint32_t*, s, SCRATCH_B, (size_t(numFixed) + 1) * sizeof(uint32_t));
    CSB_SCRATCH(scanTmp, void*, s, SCRATCH_C, scanTempBytes(size_t(numFixed) + 1));
    CSB_CHECK(cudaMemsetAsync(counts + numFixed, 0, sizeof(uint32_t), s));
    Box<Tc> box = makeBox<Tc>(lim, bnd);
    const unsigned grid = iceil(size_t(numFixed) * 32, 256);
    if (nwt == 1)
    {
        groupSplitMasksKernel<Tc, Th, 1><<<grid, 256, 0, s>>>(first, last, x, y, z, h, leaves, numLeaves, layout, box,
                                                               tolFactor, numFixed, masks, counts);
    }
    else
    {
        groupSplitMasksKernel<Tc, Th, 2><<<grid, 256, 0, s>>>(first, last, x, y, z, h, leaves, numLeaves, layout, box,
                                                               tolFactor, numFixed, masks, counts);
    }
    CSB_LAUNCH_CHECK();
    if (int e = exclusiveScanU32(counts, counts, size_t(numFixed) + 1, scanTmp, s)) { return e; }
    CSB_CHECK(cudaMemcpyAsync(numGroupsOut, counts + numFixed, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CSB_CHECK(cudaStreamSynchronize(s));
    return 0;
}

//! second half: groups[0 .. numGroups] = the ascending group boundaries, groups[numGroups] = last
int groupSplitsFinish(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, cudaStream_t s)
{
    CSB_REQUIRE(groupSize == 32 || groupSize == 64, "Unsupported spatial group size");
    const uint32_t numFixed = uint32_t(iceil(size_t(last - first), groupSize));
    if (numFixed == 0) { return 0; }
    const int nwt = int(groupSize / 32);
    CSB_SCRATCH(masks, uint32_t*, s, SCRATCH_A, size_t(numFixed) * nwt * sizeof(uint32_t));
    CSB_SCRATCH(offsets, uint32_t*, s, SCRATCH_B, (size_t(numFixed) + 1) * sizeof(uint32_t));
    const unsigned grid = iceil(size_t(numFixed) * 32, 256);
    if (nwt == 1) { groupSplitFillKernel<1><<<grid, 256, 0, s>>>(first, last, numFixed, masks, offsets, groups); }
    else { groupSplitFillKernel<2><<<grid, 256, 0, s>>>(first, last, numFixed, masks, offsets, groups); }
    CSB_LAUNCH_CHECK();
    return 0;
}

} // namespace csb

extern "C"
{

int cs_compute_fixed_groups(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, void* stream)
{
    return csb::computeFixedGroups(first, last, groupSize, groups, cudaStream_t(stream));
}

int cs_group_splits_begin_dd(uint32_t first, uint32_t last, const double* x, const double* y, const double* z,
                             const double* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream)
{
    return csb::groupSplitsBegin<double, double>(first, last, x, y, z, h, leaves, numLeaves, layout, lim, bnd, groupSize,
                                                 tolFactor, numGroupsOut, cudaStream_t(stream));
}

int cs_group_splits_begin_df(uint32_t first, uint32_t last, const double* x, const double* y, const double* z,
                             const float* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream)
{
    return csb::groupSplitsBegin<double, float>(first, last, x, y, z, h, leaves, numLeaves, layout, lim, bnd, groupSize,
                                                tolFactor, numGroupsOut, cudaStream_t(stream));
}

int cs_group_splits_begin_ff(uint32_t first, uint32_t last, const float* x, const float* y, const float* z,
                             const float* h, const uint64_t* leaves, int numLeaves, const uint32_t* layout,
                             const double* lim, const int* bnd, uint32_t groupSize, float tolFactor,
                             uint32_t* numGroupsOut, void* stream)
{
    return csb::groupSplitsBegin<float, float>(first, last, x, y, z, h, leaves, numLeaves, layout, lim, bnd, groupSize,
                                               tolFactor, numGroupsOut, cudaStream_t(stream));
}

int cs_group_splits_finish(uint32_t first, uint32_t last, uint32_t groupSize, uint32_t* groups, void* stream)
{
    return csb::groupSplitsFinish(first, last, groupSize, groups, cudaStream_t(stream));
}

} // extern "C"
