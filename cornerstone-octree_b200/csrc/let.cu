/* Kernels of the locally essential tree (LET) of a multi-rank domain:
 *   - min-distance MAC spheres and the MAC marking traversal (traversal/macs.hpp:40-57,118-260,
 *     traversal/collisions_gpu.cu:91-140, focus/octree_focus_mpi.hpp:422-499)
 *   - node counts of LET leaves taken from the replicated global tree (focus/rebalance.hpp:263-284)
 *   - index gathers/scatters of the peer count exchange (focus/exchange_focus.hpp:310-366)
 *   - range packing of the halo exchange (halos/gather_halos_gpu.cu:27-58)
 */
#include "common.cuh"
#include "cstone_b200.h"
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

/*! markMacs (traversal/macs.hpp:149-260): marks every LET node outside the focus that fails the MAC against a focus
 *  leaf whose extended box is not interior to the focus.  The reference walks the tree once per leaf; here a warp
 *  owns 32 SFC-consecutive focus leaves and walks the UNION of their traversals once, warp-uniformly: bit d of a
 *  lane's `path` says that this lane's own walk descended at depth d on the current root path, so every lane
 *  evaluates the MAC for exactly the nodes its own traversal (traversal/traversal.hpp:26-69) visits and the marks
 *  are the same set.  Stores race benignly (all write 1). */
template<class K, class T>
__global__ void __launch_bounds__(128) markMacsKernel(const K* __restrict__ prefixes,
                                                      const int* __restrict__ childOffsets,
                                                      const int* __restrict__ parents,
                                                      const T* __restrict__ centers4,
                                                      Box<T> box,
                                                      const K* __restrict__ focusNodes,
                                                      int numFocusNodes,
                                                      bool limitSource,
                                                      uint8_t* markings)
{
    __shared__ unsigned char hilbertTables[hilbertTableBytes];
    stageHilbertTables(hilbertTables);
    __syncthreads();
    const int tid          = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int maxCoord = 1 << KeyTraits<K>::maxLevel;
    constexpr T uL         = T(1) / maxCoord;
    constexpr unsigned FULL = 0xffffffffu;

    const K focusStart = focusNodes[0];
    const K focusEnd   = focusNodes[numFocusNodes];
    bool active        = tid < numFocusNodes;
    const int leaf     = active ? tid : numFocusNodes - 1;
    const K a = focusNodes[leaf], b = focusNodes[leaf + 1];

    unsigned level      = treeLevel<K>(b - a);
    // limitSource: sources deeper than one level above the target are neither marked nor entered (macs.hpp:221-222)
    const int maxSourceLevel = limitSource ? max(int(level) - 1, 0) : int(KeyTraits<K>::maxLevel);
    unsigned cubeLength = unsigned(maxCoord) >> level;
    unsigned mask       = ~(cubeLength - 1);
    unsigned ix, iy, iz;
    hilbertDecode(a, ix, iy, iz, hilbertTables);
    int lo[3] = {int(ix & mask), int(iy & mask), int(iz & mask)};
    int hi[3] = {lo[0] + int(cubeLength), lo[1] + int(cubeLength), lo[2] + int(cubeLength)};
    /* containedIn(focusStart, focusEnd, leaf box extended by one unit) (traversal/boxoverlap.hpp:96-117) without
     * encoding the two corner keys: the corners (lo - 1) and (hi) lie in the same cell of side 2^m exactly when no
     * face of the leaf box is a multiple of 2^m, so the smallest common cell is the leaf's ancestor with
     * m = 1 + max(ctz(lo), ctz(hi)) over the three dimensions, whose key range is a prefix of the leaf key */
    {
        bool contained;
        int mn = min(min(lo[0], lo[1]), lo[2]);
        int mx = max(max(hi[0], hi[1]), hi[2]);
        if (mn == 0 || mx == maxCoord) { contained = focusStart == 0 && focusEnd == nodeRange<K>(0); }
        else
        {
            int m = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
                m = max(m, max(__ffs(lo[d]), __ffs(hi[d]))); // __ffs = ctz + 1
            unsigned ancLevel = unsigned(KeyTraits<K>::maxLevel - m);
            K nodeStart       = a & ~(nodeRange<K>(ancLevel) - 1);
            K nodeEnd         = nodeStart + nodeRange<K>(ancLevel);
            contained         = nodeStart >= focusStart && nodeEnd <= focusEnd;
        }
        active = active && !contained;
    }
    if (!__any_sync(FULL, active)) { return; }

    T tc[3], ts[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
        T halfUnit = T(0.5) * uL * box.len[d];
        tc[d]      = box.lim[2 * d] + T(hi[d] + lo[d]) * halfUnit;
        ts[d]      = T(hi[d] - lo[d]) * halfUnit;
    }

    /* singleTraversal (traversal/traversal.hpp:26-69) for 32 leaves at once, one SIBLING GROUP per step: lanes 0-7 fetch
     * the 8 children of the node being expanded (one round trip to L2 instead of one per node - the walk of a leaf at
     * the focus boundary visits thousands of nodes and its latency chain is what this kernel's run time consists of),
     * then every lane evaluates the MAC of the children its own walk reaches from shared memory. */
    struct Group
    {
        T cx[8], cy[8], cz[8], mac2[8];
        int child[8];
    };
    __shared__ Group groups[128 / 32];
    __shared__ uint8_t laneMask[128 / 32][KeyTraits<K>::maxLevel + 2][32];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Group& g            = groups[warp];

    //! evaluateMacPbc (macs.hpp:118-130) of this lane's leaf box against a source centre
    auto violates = [&](T cx, T cy, T cz, T mac2)
    {
        T c[3] = {cx, cy, cz};
        T dx[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            T v = tc[d] - c[d];
            v   = rabs(pbcFold(v, d, box));
            v -= ts[d];
            v += rabs(v);
            v *= T(0.5);
            dx[d] = v;
        }
        T R2 = dx[0] * dx[0] + (dx[1] * dx[1] + dx[2] * dx[2]);
        return R2 < rabs(mac2);
    };
    auto outsideFocus = [&](int idx)
    {
        K nodePrefix         = prefixes[idx];
        unsigned sourceLevel = decodePrefixLength(nodePrefix) / 3;
        K nodeStart          = decodePlaceholderBit(nodePrefix);
        K nodeEnd            = nodeStart + nodeRange<K>(sourceLevel);
        return nodeStart < focusStart || nodeEnd > focusEnd;
    };

    /* Bounding box of the active leaves of the warp.  The expression the MAC compares, sum over d of
     * max(p(tc_d - c_d) - ts_d, 0)^2 with p = distance to the nearest multiple of the period (or |.| for open
     * dimensions), can only grow when the target box shrinks inside its bounding box: p(a + b) <= p(a) + |b| and
     * |bc_d - tc_d| <= bs_d - ts_d.  A source whose expression against the bounding box already exceeds mac^2 (with a
     * margin far above the rounding errors of either evaluation) therefore passes the MAC of every lane, and the
     * per-lane tests of that child are skipped - the marks stay exactly those of the per-leaf walks. */
    T bc[3], bs[3];
    {
        constexpr T big = sizeof(T) == 8 ? T(1e300) : T(1e30);
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
            T mn = active ? tc[d] - ts[d] : big;
            T mx = active ? tc[d] + ts[d] : -big;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                T a = __shfl_xor_sync(FULL, mn, o), b = __shfl_xor_sync(FULL, mx, o);
                mn  = a < mn ? a : mn;
                mx  = b > mx ? b : mx;
            }
            bc[d] = T(0.5) * (mx + mn);
            // inflated beyond the rounding of bc and bs themselves, so that the bounding-box expression is a lower
            // bound of every lane's also in floating point
            bs[d] = T(0.5) * (mx - mn) * (T(1) + (sizeof(T) == 8 ? T(1e-12) : T(1e-5))) +
                    (sizeof(T) == 8 ? T(1e-13) : T(1e-6)) * (rabs(mx) + rabs(mn));
        }
    }
    constexpr T cullMargin = sizeof(T) == 8 ? T(1e-9) : T(1e-3);

    //! this lane's decisions for the 8 children starting at child0 (bit c: the lane's walk marks and enters child c)
    auto testChildren = [&](int child0, bool mine, int sourceLevel) -> unsigned
    {
        __syncwarp();
        bool test = false;
        if (lane < 8)
        {
            int idx         = child0 + int(lane);
            const T cx = centers4[4 * idx], cy = centers4[4 * idx + 1], cz = centers4[4 * idx + 2];
            const T mac2    = centers4[4 * idx + 3];
            g.cx[lane]      = cx;
            g.cy[lane]      = cy;
            g.cz[lane]      = cz;
            g.mac2[lane]    = mac2;
            g.child[lane]   = childOffsets[idx];
            if (outsideFocus(idx))
            {
                const T c[3] = {cx, cy, cz};
                T r2         = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                {
                    T v = rabs(pbcFold(bc[d] - c[d], d, box)) - bs[d];
                    v   = v > T(0) ? v : T(0);
                    r2 += v * v;
                }
                test = !(r2 * (T(1) - cullMargin) > rabs(mac2)); // not certainly far for every lane
            }
        }
        const unsigned testMask = __ballot_sync(FULL, test);
        unsigned bits           = 0;
        if (mine && sourceLevel <= maxSourceLevel)
        {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (((testMask >> c) & 1u) && violates(g.cx[c], g.cy[c], g.cz[c], g.mac2[c])) { bits |= 1u << c; }
        }
        return bits;
    };

    bool viol = active && outsideFocus(0) &&
                violates(centers4[0], centers4[1], centers4[2], centers4[3]);
    if (!__any_sync(FULL, viol)) { return; }
    if (lane == 0 && !markings[0]) { markings[0] = 1; }
    int base = childOffsets[0];
    if (base == 0) { return; }

    int depth     = 1;
    unsigned lm   = testChildren(base, viol, 1);
    unsigned wm   = __reduce_or_sync(FULL, lm);
    laneMask[warp][1][lane] = uint8_t(lm);
    if (lane < 8 && ((wm >> lane) & 1u) && !markings[base + lane]) { markings[base + lane] = 1; }
    unsigned enter = 0; // children of the current group that are internal nodes and entered by some lane
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (((wm >> c) & 1u) && g.child[c] != 0) { enter |= 1u << c; }
    while (true)
    {
        if (enter == 0)
        {
            if (depth == 1) { return; }
            // back to the parent's sibling group: its per-lane decisions are in shared memory, the children still
            // to expand are recomputed from the group data (re-fetched: the walk is depth-first)
            const int up = parents[(base - 1) >> 3];
            --depth;
            base = ((up - 1) & ~7) + 1;
            lm   = laneMask[warp][depth][lane];
            wm   = __reduce_or_sync(FULL, lm);
            __syncwarp();
            if (lane < 8) { g.child[lane] = childOffsets[base + int(lane)]; }
            __syncwarp();
            enter = 0;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (((wm >> c) & 1u) && g.child[c] != 0 && c > ((up - 1) & 7)) { enter |= 1u << c; }
            continue;
        }
        const int c = __ffs(int(enter)) - 1;
        enter &= enter - 1;
        const bool mine = (lm >> c) & 1u;
        const int child = g.child[c];
        ++depth;
        base = child;
        lm   = testChildren(child, mine, depth); // the children of a node at tree level depth - 1
        wm   = __reduce_or_sync(FULL, lm);
        laneMask[warp][depth][lane] = uint8_t(lm);
        if (lane < 8 && ((wm >> lane) & 1u) && !markings[base + lane]) { markings[base + lane] = 1; }
        enter = 0;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
            if (((wm >> cc) & 1u) && g.child[cc] != 0) { enter |= 1u << cc; }
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


/* ---- device-side pieces of FocusedOctree::updateTree's rank-to-rank part (octree_focus_mpi.hpp:137-165); the leaf and
 *      prefix arrays never leave HBM, only O(numRanks) scalars are read back by the host ---- */

//! findNodeAbove / findNodeBelow (tree/csarray.hpp:73-95) of a handful of keys: out[t] = lower_bound,
//! out[numBounds + t] = upper_bound - 1 over the numLeaves + 1 leaf keys
template<class K>
__global__ void focusBoundsKernel(const K* __restrict__ leaves, int numKeys, const K* __restrict__ bounds,
                                  int numBounds, int* __restrict__ out)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= numBounds) { return; }
    out[t]             = lowerBound(leaves, numKeys, bounds[t]);
    out[numBounds + t] = upperBound(leaves, numKeys, bounds[t]) - 1;
}

//! !std::includes(global leaves of a rank, my focus leaves in that rank's range) (focus/peer_flags.hpp:33-57)
template<class K>
__global__ void notIncludedKernel(const K* __restrict__ fLeaves, int count, const K* __restrict__ gLeaves, int gCount,
                                  int* __restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    K k   = fLeaves[i];
    int j = lowerBound(gLeaves, gCount, k);
    if (j == gCount || gLeaves[j] != k) { *flag = 1; }
}

//! checkTreelets (focus/exchange_focus.hpp:60-115): a key of a peer's treelet is valid if it is one of my leaf keys;
//! the last key of a treelet and the curve ends are always valid
template<class K>
__global__ void checkTreeletKernel(const K* __restrict__ treelet, int count, const K* __restrict__ leaves,
                                   int numLeaves, uint32_t* __restrict__ valid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    K k    = treelet[i];
    bool v = true;
    if (i + 1 < count && k != 0 && k != nodeRange<K>(0))
    {
        int j = lowerBound(leaves, numLeaves, k);
        v     = (leaves[j] == k); // leaves has numLeaves + 1 entries
    }
    valid[i] = v ? 1u : 0u;
}

//! pruneTreelets: stable split of the received keys into accepted (treelet) and rejected keys
template<class K>
__global__ void splitTreeletKernel(const K* __restrict__ keys, int count, const uint32_t* __restrict__ validScan,
                                   K* __restrict__ accepted, K* __restrict__ rejected)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    uint32_t pos = validScan[i] - validScan[0]; // accepted keys of this treelet before key i
    bool v       = validScan[i + 1] != validScan[i];
    if (v) { accepted[pos] = keys[i]; }
    else { rejected[uint32_t(i) - pos] = keys[i]; }
}

//! leaves named by rejected keys are removed: nodeOps[findNodeAbove(key)] = 0
template<class K>
__global__ void rejectLeavesKernel(const K* __restrict__ rejected, int count, const K* __restrict__ leaves, int numKeys,
                                   int* __restrict__ nodeOps)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) { return; }
    nodeOps[lowerBound(leaves, numKeys, rejected[i])] = 0;
}

__global__ void fillIntKernel(int* a, int n, int v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = v; }
}

//! indexTreelets (focus/exchange_focus.hpp:286-308): node index of every treelet leaf in the level-sorted prefixes
template<class K>
__global__ void indexTreeletKernel(const K* __restrict__ treelet, int numNodes, const K* __restrict__ prefixes,
                                   const int* __restrict__ levelRange, int* __restrict__ out, int* __restrict__ error)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) { return; }
    K a = treelet[i], b = treelet[i + 1];
    unsigned level = treeLevel<K>(b - a);
    K prefix       = encodePlaceholderBit(a, int(3 * level));
    int first = levelRange[level], last = levelRange[level + 1];
    int j     = first + lowerBound(prefixes + first, last - first, prefix);
    if (j == last || prefixes[j] != prefix) { *error = 1; }
    out[i] = j;
}

//! gatherRanges for one array of elements made of `words` 32-bit words (any field type of the client, as the
//! reference reinterprets its payloads to int / util::array<float, 1..4>, halos/pack_buffers.hpp:50-54)
__global__ void gatherRangesWordsKernel(const uint32_t* __restrict__ rangeScan, const uint32_t* __restrict__ rangeStart,
                                        int numRanges, uint32_t total, int words, const uint32_t* __restrict__ src,
                                        uint32_t* __restrict__ out)
{
    size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(total) * words) { return; }
    uint32_t k = uint32_t(t / words);
    int w      = int(t - size_t(k) * words);
    int r      = int(upperBound(rangeScan, numRanges, k)) - 1;
    size_t e   = size_t(rangeStart[r]) + (k - rangeScan[r]);
    out[t]     = src[e * words + w];
}

/*! out[k] = value of the particle at exchange-buffer position order[k]: positions inside [recvStart, recvStart+numRecv)
 *  come from the receive staging buffer, all others from the array as it was before the sync (the replay of
 *  Domain::reapplySync, domain/domain.hpp:297-329, in one pass instead of redoExchange + gatherArrays) */
__global__ void replayGatherWordsKernel(const uint32_t* __restrict__ order, uint32_t n, int words,
                                        const uint32_t* __restrict__ before, const uint32_t* __restrict__ received,
                                        uint32_t recvStart, uint32_t numRecv, uint32_t* __restrict__ out)
{
    size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(n) * words) { return; }
    uint32_t k = uint32_t(t / words);
    int w      = int(t - size_t(k) * words);
    uint32_t p = order[k];
    uint32_t q = p - recvStart;
    out[t]     = q < numRecv ? received[size_t(q) * words + w] : before[size_t(p) * words + w];
}

//! out[k] = src[order[k]] for elements of `words` 32-bit words
__global__ void gatherWordsKernel(const uint32_t* __restrict__ order, uint32_t n, int words,
                                  const uint32_t* __restrict__ src, uint32_t* __restrict__ out)
{
    size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(n) * words) { return; }
    uint32_t k = uint32_t(t / words);
    int w      = int(t - size_t(k) * words);
    out[t]     = src[size_t(order[k]) * words + w];
}

/* ---- device-side checkLayout (domain/layout.hpp:187-219), halo request keys (extractMarkedElements,
 *      domain/layout.hpp:110-141, per peer range) and their translation into outgoing index ranges
 *      (halos/halos.hpp:64-80, domain/exchange_keys.hpp:45-99) ---- */

//! rank r != me whose focus range [fa[2r], fa[2r+1]) holds leaf i, or -1
__device__ inline int peerOfLeaf(const int* __restrict__ fa, int numRanks, int me, int i)
{
    for (int r = 0; r < numRanks; ++r)
        if (r != me && fa[2 * r] <= i && i < fa[2 * r + 1]) { return r; }
    return -1;
}

/*! flags[i] = 1 if leaf i starts a run of consecutive halo leaves (layout count > 0) inside a peer's range;
 *  status[0] |= a halo leaf lies in no rank's range, status[1] |= a foreign leaf holds more than maxParticles */
__global__ void haloRunStartKernel(const uint32_t* __restrict__ layout, int numLeaves, const int* __restrict__ fa,
                                   int numRanks, int me, uint32_t maxParticles, uint32_t* __restrict__ flags,
                                   int* __restrict__ status)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numLeaves) { return; }
    uint32_t f = 0;
    if (i < numLeaves && !(fa[2 * me] <= i && i < fa[2 * me + 1]))
    {
        uint32_t cnt = layout[i + 1] - layout[i];
        if (cnt > 0)
        {
            int r = peerOfLeaf(fa, numRanks, me, i);
            if (r < 0) { status[0] = 1; }
            else
            {
                bool prevMarked = i > fa[2 * r] && layout[i] > layout[i - 1];
                f               = prevMarked ? 0u : 1u;
            }
        }
        if (cnt > maxParticles) { status[1] = 1; }
    }
    flags[i] = f;
}

//! request key pairs: run k (global numbering from the scan of the run starts) -> req[2k] = first key, req[2k+1] = end key
template<class K>
__global__ void haloRequestKeysKernel(const uint32_t* __restrict__ layout, int numLeaves, const int* __restrict__ fa,
                                      int numRanks, int me, const uint32_t* __restrict__ startScan,
                                      const K* __restrict__ leaves, K* __restrict__ req)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numLeaves) { return; }
    if (fa[2 * me] <= i && i < fa[2 * me + 1]) { return; }
    if (layout[i + 1] == layout[i]) { return; }
    int r = peerOfLeaf(fa, numRanks, me, i);
    if (r < 0) { return; }
    bool isStart = startScan[i + 1] != startScan[i];
    if (isStart) { req[2 * size_t(startScan[i])] = leaves[i]; }
    bool isEnd = (i + 1 == fa[2 * r + 1]) || layout[i + 2] == layout[i + 1];
    if (isEnd) { req[2 * size_t(startScan[i + 1] - 1) + 1] = leaves[i + 1]; }
}

//! requested key pairs -> particle index ranges of my layout: start[q] = layout[nodeAbove(k0)], len[q] = extent
template<class K>
__global__ void haloRangesKernel(const K* __restrict__ pairs, int numPairs, const K* __restrict__ leaves, int numKeys,
                                 const uint32_t* __restrict__ layout, uint32_t* __restrict__ len,
                                 uint32_t* __restrict__ start)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > numPairs) { return; }
    if (q == numPairs)
    {
        len[q] = 0;
        return;
    }
    uint32_t lo = layout[lowerBound(leaves, numKeys, pairs[2 * q])];
    uint32_t hi = layout[lowerBound(leaves, numKeys, pairs[2 * q + 1])];
    start[q]    = lo;
    len[q]      = hi - lo;
}

__global__ void gatherU32PlainKernel(const uint32_t* __restrict__ src, const int* __restrict__ idx, int n,
                                     uint32_t* __restrict__ dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { dst[i] = src[idx[i]]; }
}

} // namespace

int haloRunStarts(const uint32_t* layout, int numLeaves, const int* fa, int numRanks, int me, uint32_t maxParticles,
                  uint32_t* flags, int* status, cudaStream_t s)
{
    haloRunStartKernel<<<iceil(numLeaves + 1, 256), 256, 0, s>>>(layout, numLeaves, fa, numRanks, me, maxParticles,
                                                                 flags, status);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int haloRequestKeys(const uint32_t* layout, int numLeaves, const int* fa, int numRanks, int me,
                    const uint32_t* startScan, const K* leaves, K* req, cudaStream_t s)
{
    haloRequestKeysKernel<K><<<iceil(numLeaves, 256), 256, 0, s>>>(layout, numLeaves, fa, numRanks, me, startScan,
                                                                   leaves, req);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int haloRanges(const K* pairs, int numPairs, const K* leaves, int numKeys, const uint32_t* layout, uint32_t* len,
               uint32_t* start, cudaStream_t s)
{
    haloRangesKernel<K><<<iceil(numPairs + 1, 256), 256, 0, s>>>(pairs, numPairs, leaves, numKeys, layout, len, start);
    CSB_LAUNCH_CHECK();
    return 0;
}
template int haloRequestKeys<uint32_t>(const uint32_t*, int, const int*, int, int, const uint32_t*, const uint32_t*,
                                       uint32_t*, cudaStream_t);
template int haloRequestKeys<uint64_t>(const uint32_t*, int, const int*, int, int, const uint32_t*, const uint64_t*,
                                       uint64_t*, cudaStream_t);
template int haloRanges<uint32_t>(const uint32_t*, int, const uint32_t*, int, const uint32_t*, uint32_t*, uint32_t*,
                                  cudaStream_t);
template int haloRanges<uint64_t>(const uint64_t*, int, const uint64_t*, int, const uint32_t*, uint32_t*, uint32_t*,
                                  cudaStream_t);

//! dst[i] = src[idx[i]] for a handful of indices (scalars the host needs from a device array)
int pickU32(const uint32_t* src, const int* idx, int n, uint32_t* dst, cudaStream_t s)
{
    if (n <= 0) { return 0; }
    gatherU32PlainKernel<<<iceil(n, 64), 64, 0, s>>>(src, idx, n, dst);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int focusBounds(const K* leaves, int numKeys, const K* bounds, int numBounds, int* out, cudaStream_t s)
{
    focusBoundsKernel<K><<<iceil(numBounds, 64), 64, 0, s>>>(leaves, numKeys, bounds, numBounds, out);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int notIncluded(const K* fLeaves, int count, const K* gLeaves, int gCount, int* flag, cudaStream_t s)
{
    if (count <= 0) { return 0; }
    notIncludedKernel<K><<<iceil(count, 256), 256, 0, s>>>(fLeaves, count, gLeaves, gCount, flag);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int checkTreelet(const K* treelet, int count, const K* leaves, int numLeaves, uint32_t* valid, cudaStream_t s)
{
    if (count <= 0) { return 0; }
    checkTreeletKernel<K><<<iceil(count, 256), 256, 0, s>>>(treelet, count, leaves, numLeaves, valid);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int splitTreelet(const K* keys, int count, const uint32_t* validScan, K* accepted, K* rejected, cudaStream_t s)
{
    if (count <= 0) { return 0; }
    splitTreeletKernel<K><<<iceil(count, 256), 256, 0, s>>>(keys, count, validScan, accepted, rejected);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int rejectLeaves(const K* rejected, int count, const K* leaves, int numKeys, int* nodeOps, cudaStream_t s)
{
    if (count <= 0) { return 0; }
    rejectLeavesKernel<K><<<iceil(count, 256), 256, 0, s>>>(rejected, count, leaves, numKeys, nodeOps);
    CSB_LAUNCH_CHECK();
    return 0;
}

int fillInt(int* a, int n, int v, cudaStream_t s)
{
    if (n <= 0) { return 0; }
    fillIntKernel<<<iceil(n, 256), 256, 0, s>>>(a, n, v);
    CSB_LAUNCH_CHECK();
    return 0;
}

template<class K>
int indexTreelet(const K* treelet, int numNodes, const K* prefixes, const int* levelRange, int* out, int* error,
                 cudaStream_t s)
{
    if (numNodes <= 0) { return 0; }
    indexTreeletKernel<K><<<iceil(numNodes, 256), 256, 0, s>>>(treelet, numNodes, prefixes, levelRange, out, error);
    CSB_LAUNCH_CHECK();
    return 0;
}

#define CSB_INST_LET(K)                                                                                                \
    template int focusBounds<K>(const K*, int, const K*, int, int*, cudaStream_t);                                     \
    template int notIncluded<K>(const K*, int, const K*, int, int*, cudaStream_t);                                     \
    template int checkTreelet<K>(const K*, int, const K*, int, uint32_t*, cudaStream_t);                               \
    template int splitTreelet<K>(const K*, int, const uint32_t*, K*, K*, cudaStream_t);                                \
    template int rejectLeaves<K>(const K*, int, const K*, int, int*, cudaStream_t);                                    \
    template int indexTreelet<K>(const K*, int, const K*, const int*, int*, int*, cudaStream_t);
CSB_INST_LET(uint32_t)
CSB_INST_LET(uint64_t)
#undef CSB_INST_LET

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
             const int* bnd, const K* focusNodes, int numFocusNodes, uint8_t* markings, cudaStream_t s, bool limitSource)
{
    if (numFocusNodes <= 0) { return 0; }
    Box<T> box = makeBox<T>(lim, bnd);
    markMacsKernel<K, T><<<iceil(numFocusNodes, 128), 128, 0, s>>>(prefixes, childOffsets, parents, centers4, box,
                                                                   focusNodes, numFocusNodes, limitSource, markings);
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

int gatherRangesWords(const uint32_t* rangeScan, const uint32_t* rangeStart, int numRanges, uint32_t total, int words,
                      const void* src, void* out, cudaStream_t s)
{
    if (total == 0) { return 0; }
    gatherRangesWordsKernel<<<iceil(size_t(total) * words, 256), 256, 0, s>>>(
        rangeScan, rangeStart, numRanges, total, words, static_cast<const uint32_t*>(src), static_cast<uint32_t*>(out));
    CSB_LAUNCH_CHECK();
    return 0;
}

int gatherWords(const uint32_t* order, uint32_t n, int words, const void* src, void* out, cudaStream_t s)
{
    if (n == 0) { return 0; }
    gatherWordsKernel<<<iceil(size_t(n) * words, 256), 256, 0, s>>>(order, n, words, static_cast<const uint32_t*>(src),
                                                                    static_cast<uint32_t*>(out));
    CSB_LAUNCH_CHECK();
    return 0;
}

int replayGatherWords(const uint32_t* order, uint32_t n, int words, const void* before, const void* received,
                      uint32_t recvStart, uint32_t numRecv, void* out, cudaStream_t s)
{
    if (n == 0) { return 0; }
    replayGatherWordsKernel<<<iceil(size_t(n) * words, 256), 256, 0, s>>>(
        order, n, words, static_cast<const uint32_t*>(before), static_cast<const uint32_t*>(received), recvStart,
        numRecv, static_cast<uint32_t*>(out));
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
                                       const uint32_t*, int, uint8_t*, cudaStream_t, bool);
template int markMacs<uint64_t, float>(const uint64_t*, const int*, const int*, const float*, const double*, const int*,
                                       const uint64_t*, int, uint8_t*, cudaStream_t, bool);
template int markMacs<uint64_t, double>(const uint64_t*, const int*, const int*, const double*, const double*,
                                        const int*, const uint64_t*, int, uint8_t*, cudaStream_t, bool);
template int rangeCount<uint32_t>(const uint32_t*, int, const uint64_t*, const uint32_t*, const int*, int, uint32_t*,
                                  cudaStream_t);
template int rangeCount<uint64_t>(const uint64_t*, int, const uint64_t*, const uint64_t*, const int*, int, uint32_t*,
                                  cudaStream_t);
template int gatherRanges4<float>(const uint32_t*, const uint32_t*, int, uint32_t, const float*, const float*,
                                  const float*, const float*, float*, size_t, cudaStream_t);
template int gatherRanges4<double>(const uint32_t*, const uint32_t*, int, uint32_t, const double*, const double*,
                                   const double*, const double*, double*, size_t, cudaStream_t);

} // namespace csb

extern "C"
{

/* extractMarkedElements (domain/layout.hpp:110-141) on the device: the request keys of the leaf range
 * [firstReqIdx, secondReqIdx) - one (first key, end key) pair per run of consecutive leaves with a non-zero layout
 * count.  Returns the number of keys written (2 per run), -1 on error, -needed if capacity is too small. */
#define CSB_EXTRACT_ABI(SFX, K)                                                                                        \
    long cs_extract_marked_elements_##SFX(const K* leaves, const uint32_t* layout, int numLeaves, int firstReqIdx,    \
                                          int secondReqIdx, K* out, long capacity, void* stream)                       \
    {                                                                                                                  \
        cudaStream_t s = cudaStream_t(stream);                                                                         \
        if (firstReqIdx < 0 || secondReqIdx > numLeaves || firstReqIdx > secondReqIdx)                                 \
        {                                                                                                              \
            csb::setLastError("cs_extract_marked_elements: invalid leaf range");                                       \
            return -1;                                                                                                 \
        }                                                                                                              \
        /* two "ranks": the caller owns nothing, the requested range is the peer's */                                 \
        const int fa[4] = {0, 0, firstReqIdx, secondReqIdx};                                                           \
        void* scratchA = csb::scratch(s, csb::SCRATCH_A, (size_t(numLeaves) + 1) * sizeof(uint32_t) + 64);             \
        void* scratchB = csb::scratch(s, csb::SCRATCH_B, csb::scanTempBytes(size_t(numLeaves) + 1));                    \
        void* scratchC = csb::scratch(s, csb::SCRATCH_C, 64);                                                          \
        if (!scratchA || !scratchB || !scratchC) { return -1; }                                                        \
        uint32_t* starts = static_cast<uint32_t*>(scratchA);                                                           \
        int* faDev       = static_cast<int*>(scratchC);                                                                \
        int* status      = faDev + 4;                                                                                  \
        if (cudaMemcpyAsync(faDev, fa, sizeof(fa), cudaMemcpyHostToDevice, s) != cudaSuccess ||                        \
            cudaMemsetAsync(status, 0, 2 * sizeof(int), s) != cudaSuccess)                                             \
        {                                                                                                              \
            return -1;                                                                                                 \
        }                                                                                                              \
        if (csb::haloRunStarts(layout, numLeaves, faDev, 2, 0, 0xFFFFFFFFu, starts, status, s)) { return -1; }         \
        if (csb::exclusiveScanU32(starts, starts, size_t(numLeaves) + 1, scratchB, s)) { return -1; }                  \
        uint32_t numRuns = 0;                                                                                          \
        if (cudaMemcpyAsync(&numRuns, starts + numLeaves, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess || \
            cudaStreamSynchronize(s) != cudaSuccess)                                                                   \
        {                                                                                                              \
            return -1;                                                                                                 \
        }                                                                                                              \
        if (long(2 * numRuns) > capacity) { return -long(2 * numRuns); }                                               \
        if (numRuns && csb::haloRequestKeys<K>(layout, numLeaves, faDev, 2, 0, starts, leaves, out, s)) { return -1; } \
        if (cudaStreamSynchronize(s) != cudaSuccess) { return -1; }                                                    \
        return long(2 * numRuns);                                                                                      \
    }
CSB_EXTRACT_ABI(u32, uint32_t)
CSB_EXTRACT_ABI(u64, uint64_t)
#undef CSB_EXTRACT_ABI

/* markMacsGpu (traversal/collisions_gpu.h:62-71, macs.hpp:185-229): markings[i] = 1 for every node of the linked tree
 * that fails the MAC against one of the numFocusNodes leaves focusNodes[0..numFocusNodes] (and is not contained in
 * their key range); centers4 = (x, y, z, mac^2) per node; limitSource: nodes deeper than one level above the target leaf
 * are neither marked nor entered.  Marks are only set, never cleared. */
int cs_mark_macs_u32f(const uint32_t* prefixes, const int* childOffsets, const int* parents, const float* centers4,
                      const double* lim, const int* bnd, const uint32_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream)
{
    return csb::markMacs<uint32_t, float>(prefixes, childOffsets, parents, centers4, lim, bnd, focusNodes, numFocusNodes,
                                          markings, cudaStream_t(stream), limitSource != 0);
}
int cs_mark_macs_u64f(const uint64_t* prefixes, const int* childOffsets, const int* parents, const float* centers4,
                      const double* lim, const int* bnd, const uint64_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream)
{
    return csb::markMacs<uint64_t, float>(prefixes, childOffsets, parents, centers4, lim, bnd, focusNodes, numFocusNodes,
                                          markings, cudaStream_t(stream), limitSource != 0);
}
int cs_mark_macs_u64d(const uint64_t* prefixes, const int* childOffsets, const int* parents, const double* centers4,
                      const double* lim, const int* bnd, const uint64_t* focusNodes, int numFocusNodes, int limitSource,
                      uint8_t* markings, void* stream)
{
    return csb::markMacs<uint64_t, double>(prefixes, childOffsets, parents, centers4, lim, bnd, focusNodes,
                                           numFocusNodes, markings, cudaStream_t(stream), limitSource != 0);
}

/* gatherRanges (halos/gather_halos_gpu.h:23-30): buffer[rangeScan[r] + k] = src[rangeOffsets[r] + k] for the numRanges
 * ranges whose lengths are the differences of rangeScan (the last one ends at bufferSize); elements of elemBytes bytes
 * (a multiple of 4: int, util::array<float, 1..4>) */
int cs_gather_ranges(const uint32_t* rangeScan, const uint32_t* rangeOffsets, int numRanges, const void* src,
                     void* buffer, size_t bufferSize, int elemBytes, void* stream)
{
    CSB_REQUIRE(elemBytes > 0 && elemBytes % 4 == 0, "gatherRanges: element sizes must be positive multiples of 4 bytes");
    return csb::gatherRangesWords(rangeScan, rangeOffsets, numRanges, uint32_t(bufferSize), elemBytes / 4, src, buffer,
                                  cudaStream_t(stream));
}

} // extern "C"
