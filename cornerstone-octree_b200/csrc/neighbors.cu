/* Radius neighbour search for sm_100a producing the reference's CPU list layout:
 * cstone::findNeighbors (findneighbors.hpp:77-177): neighbors[(i-first)*ngmax + k], neighborsCount[i-first].
 *
 * Design
 *  - Target groups are LEAF-ALIGNED: every tree leaf is cut into ceil(count/32) equal groups of at most 32 consecutive
 *    particles, and consecutive SIBLING leaves with few particles are packed into one group (the role of the
 *    reference's GroupView / computeFixedGroups, traversal/groups.hpp:28-64, but aligned to tree cells).  All targets
 *    of a group sit in one leaf or parent cell, so the union of the tree cells their search spheres touch is close to
 *    what a single target touches (27 cells instead of ~85 for arbitrary 32-particle SFC slices).
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
 *  - The default search (variant 2, second half of this file) shares the tree walk and the staging of the candidates
 *    between the four warps of a CTA; the per-warp search described above remains as variant 0 and as its fall-back.
 */
#include <type_traits>

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

/*! Groups are built per internal node from its leaf children.
 *  POLICY 0: consecutive sibling leaves are packed greedily into one group while they hold at most 32 targets together
 *  (deep trees have leaves with a handful of particles; a warp per such leaf would run mostly empty), a leaf with more
 *  than 32 targets is cut into ceil(count/32) balanced groups: groups never straddle a leaf boundary unless they hold
 *  whole leaves.
 *  POLICY 1: every maximal run of consecutive leaf siblings (their particles are contiguous) is cut into
 *  ceil(total/LIMIT) balanced groups regardless of the leaf boundaries inside the run: groups are full (a uniform tree
 *  with 32 particles per leaf gives 8 groups of 32 per parent instead of ~12 of 21), at the price of a slightly larger
 *  bounding box where a group takes particles from two curve-adjacent leaves.  LIMIT = 128 gives the super-groups of
 *  the cooperative search.
 *  FILL = false counts the groups that start at each leaf, FILL = true writes them at the scanned offsets. */
template<bool FILL, int POLICY, int LIMIT>
__global__ void groupBuildKernel(const int* __restrict__ childOffsets, const int* __restrict__ internalToLeaf,
                                 const uint32_t* __restrict__ layout, int numNodes, uint32_t first, uint32_t last,
                                 uint32_t* __restrict__ groupCounts, const uint32_t* __restrict__ groupOffsets,
                                 uint2* __restrict__ groups)
{
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= numNodes) { return; }

    auto standalone = [&](int leaf, uint32_t s, uint32_t e)
    {
        uint32_t c  = e - s;
        uint32_t ng = (c + LIMIT - 1) / LIMIT;
        if (!FILL) { groupCounts[leaf] = ng; }
        else
        {
            // balanced split: sizes differ by at most one, none is empty, all are <= 32
            uint32_t off = groupOffsets[leaf];
            for (uint32_t k = 0; k < ng; ++k)
                groups[off + k] = make_uint2(s + uint32_t(uint64_t(k) * c / ng), s + uint32_t(uint64_t(k + 1) * c / ng));
        }
    };

    const int child0 = childOffsets[node];
    if (child0 == 0)
    {
        if (numNodes == 1) // the root is the only leaf
        {
            uint32_t s, e;
            leafTargets(layout, 0, first, last, s, e);
            if (e > s) { standalone(0, s, e); }
        }
        return;
    }

    int packLeaf       = -1;
    uint32_t packStart = 0, packEnd = 0;
    auto flush = [&]()
    {
        if (packLeaf >= 0)
        {
            if (POLICY == 1) { standalone(packLeaf, packStart, packEnd); }
            else if (!FILL) { groupCounts[packLeaf] = 1; }
            else { groups[groupOffsets[packLeaf]] = make_uint2(packStart, packEnd); }
        }
        packLeaf = -1;
    };
    for (int j = 0; j < 8; ++j)
    {
        if (childOffsets[child0 + j] != 0)
        {
            flush();
            continue;
        }
        const int leaf = internalToLeaf[child0 + j];
        uint32_t s, e;
        leafTargets(layout, leaf, first, last, s, e);
        if (e == s) { continue; }
        if (POLICY == 1)
        {
            // runs of leaf siblings: contiguous particle ranges are merged, cut into groups when the run ends
            if (packLeaf >= 0 && packEnd == s) { packEnd = e; }
            else
            {
                flush();
                packLeaf  = leaf;
                packStart = s;
                packEnd   = e;
            }
            continue;
        }
        if (packLeaf >= 0 && packEnd == s && e - packStart <= 32) { packEnd = e; }
        else
        {
            flush();
            if (e - s <= 32)
            {
                packLeaf  = leaf;
                packStart = s;
                packEnd   = e;
            }
            else { standalone(leaf, s, e); }
        }
    }
    flush();
}

/* ---------------------------------------------------------------- traversal */

template<class T>
struct Target
{
    T x, y, z;
    T radiusSq;
    bool usePbc;
};

//! continuation test of findneighbors.hpp:108-112 in the reference's precision and operation order
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

/* ---- certified single-precision pre-filter for double-precision searches ----
 * B200 issues FP64 at half the FP32 rate and the search is instruction-issue bound, so for T = double every distance
 * test is first evaluated in float on coordinates taken RELATIVE to the first target of the warp (that removes the
 * magnitude of the box from the rounding error).  With every float coordinate difference within
 * e = k * 2^-24 * D of the true one (D: largest relative coordinate involved; k = 4 for particle pairs, 8 for
 * point-box distances which have more roundings), the float sum of squares s satisfies
 *      |s - d2| <= delta * s + 3 e^2 (1 + 1/delta)         (AM-GM on the cross term) + 4 * 2^-24 * s (float rounding)
 * so with delta = 2^-12:   s < A := (r2(1-2^-22) - E)(1-2^-11)  =>  d2 < r2   in the reference's double arithmetic
 *                          s > B := (r2(1+2^-22) + E)(1+2^-11)  =>  d2 >= r2
 * where E = 2^-29 D^2 (pairs) or 2^-27 D^2 (boxes) over-covers 3 e^2 (1 + 2^12).  Only candidates inside the band
 * [A, B] (a shell of relative width 2^-10 around the search sphere, < 1 % of the neighbours) are re-evaluated with the
 * reference's double expression, so the accepted set — and therefore list order, counts and truncation — is identical
 * to the CPU result bit for bit.  Degenerate magnitudes (radius^2 below 1e-30 in float, overflow, NaN) make the band
 * cover everything, i.e. fall back to the double expression.
 *
 * The same bound gives a warp-level cull: a candidate whose distance to the bounding box of the warp's targets is
 * certainly larger than the largest search radius of the warp cannot be a neighbour of any lane and is dropped while
 * staging (about 40 % of the particles of the 27 cells around a leaf), for float searches as well. */
constexpr float BAND_KA = (1.0f - 0x1p-22f) * (1.0f - 0x1p-11f);
constexpr float BAND_KB = (1.0f + 0x1p-22f) * (1.0f + 0x1p-11f);
constexpr float BAND_SA = 1.0f - 0x1p-11f;
constexpr float BAND_SB = 1.0f + 0x1p-11f;

//! order-preserving map float -> int (for REDUX min/max); NaNs map above +inf / below -inf by sign
__device__ inline int floatKey(float f)
{
    int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ inline float keyFloat(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ inline float warpMinF(float v) { return keyFloat(__reduce_min_sync(0xffffffffu, floatKey(v))); }
__device__ inline float warpMaxF(float v) { return keyFloat(__reduce_max_sync(0xffffffffu, floatKey(v))); }

//! the reference's acceptance test (findneighbors.hpp:33-60,134); out of line so that the rare call does not get
//! if-converted into the hot loop
template<class T>
__device__ __noinline__ bool exactInside(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                         uint32_t j, T tx, T ty, T tz, T radiusSq)
{
    T ex = x[j] - tx;
    T ey = y[j] - ty;
    T ez = z[j] - tz;
    return ex * ex + ey * ey + ez * ez < radiusSq;
}

constexpr int NB_MAX_DEPTH = 23; // >= maxTreeLevel<uint64_t> + 2
constexpr int NB_STAGE     = 64; // staged candidates per round (two half-rounds of 32 loads)

struct WarpShared
{
    float4 cand[NB_STAGE];          // x, y, z (relative floats for T = double, the values themselves for T = float), j
    float4 geoC[8], geoS[8];        // centres / sizes of the 8 children being tested; geoS.w = error term E of the child
    uint8_t mask[NB_MAX_DEPTH][32]; // per tree depth, per lane: which of the 8 siblings this lane's own walk enters
};

/*! The search of ONE warp for the targets [grp.x, grp.y) (at most 32), walking the tree itself: the kernel of variant 0
 *  and the fall-back of the cooperative search (periodic-boundary groups, range-list overflow).
 *  PBC = false: the box has no periodic dimension, the fold code is not even compiled in.  PBC = true: whether the
 *  fold is needed is decided per warp (any lane whose search sphere leaves the box); such warps run the reference
 *  expressions directly on broadcast loads, lanes that do not need the fold select the unfolded difference exactly as
 *  the reference picks per particle (findneighbors.hpp:104-106,150-151).  Interior warps take the staged path. */
template<class T, bool PBC>
__device__ __noinline__ void warpSearch(WarpShared& sh, const uint2 grp, const T* __restrict__ x, const T* __restrict__ y,
                           const T* __restrict__ z, const T* __restrict__ h, uint32_t first, const Box<T>& box,
                           const int* __restrict__ childOffsets, const int* __restrict__ parents,
                           const int* __restrict__ internalToLeaf, const uint32_t* __restrict__ layout,
                           const T* __restrict__ centers, const T* __restrict__ sizes, uint32_t ngmax,
                           uint32_t* __restrict__ neighbors, uint32_t* __restrict__ neighborsCount)
{
    constexpr bool Filt   = sizeof(T) == 8;
    const unsigned lane   = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const bool valid      = grp.x + lane < grp.y;
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

    // single-precision frame: relative to the first target of the group for double searches, absolute for float
    const T ox = Filt ? __shfl_sync(0xffffffffu, t.x, 0) : T(0);
    const T oy = Filt ? __shfl_sync(0xffffffffu, t.y, 0) : T(0);
    const T oz = Filt ? __shfl_sync(0xffffffffu, t.z, 0) : T(0);
    const float txf = float(t.x - ox);
    const float tyf = float(t.y - oy);
    const float tzf = float(t.z - oz);
    const float r2f = float(t.radiusSq);

    // bounding box of the targets, largest radius, and the magnitude bounds of the error terms (group constants)
    const float lox = warpMinF(txf), hix = warpMaxF(txf);
    const float loy = warpMinF(tyf), hiy = warpMaxF(tyf);
    const float loz = warpMinF(tzf), hiz = warpMaxF(tzf);
    const float DwT = fmaxf(fmaxf(fmaxf(fabsf(lox), fabsf(hix)), fmaxf(fabsf(loy), fabsf(hiy))),
                            fmaxf(fabsf(loz), fabsf(hiz)));
    const float r2max = warpMaxF(r2f);
    const float r2bMax = (r2max > 1e-30f) ? r2max * BAND_KB : __int_as_float(0x7f800000);

    float r2a = r2f * BAND_KA, r2b = r2f * BAND_KB;
    if (!(r2f > 1e-30f))
    {
        r2a = -1.0f;
        r2b = __int_as_float(0x7f800000);
    }
    // every staged (un-culled) candidate has |coordinate| <= 1.01 (DwT + sqrt(r2max)), see the cull test
    const float Dpair = 1.01f * (DwT + sqrtf(r2max));
    const float Epair = Dpair * Dpair * 0x1p-29f;
    const float pairA = fmaf(-Epair, BAND_SA, r2a);
    const float pairB = fmaf(Epair, BAND_SB, r2b);

    // out == row + numFound at all times; entries beyond ngmax are counted but not stored (findneighbors.hpp:139-146)
    uint32_t* out     = neighbors + size_t(i - first) * size_t(ngmax);
    uint32_t numFound = 0;

    auto append = [&](uint32_t j)
    {
        if (numFound < ngmax) { *out = j; }
        ++out;
        ++numFound;
    };

    auto scanLeaf = [&](int node, bool mine)
    {
        int leafIdx = internalToLeaf[node];
        uint32_t jb = layout[leafIdx];
        uint32_t je = layout[leafIdx + 1];
        if (warpPbc)
        {
            // warps touching a periodic boundary: the reference expressions on broadcast loads
            for (uint32_t j = jb; j < je; ++j)
            {
                T dx = x[j] - t.x;
                T dy = y[j] - t.y;
                T dz = z[j] - t.z;
                T fx = pbcFold(dx, 0, box);
                T fy = pbcFold(dy, 1, box);
                T fz = pbcFold(dz, 2, box);
                dx   = t.usePbc ? fx : dx;
                dy   = t.usePbc ? fy : dy;
                dz   = t.usePbc ? fz : dz;
                T d2 = dx * dx + dy * dy + dz * dz;
                if (mine && j != i && d2 < t.radiusSq) { append(j); }
            }
            return;
        }
        const float bandA = mine ? pairA : -1.0f; // not mine: never inside ...
        const float bandB = mine ? pairB : -1.0f; // ... and always surely outside (sums of squares are >= 0)
        for (uint32_t base = jb; base < je; base += NB_STAGE)
        {
            // stage up to 64 candidates with coalesced loads, dropping those no lane can reach
            uint32_t cnt = 0;
            __syncwarp();
#pragma unroll
            for (int half = 0; half < NB_STAGE / 32; ++half)
            {
                if (base + half * 32 >= je) { break; }
                const uint32_t j = base + half * 32 + lane;
                bool keep        = false;
                float4 c;
                if (j < je)
                {
                    c.x = float(x[j] - ox);
                    c.y = float(y[j] - oy);
                    c.z = float(z[j] - oz);
                    c.w = __uint_as_float(j);
                    float D  = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), DwT));
                    float bc = fmaf(D * D * 0x1p-27f, BAND_SB, r2bMax);
                    float ex = fmaxf(fmaxf(lox - c.x, c.x - hix), 0.0f);
                    float ey = fmaxf(fmaxf(loy - c.y, c.y - hiy), 0.0f);
                    float ez = fmaxf(fmaxf(loz - c.z, c.z - hiz), 0.0f);
                    keep     = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > bc);
                }
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                if (keep) { sh.cand[cnt + __popc(km & ltMask)] = c; }
                cnt += __popc(km);
            }
            __syncwarp();
            if (Filt)
            {
#pragma unroll 4
                for (uint32_t k = 0; k < cnt; ++k)
                {
                    const float4 c = sh.cand[k];
                    float dx       = c.x - txf;
                    float dy       = c.y - tyf;
                    float dz       = c.z - tzf;
                    float s2       = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    const uint32_t j = __float_as_uint(c.w);
                    bool in          = s2 < bandA;
                    // inside the uncertainty band (rare, < 1 % of the neighbours; never for lanes that do not own
                    // the leaf, their band is empty): the reference's own expression decides.  The branch is
                    // warp-uniform and the callee is evaluated by all lanes, so the common path carries no
                    // divergence bookkeeping.
                    const bool amb = !in && !(s2 > bandB);
                    if (__any_sync(0xffffffffu, amb))
                    {
                        const bool e = exactInside(x, y, z, j, t.x, t.y, t.z, t.radiusSq);
                        in           = amb ? e : in;
                    }
                    in = in && j != i;
                    if (in && numFound < ngmax) { *out = j; }
                    out += in;
                    numFound += in;
                }
            }
            else
            {
#pragma unroll 4
                for (uint32_t k = 0; k < cnt; ++k)
                {
                    const float4 c   = sh.cand[k];
                    const uint32_t j = __float_as_uint(c.w);
                    T dx = T(c.x) - t.x;
                    T dy = T(c.y) - t.y;
                    T dz = T(c.z) - t.z;
                    T d2 = dx * dx + dy * dy + dz * dz;
                    if (mine && j != i && d2 < t.radiusSq) { append(j); }
                }
            }
        }
    };

    //! this lane's continuation decisions for the 8 children of an internal node its own walk has entered
    auto testChildren = [&](int child0, bool mine) -> uint32_t
    {
        uint32_t bits = 0;
        if (warpPbc)
        {
            if (mine)
            {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    bits |= uint32_t(cellOverlap<PBC>(t, centers, sizes, child0 + c, box)) << c;
            }
            return bits;
        }
        __syncwarp();
        bool reach = false;
        if (lane < 8)
        {
            int node = child0 + int(lane);
            float4 gc, gs;
            gc.x = float(centers[3 * node] - ox);
            gc.y = float(centers[3 * node + 1] - oy);
            gc.z = float(centers[3 * node + 2] - oz);
            gc.w = 0.0f;
            gs.x = float(sizes[3 * node]);
            gs.y = float(sizes[3 * node + 1]);
            gs.z = float(sizes[3 * node + 2]);
            float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), DwT)),
                            fmaxf(gs.x, fmaxf(gs.y, gs.z)));
            gs.w          = D * D * 0x1p-27f;
            sh.geoC[lane] = gc;
            sh.geoS[lane] = gs;
            // warp-level cull: a child whose box is certainly farther from the targets' bounding box than the
            // largest radius fails the continuation test of every lane
            float ex = fmaxf(fmaxf(lox - (gc.x + gs.x), (gc.x - gs.x) - hix), 0.0f);
            float ey = fmaxf(fmaxf(loy - (gc.y + gs.y), (gc.y - gs.y) - hiy), 0.0f);
            float ez = fmaxf(fmaxf(loz - (gc.z + gs.z), (gc.z - gs.z) - hiz), 0.0f);
            reach    = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(gs.w, BAND_SB, r2bMax));
        }
        const unsigned reachable = __ballot_sync(0xffffffffu, reach);
        if (mine)
        {
#pragma unroll
            for (int c = 0; c < 8; ++c)
            {
                if (!((reachable >> c) & 1u)) { continue; }
                const float4 gc = sh.geoC[c];
                const float4 gs = sh.geoS[c];
                bool pass;
                if (Filt)
                {
                    float dx = fmaxf(fabsf(gc.x - txf) - gs.x, 0.0f);
                    float dy = fmaxf(fabsf(gc.y - tyf) - gs.y, 0.0f);
                    float dz = fmaxf(fabsf(gc.z - tzf) - gs.z, 0.0f);
                    float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    pass     = s2 < fmaf(-gs.w, BAND_SA, r2a);
                    if (!pass && !(s2 > fmaf(gs.w, BAND_SB, r2b)))
                    {
                        pass = cellOverlap<false>(t, centers, sizes, child0 + c, box);
                    }
                }
                else
                {
                    // float searches: the staged values are the reference's operands, evaluate its expression
                    T dx = rabs(T(gc.x) - t.x) - T(gs.x);
                    T dy = rabs(T(gc.y) - t.y) - T(gs.y);
                    T dz = rabs(T(gc.z) - t.z) - T(gs.z);
                    dx += rabs(dx);
                    dy += rabs(dy);
                    dz += rabs(dz);
                    dx *= T(0.5);
                    dy *= T(0.5);
                    dz *= T(0.5);
                    pass = dx * dx + (dy * dy + dz * dz) < t.radiusSq;
                }
                bits |= uint32_t(pass) << c;
            }
        }
        return bits;
    };

    const bool rootMine = valid && cellOverlap<PBC>(t, centers, sizes, 0, box);
    if (__any_sync(0xffffffffu, rootMine))
    {
        int rootChild = childOffsets[0];
        if (rootChild == 0) { scanLeaf(0, rootMine); }
        else
        {
            // depth-first walk in SFC order over the children that at least one lane enters: `lm` holds this lane's
            // decisions for the 8 siblings starting at `base` (kept per depth in shared memory for the way back up),
            // `wm` the siblings still to visit for the warp
            int depth        = 1;
            int base         = rootChild;
            uint32_t lm      = testChildren(rootChild, rootMine);
            sh.mask[1][lane] = uint8_t(lm);
            uint32_t wm      = __reduce_or_sync(0xffffffffu, lm);
            while (true)
            {
                if (wm == 0)
                {
                    if (depth == 1) { break; }
                    const int up = parents[(base - 1) >> 3];
                    --depth;
                    base = ((up - 1) & ~7) + 1;
                    lm   = sh.mask[depth][lane];
                    wm   = __reduce_or_sync(0xffffffffu, lm) & ~((2u << ((up - 1) & 7)) - 1u);
                    continue;
                }
                const int c = __ffs(int(wm)) - 1;
                wm &= wm - 1;
                const int node  = base + c;
                const bool mine = (lm >> c) & 1u;
                const int child = childOffsets[node];
                if (child == 0) { scanLeaf(node, mine); }
                else
                {
                    ++depth;
                    lm                   = testChildren(child, mine);
                    sh.mask[depth][lane] = uint8_t(lm);
                    wm                   = __reduce_or_sync(0xffffffffu, lm);
                    base                 = child;
                }
            }
        }
    }

    if (valid) { neighborsCount[i - first] = numFound; }
}

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
    __shared__ WarpShared shAll[NB_THREADS / 32];
    const size_t warpId = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    if (warpId >= size_t(*numGroupsPtr)) { return; }
    warpSearch<T, PBC>(shAll[threadIdx.x >> 5], groups[warpId], x, y, z, h, first, box, childOffsets, parents,
                       internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount);
}

/* ================================================================ cooperative search (variant 2, the default)
 *
 * The per-warp walk above spends most of its instructions outside the distance tests: every warp traverses the tree
 * for itself, converts and culls the particles of ~50 leaf visits, and the sibling warps of a parent cell repeat nearly
 * the same work.  Here a CTA of four warps owns a SUPER-GROUP of up to 128 consecutive targets taken from a run of
 * sibling leaves:
 *   1. nbRangesKernel (one warp per super-group, all lanes busy with the 8 children of a node at a time) walks the
 *      tree ONCE for the bounding box of the super-group and records the particle ranges of the leaves that any of
 *      its targets can reach, in SFC order.  No per-target tests are made here.
 *   2. nbSearchKernel stages those ranges chunk by chunk in shared memory (one load + conversion per particle and
 *      CTA); each warp (32 targets in lanes) culls the staged leaves and particles against ITS bounding box, buffers
 *      the survivors and tests them in batches of 128 with packed two-wide single-precision arithmetic, one result bit
 *      per candidate; accepted bits are written out in buffer order = ascending particle index.
 *
 * Exactness without per-target tree tests.  The reference accepts particle j for target i iff the continuation test
 * passes for EVERY node on the path from the root to j's leaf and d2(i,j) < r2 (findneighbors.hpp:108-146).  If j lies
 * in its leaf's box and every node's box lies in its parent's box, each of those boxes is at most |x_i - x_j| away from
 * target i, so d(i,j) < r - Delta implies that all continuation tests pass, Delta covering the rounding of the
 * reference's box arithmetic.  Both containments are CHECKED here (particles against their leaf box while they are
 * staged, child boxes against parent boxes during the walk, each with an explicit tolerance that is part of Delta);
 * whatever fails the check is marked and takes the exact route.  Candidates that are not certainly inside by that
 * margin - a shell of relative width ~2^-11 below the search radius - are decided by the reference's own expressions:
 * the double-precision distance and the continuation tests of j's leaf and all its ancestors (exactMine).  The result
 * is therefore identical to the per-target walk for ANY input, not only for consistent trees.
 * Super-groups with a target whose search sphere crosses a periodic boundary, or whose range list overflows, are
 * searched by warpSearch. */

constexpr int SG_TARGETS = 128;             // targets per super-group = threads of a search CTA
constexpr int SG_WARPS   = SG_TARGETS / 32;
constexpr int SG_CAP     = 192;             // range-list entries per super-group
constexpr int SG_SPLIT   = 64;              // particles per list entry (larger leaves are cut)
constexpr int SG_CHUNK   = 1024;            // staged particles per round (>= 16 entries)
constexpr int SG_BATCH   = 128;             // candidates per evaluation batch of a warp
constexpr int SG_PCAP    = SG_BATCH + 32;   // buffered candidates of a warp
constexpr uint32_t SG_UNCERT = 0x80000000u; // entry / particle failed a containment check: exact route
constexpr int SG_FLOAT_DEPTH = 10;          // float searches: depth covered by the tolerance budget (deeper: exact route)

//! tolerances of the containment checks and the resulting margin Delta, as multiples of the coordinate magnitude
template<class T>
struct CertTol;
template<>
struct CertTol<double>
{
    static constexpr double check = 0x1p-46; // per check (particle in leaf box, child box in parent box)
    static constexpr double delta = 0x1p-40; // >= (1 + 21 levels) * (check + rounding of the check) + box arithmetic
};
template<>
struct CertTol<float>
{
    static constexpr double check = 0x1p-21; // 8 ulp of the magnitude
    static constexpr double delta = 0x1p-17; // >= (1 + 10 levels) * 11 ulp + 4 ulp
};

__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, uint32_t& lo, uint32_t& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
//! two-wide IEEE single precision (round to nearest even per component): SASS FADD2 / FMUL2 / FFMA2
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

//! bounding box / radius / frame of a set of targets in the single-precision frame of the search
struct ReachBox
{
    float lox, hix, loy, hiy, loz, hiz; // bounding box of the targets
    float DwT;                          // largest |coordinate| of a target
    float r2bMax;                       // upper band edge of the largest search radius
};

//! can a box (centre gc, half sizes gs, error term gs.w) contain a neighbour of any target inside rb?  Certified: false
//! only if the box is farther from the targets' bounding box than the largest radius, whatever the rounding
__device__ __forceinline__ bool boxReachable(const ReachBox& rb, const float4& gc, const float4& gs)
{
    float ex = fmaxf(fmaxf(rb.lox - (gc.x + gs.x), (gc.x - gs.x) - rb.hix), 0.0f);
    float ey = fmaxf(fmaxf(rb.loy - (gc.y + gs.y), (gc.y - gs.y) - rb.hiy), 0.0f);
    float ez = fmaxf(fmaxf(rb.loz - (gc.z + gs.z), (gc.z - gs.z) - rb.hiz), 0.0f);
    return !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(gs.w, BAND_SB, rb.r2bMax));
}

//! node box in the single-precision frame with origin (ox,oy,oz); gs.w = error term of tests against this box
template<class T>
__device__ __forceinline__ void relBox(const T* __restrict__ centers, const T* __restrict__ sizes, int node, T ox, T oy,
                                       T oz, float DwT, float4& gc, float4& gs)
{
    gc.x = float(centers[3 * node] - ox);
    gc.y = float(centers[3 * node + 1] - oy);
    gc.z = float(centers[3 * node + 2] - oz);
    gc.w = 0.0f;
    gs.x = float(sizes[3 * node]);
    gs.y = float(sizes[3 * node + 1]);
    gs.z = float(sizes[3 * node + 2]);
    float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), DwT)), fmaxf(gs.x, fmaxf(gs.y, gs.z)));
    gs.w    = D * D * 0x1p-27f;
}

//! does the search sphere of a target with these coordinates leave the box in a periodic dimension?
template<class T>
__device__ __forceinline__ bool needsPbc(T tx, T ty, T tz, T hi, const Box<T>& box)
{
    bool anyPbc = box.pbc(0) || box.pbc(1) || box.pbc(2);
    T s         = T(2) * hi;
    bool inside = (tx - s >= box.lim[0]) && (ty - s >= box.lim[2]) && (tz - s >= box.lim[4]) &&
                  (tx + s <= box.lim[1]) && (ty + s <= box.lim[3]) && (tz + s <= box.lim[5]);
    return anyPbc && !inside;
}

//! coordinate magnitude that scales the tolerances: the box and the target itself
template<class T>
__device__ __forceinline__ double coordMagnitude(const Box<T>& box, T tx, T ty, T tz)
{
    double m = fmax(fmax(fabs(double(tx)), fabs(double(ty))), fabs(double(tz)));
    for (int k = 0; k < 6; ++k)
        m = fmax(m, fabs(double(box.lim[k])));
    return m;
}

struct RangeWarpShared
{
    int co[NB_MAX_DEPTH][8];       // child offsets of the sibling group on the path at each depth (0: leaf)
    uint32_t jb[NB_MAX_DEPTH][8];  // first particle of the leaf children
    uint32_t cnt[NB_MAX_DEPTH][8]; // their particle counts
    uint8_t wm[NB_MAX_DEPTH];      // siblings the walk enters
};

/*! range lists: entries[sg * SG_CAP + k] = {first particle, count (<= SG_SPLIT) | SG_UNCERT, leaf node, 0} for the
 *  leaves that can hold a neighbour of a target of super-group sg, in SFC order; header[sg] = number of entries, or -1
 *  if the super-group is to be searched by warpSearch (periodic boundary, overflow) */
template<class T, bool PBC>
__global__ void __launch_bounds__(SG_TARGETS) nbRangesKernel(const T* __restrict__ x,
                                                             const T* __restrict__ y,
                                                             const T* __restrict__ z,
                                                             const T* __restrict__ h,
                                                             const uint2* __restrict__ superGroups,
                                                             const uint32_t* __restrict__ numSuperGroupsPtr,
                                                             Box<T> box,
                                                             const int* __restrict__ childOffsets,
                                                             const int* __restrict__ parents,
                                                             const int* __restrict__ internalToLeaf,
                                                             const uint32_t* __restrict__ layout,
                                                             const T* __restrict__ centers,
                                                             const T* __restrict__ sizes,
                                                             uint4* __restrict__ entries,
                                                             int* __restrict__ header)
{
    __shared__ RangeWarpShared shAll[SG_WARPS];
    const unsigned lane = threadIdx.x & 31;
    const size_t sg     = (size_t(blockIdx.x) * SG_TARGETS + threadIdx.x) >> 5;
    if (sg >= size_t(*numSuperGroupsPtr)) { return; }
    RangeWarpShared& sh = shAll[threadIdx.x >> 5];
    const uint2 grp     = superGroups[sg];
    constexpr bool Filt = sizeof(T) == 8;

    // ---- bounding box, largest radius and coordinate magnitude of the (up to 128) targets: 4 per lane
    const T ox = Filt ? x[grp.x] : T(0), oy = Filt ? y[grp.x] : T(0), oz = Filt ? z[grp.x] : T(0);
    float lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f, loz = 3.0e38f, hiz = -3.0e38f, r2max = 0.0f;
    double mag  = 0.0;
    bool usePbc = false;
    for (uint32_t i = grp.x + lane; i < grp.y; i += 32)
    {
        const T tx = x[i], ty = y[i], tz = z[i], hi = h[i];
        const float fx = float(tx - ox), fy = float(ty - oy), fz = float(tz - oz);
        lox   = fminf(lox, fx), hix = fmaxf(hix, fx);
        loy   = fminf(loy, fy), hiy = fmaxf(hiy, fy);
        loz   = fminf(loz, fz), hiz = fmaxf(hiz, fz);
        r2max = fmaxf(r2max, float(T(4.0) * hi * hi));
        mag   = fmax(mag, coordMagnitude(box, tx, ty, tz));
        usePbc = usePbc || (PBC && needsPbc(tx, ty, tz, hi, box));
    }
    if (PBC && __any_sync(0xffffffffu, usePbc))
    {
        if (lane == 0) { header[sg] = -1; }
        return;
    }
    ReachBox rb;
    rb.lox = warpMinF(lox), rb.hix = warpMaxF(hix);
    rb.loy = warpMinF(loy), rb.hiy = warpMaxF(hiy);
    rb.loz = warpMinF(loz), rb.hiz = warpMaxF(hiz);
    rb.DwT = fmaxf(fmaxf(fmaxf(fabsf(rb.lox), fabsf(rb.hix)), fmaxf(fabsf(rb.loy), fabsf(rb.hiy))),
                   fmaxf(fabsf(rb.loz), fabsf(rb.hiz)));
    r2max     = warpMaxF(r2max);
    rb.r2bMax = (r2max > 1e-30f) ? r2max * BAND_KB : __int_as_float(0x7f800000);
    {
        // the magnitude is a double: reduce it through its order-preserving float upper bound
        float mf = __double2float_ru(mag);
        mag      = double(warpMaxF(mf));
    }
    const T tolCheck = T(CertTol<T>::check * mag);

    uint4* const list = entries + sg * size_t(SG_CAP);
    int numEntries    = 0;
    bool overflow     = false;

    //! append the particles [jb, jb + cnt) of leaf `node`, cut into entries of at most SG_SPLIT particles
    auto emitLeaf = [&](int node, uint32_t jb, uint32_t cnt, bool uncertified)
    {
        const int ne = int((cnt + SG_SPLIT - 1) / SG_SPLIT);
        if (numEntries + ne > SG_CAP)
        {
            overflow = true;
            return;
        }
        for (int k = int(lane); k < ne; k += 32)
        {
            uint32_t c = min(uint32_t(SG_SPLIT), cnt - uint32_t(k) * SG_SPLIT);
            list[numEntries + k] =
                make_uint4(jb + uint32_t(k) * SG_SPLIT, c | (uncertified ? SG_UNCERT : 0u), uint32_t(node), 0u);
        }
        numEntries += ne;
    };

    /*! the 8 children of `parent`, one per lane: child offsets, reach test against the targets' bounding box,
     *  containment of the child boxes in the parent's box, particle ranges of the leaf children.  Returns the mask of
     *  children the walk enters; `bad` is set if a child box sticks out of the parent box. */
    auto enterGroup = [&](int parent, int child0, int depth, bool& bad) -> uint32_t
    {
        bool reach = false, out = false;
        int co     = 0;
        uint32_t jb = 0, cnt = 0;
        // parent box: lanes 8..13 load one value each
        T pv = T(0);
        if (lane >= 8 && lane < 14)
        {
            pv = lane < 11 ? centers[3 * parent + (lane - 8)] : sizes[3 * parent + (lane - 11)];
        }
        const T pcx = __shfl_sync(0xffffffffu, pv, 8), pcy = __shfl_sync(0xffffffffu, pv, 9),
                pcz = __shfl_sync(0xffffffffu, pv, 10), psx = __shfl_sync(0xffffffffu, pv, 11),
                psy = __shfl_sync(0xffffffffu, pv, 12), psz = __shfl_sync(0xffffffffu, pv, 13);
        if (lane < 8)
        {
            const int node = child0 + int(lane);
            co             = childOffsets[node];
            const T cx = centers[3 * node], cy = centers[3 * node + 1], cz = centers[3 * node + 2];
            const T sx = sizes[3 * node], sy = sizes[3 * node + 1], sz = sizes[3 * node + 2];
            float4 gc, gs;
            gc.x = float(cx - ox), gc.y = float(cy - oy), gc.z = float(cz - oz), gc.w = 0.0f;
            gs.x = float(sx), gs.y = float(sy), gs.z = float(sz);
            float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), rb.DwT)),
                            fmaxf(gs.x, fmaxf(gs.y, gs.z)));
            gs.w    = D * D * 0x1p-27f;
            reach   = boxReachable(rb, gc, gs);
            out     = !(rabs(cx - pcx) + sx <= psx + tolCheck && rabs(cy - pcy) + sy <= psy + tolCheck &&
                    rabs(cz - pcz) + sz <= psz + tolCheck);
            if (reach && co == 0)
            {
                const int leafIdx = internalToLeaf[node];
                jb                = layout[leafIdx];
                cnt               = layout[leafIdx + 1] - jb;
            }
            sh.co[depth][lane]  = co;
            sh.jb[depth][lane]  = jb;
            sh.cnt[depth][lane] = cnt;
        }
        const uint32_t reachMask = __ballot_sync(0xffffffffu, reach) & 0xffu;
        bad                      = (__ballot_sync(0xffffffffu, out) & reachMask) != 0;
        if (lane == 0) { sh.wm[depth] = uint8_t(reachMask); }
        __syncwarp();
        return reachMask;
    };

    const int rootChild = childOffsets[0];
    if (rootChild == 0)
    {
        // the root is the only leaf
        emitLeaf(0, layout[0], layout[1] - layout[0], false);
    }
    else
    {
        int depth    = 1;
        int base     = rootChild;
        int uncDepth = 0; // smallest depth at which a containment check failed on the current path (0: none)
        bool bad;
        uint32_t wm = enterGroup(0, rootChild, 1, bad);
        if (bad) { uncDepth = 1; }
        while (!overflow)
        {
            if (wm == 0)
            {
                if (depth == 1) { break; }
                const int up = parents[(base - 1) >> 3];
                --depth;
                if (depth < uncDepth) { uncDepth = 0; }
                base = ((up - 1) & ~7) + 1;
                wm   = uint32_t(sh.wm[depth]) & ~((2u << ((up - 1) & 7)) - 1u);
                continue;
            }
            const int c = __ffs(int(wm)) - 1;
            wm &= wm - 1;
            const int node  = base + c;
            const int child = sh.co[depth][c];
            if (child == 0)
            {
                const uint32_t cnt = sh.cnt[depth][c];
                const bool unc     = uncDepth != 0 || (!Filt && depth > SG_FLOAT_DEPTH);
                if (cnt) { emitLeaf(node, sh.jb[depth][c], cnt, unc); }
            }
            else
            {
                ++depth;
                wm   = enterGroup(node, child, depth, bad);
                base = child;
                if (bad && uncDepth == 0) { uncDepth = depth; }
            }
        }
    }
    __syncwarp();
    if (lane == 0) { header[sg] = overflow ? -1 : numEntries; }
}

struct alignas(16) SearchWarpShared
{
    float cx[SG_PCAP], cy[SG_PCAP], cz[SG_PCAP]; // buffered candidates of the warp
    uint32_t cj[SG_PCAP];                        // particle index
    int cnode[SG_PCAP];                          // leaf node (for the exact route)
    uint32_t res[SG_BATCH / 32][32];             // accepted masks of the batch, per word and lane
    uint32_t unc[SG_PCAP / 32];                  // buffered candidates that failed the containment check
};

struct alignas(16) SearchShared
{
    float4 chunk[SG_CHUNK]; // staged particles: x, y, z in the single-precision frame, bits of (index | SG_UNCERT)
    float4 rBoxC[32], rBoxS[32];
    uint32_t rJb[32], rCnt[32], rOff[32];
    int rNode[32];
    int numRanges;
    float magW[SG_WARPS];
    SearchWarpShared w[SG_WARPS];
};
static_assert(sizeof(WarpShared) * SG_WARPS <= sizeof(float4) * SG_CHUNK, "warpSearch scratch overlays the chunk buffer");

//! the continuation tests of the reference for the leaf `node` and all its ancestors (findneighbors.hpp:108-112)
template<class T>
__device__ __noinline__ bool exactMine(const Target<T>& t, const T* __restrict__ centers, const T* __restrict__ sizes,
                                       const int* __restrict__ parents, int node, const Box<T>& box)
{
    for (int a = node;;)
    {
        if (!cellOverlap<false>(t, centers, sizes, a, box)) { return false; }
        if (a == 0) { return true; }
        a = parents[(a - 1) >> 3];
    }
}

/*! the reference's own decision for candidate j of target i (leaf `node`): the distance test in T
 *  (findneighbors.hpp:33-60,134), then - unless the distance is below the certified margin in T as well - the
 *  continuation tests of the leaf and all its ancestors.  tolCheck = CertTol::check * coordinate magnitude. */
template<class T>
__device__ __noinline__ bool exactDecision(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                           const T* __restrict__ h, uint32_t i, uint32_t j, int node, bool flagged,
                                           T tolCheck, const T* __restrict__ centers, const T* __restrict__ sizes,
                                           const int* __restrict__ parents, const Box<T>& box)
{
    Target<T> t;
    t.x        = x[i];
    t.y        = y[i];
    t.z        = z[i];
    const T hi = h[i];
    t.radiusSq = T(4.0) * hi * hi;
    t.usePbc   = false;
    const T ex = x[j] - t.x, ey = y[j] - t.y, ez = z[j] - t.z;
    const T d2 = ex * ex + ey * ey + ez * ez;
    if (!(d2 < t.radiusSq)) { return false; }
    if (sizeof(T) == 8 && !flagged)
    {
        const double mag   = double(tolCheck) / CertTol<T>::check;
        const double vSure = sqrt(double(t.radiusSq)) * (1.0 - 0x1p-20) - CertTol<T>::delta * mag;
        if (vSure > 0.0 && double(d2) < vSure * vSure * (1.0 - 0x1p-19)) { return true; }
    }
    return exactMine(t, centers, sizes, parents, node, box);
}

//! if (s < thr [&& j != self] [&& out < end]) *out = j; if (s < thr [&& j != self]) ++out  - predicated, no branches
template<bool SELF, bool GUARD>
__device__ __forceinline__ void appendIf(unsigned long long& out, unsigned long long end, uint32_t j, float s,
                                         float thr, uint32_t self)
{
    if (SELF && GUARD)
    {
        asm volatile("{\n .reg .pred q, g;\n setp.lt.f32 q, %2, %3;\n setp.ne.and.u32 q, %1, %4, q;\n"
                     " setp.lt.and.u64 g, %0, %5, q;\n @g st.global.u32 [%0], %1;\n @q add.u64 %0, %0, 4;\n}"
                     : "+l"(out)
                     : "r"(j), "f"(s), "f"(thr), "r"(self), "l"(end)
                     : "memory");
    }
    else if (SELF)
    {
        asm volatile("{\n .reg .pred q;\n setp.lt.f32 q, %2, %3;\n setp.ne.and.u32 q, %1, %4, q;\n"
                     " @q st.global.u32 [%0], %1;\n @q add.u64 %0, %0, 4;\n}"
                     : "+l"(out)
                     : "r"(j), "f"(s), "f"(thr), "r"(self)
                     : "memory");
    }
    else if (GUARD)
    {
        asm volatile("{\n .reg .pred q, g;\n setp.lt.f32 q, %2, %3;\n setp.lt.and.u64 g, %0, %4, q;\n"
                     " @g st.global.u32 [%0], %1;\n @q add.u64 %0, %0, 4;\n}"
                     : "+l"(out)
                     : "r"(j), "f"(s), "f"(thr), "l"(end)
                     : "memory");
    }
    else
    {
        asm volatile("{\n .reg .pred q;\n setp.lt.f32 q, %2, %3;\n @q st.global.u32 [%0], %1;\n"
                     " @q add.u64 %0, %0, 4;\n}"
                     : "+l"(out)
                     : "r"(j), "f"(s), "f"(thr)
                     : "memory");
    }
}

/* ================================================================ certified per-warp search (variant 1, the default)
 *
 * Same work distribution as warpSearch - one warp, up to 32 targets in lanes, one walk of the tree for all of them -
 * but without the per-lane continuation tests: the walk is steered by the warp's bounding box alone (8 lanes test the
 * 8 children of a node), and whether a target's OWN walk would have reached the leaf of an accepted particle follows
 * from the containment argument at the top of the cooperative section (CertTol, exactMine): candidates that are
 * inside by the margin Delta are accepted directly, the thin shell below the search radius and everything that fails
 * a containment check is decided by the reference's own expressions.  Candidates are tested four at a time with
 * packed two-wide arithmetic; the stores follow in candidate order, so lists stay in ascending particle index.
 * Per-lane state is kept small (the double-precision target, frame origin and leaf box live in shared / global
 * memory and are only touched on the rare exact route and while staging). */

constexpr int CW_STAGE = 64; // staged candidates per round (two half-rounds of 32 loads)

template<class T>
struct alignas(16) CertWarpShared
{
    float cx[CW_STAGE + 4], cy[CW_STAGE + 4], cz[CW_STAGE + 4]; // + padding up to a multiple of four
    uint32_t cj[CW_STAGE + 4];
    T org[4];                      // origin of the single-precision frame, tolerance of the containment checks
    T lbox[6];                     // centre and half sizes of the leaf being staged
    int co[NB_MAX_DEPTH][8];       // sibling groups on the current path: child offsets (0: leaf),
    uint32_t jb[NB_MAX_DEPTH][8];  // first particle and
    uint32_t cnt[NB_MAX_DEPTH][8]; // particle count of the leaf children,
    uint8_t wm[NB_MAX_DEPTH];      // siblings the walk enters
};

template<class T>
union CertWarpUnion
{
    CertWarpShared<T> cert;
    WarpShared legacy; // groups that touch a periodic boundary are searched by warpSearch
};

template<class T, bool PBC>
__global__ void __launch_bounds__(NB_THREADS, 8) findNeighborsCertKernel(const T* __restrict__ x,
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
    constexpr bool Filt = sizeof(T) == 8;
    __shared__ CertWarpUnion<T> shAll[NB_THREADS / 32];
    const size_t warpId = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    if (warpId >= size_t(*numGroupsPtr)) { return; }
    const uint2 grp       = groups[warpId];
    const unsigned lane   = threadIdx.x & 31;
    const unsigned ltMask = (1u << lane) - 1u;
    const bool valid      = grp.x + lane < grp.y;
    const uint32_t i      = valid ? grp.x + lane : grp.y - 1;
    CertWarpShared<T>& sh = shAll[threadIdx.x >> 5].cert;

    ReachBox rb;
    float thrLo, thrHi;         // see below
    uint32_t bandLo, bandSpan;  // bit patterns: a sum s >= 0 is undecided iff bits(s) - bandLo <= bandSpan (unsigned)
    uint64_t ntx2, nty2, ntz2;
    {
        const T tx = x[i], ty = y[i], tz = z[i], hi = h[i];
        if (PBC && __any_sync(0xffffffffu, needsPbc(tx, ty, tz, hi, box)))
        {
            warpSearch<T, PBC>(shAll[threadIdx.x >> 5].legacy, grp, x, y, z, h, first, box, childOffsets, parents,
                               internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount);
            return;
        }
        const T radiusSq = T(4.0) * hi * hi;
        // single-precision frame: relative to the first target of the group for double searches, absolute for float
        const T ox = Filt ? __shfl_sync(0xffffffffu, tx, 0) : T(0);
        const T oy = Filt ? __shfl_sync(0xffffffffu, ty, 0) : T(0);
        const T oz = Filt ? __shfl_sync(0xffffffffu, tz, 0) : T(0);
        const float txf = float(tx - ox);
        const float tyf = float(ty - oy);
        const float tzf = float(tz - oz);
        const float r2f = float(radiusSq);
        rb.lox = warpMinF(txf), rb.hix = warpMaxF(txf);
        rb.loy = warpMinF(tyf), rb.hiy = warpMaxF(tyf);
        rb.loz = warpMinF(tzf), rb.hiz = warpMaxF(tzf);
        rb.DwT = fmaxf(fmaxf(fmaxf(fabsf(rb.lox), fabsf(rb.hix)), fmaxf(fabsf(rb.loy), fabsf(rb.hiy))),
                       fmaxf(fabsf(rb.loz), fabsf(rb.hiz)));
        const float r2max = warpMaxF(r2f);
        rb.r2bMax         = (r2max > 1e-30f) ? r2max * BAND_KB : __int_as_float(0x7f800000);

        // thresholds (see nbSearchKernel): surely inside and reached below thrLo; T = double: surely outside above
        // thrHi, T = float: thrHi = radiusSq, the reference's own comparison on its own float expression
        const double mag    = double(warpMaxF(__double2float_ru(coordMagnitude(box, tx, ty, tz))));
        const double rD     = sqrt(double(radiusSq));
        const double vSure  = rD * (1.0 - 0x1p-20) - CertTol<T>::delta * mag;
        const double r2sure = vSure > 0.0 ? vSure * vSure * (1.0 - 0x1p-19) : -1.0;
        const float Dpair   = 1.01f * (rb.DwT + sqrtf(r2max));
        const float Epair   = Dpair * Dpair * 0x1p-29f;
        if (Filt)
        {
            const float r2sf = __double2float_rd(r2sure);
            thrLo            = (valid && r2sf > 1e-30f) ? fmaf(-Epair, BAND_SA, r2sf * BAND_KA) : -1.0f;
            thrHi            = (r2f > 1e-30f) ? fmaf(Epair, BAND_SB, r2f * BAND_KB) : __int_as_float(0x7f800000);
            if (!valid) { thrHi = -1.0f; } // padding lanes: everything is surely outside
        }
        else
        {
            thrLo = valid ? __double2float_rd(r2sure) : -1.0f;
            thrHi = valid ? r2f : -1.0f;
        }
        // a sum s (>= 0, or NaN / +inf) is undecided iff thrLo <= s <= thrHi: non-negative floats order like their bit
        // patterns, NaN patterns lie above +inf and never match
        if (thrHi >= 0.0f)
        {
            bandLo                = __float_as_uint(fmaxf(thrLo, 0.0f));
            const uint32_t bandHi = __float_as_uint(thrHi);
            bandSpan              = bandHi >= bandLo ? bandHi - bandLo : 0u;
            if (bandHi < bandLo) { bandLo = 0xffffffffu; }
        }
        else
        {
            bandLo   = 0xffffffffu;
            bandSpan = 0u;
        }
        ntx2 = pack2(-txf, -txf), nty2 = pack2(-tyf, -tyf), ntz2 = pack2(-tzf, -tzf);
        if (lane == 0)
        {
            sh.org[0] = ox, sh.org[1] = oy, sh.org[2] = oz;
            sh.org[3] = T(CertTol<T>::check * mag);
        }
        __syncwarp();
    }
    const uint64_t zero2 = pack2(0.0f, 0.0f);

    // out == row + number of neighbours found so far; entries at and beyond rowEnd are counted but not stored
    // (global-space byte addresses: the stores are issued from inline PTX)
    unsigned long long out       = __cvta_generic_to_global(neighbors + size_t(i - first) * size_t(ngmax));
    const unsigned long long rowEnd = out + 4ull * ngmax;

    //! squared distances of this lane's target to the staged candidates k, k + 1 (packed)
    auto dist2 = [&](uint64_t x2, uint64_t y2, uint64_t z2) -> uint64_t
    {
        const uint64_t dx = add2(x2, ntx2), dy = add2(y2, nty2), dz = add2(z2, ntz2);
        if (Filt) { return fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))); }
        // the reference's float expression (dx*dx + dy*dy) + dz*dz (findneighbors.hpp:33-60); products as fma(d, d, +0)
        // = RN(d*d): separate mul.rn / add.rn.f32x2 get contracted to FFMA2 by ptxas even under --fmad=false
        return add2(add2(fma2(dx, dx, zero2), fma2(dy, dy, zero2)), fma2(dz, dz, zero2));
    };

    //! the reference's own decision for candidate j of leaf `node` (rare, out of line)
    auto exact = [&](uint32_t j, int node, bool flagged) -> bool
    {
        return exactDecision<T>(x, y, z, h, i, j, node, flagged, sh.org[3], centers, sizes, parents, box);
    };

    //! one candidate on the careful route (blocks with a possibly undecided sum or a flagged candidate; rare)
    auto carefulOne = [&](float s, uint32_t kk, uint32_t cnt, uint32_t j, int node, bool f)
    {
        bool in              = !f && s < thrLo;
        const bool undecided = f || (!in && (Filt ? !(s > thrHi) : !(s >= thrHi)));
        if (undecided && valid && kk < cnt) { in = exact(j, node, f); }
        if (in && j != i)
        {
            if (out < rowEnd) { asm volatile("st.global.u32 [%0], %1;" ::"l"(out), "r"(j) : "memory"); }
            out += 4;
        }
    };

    /*! tests the `cnt` staged candidates (all from leaf `node`) four at a time.  SELF: the leaf holds targets of this
     *  group, so a candidate can be the target itself (excluded by index, findneighbors.hpp:131).  GUARD: a list may
     *  reach ngmax during this call.  flagged: staged candidates that failed a containment check (bit k = entry k) */
    auto testStaged = [&](uint32_t cnt, int node, uint64_t flagged, auto selfTag, auto guardTag)
    {
        constexpr bool SELF  = decltype(selfTag)::value;
        constexpr bool GUARD = decltype(guardTag)::value;
        for (uint32_t k = 0; k < cnt; k += 4)
        {
            const float4 X = *reinterpret_cast<const float4*>(&sh.cx[k]);
            const float4 Y = *reinterpret_cast<const float4*>(&sh.cy[k]);
            const float4 Z = *reinterpret_cast<const float4*>(&sh.cz[k]);
            const uint4 J  = *reinterpret_cast<const uint4*>(&sh.cj[k]);
            float s0, s1, s2, s3;
            {
                uint32_t a, b;
                unpack2(dist2(pack2(X.x, X.y), pack2(Y.x, Y.y), pack2(Z.x, Z.y)), a, b);
                s0 = __uint_as_float(a), s1 = __uint_as_float(b);
                unpack2(dist2(pack2(X.z, X.w), pack2(Y.z, Y.w), pack2(Z.z, Z.w)), a, b);
                s2 = __uint_as_float(a), s3 = __uint_as_float(b);
            }
            // some sum of the block undecided (between the thresholds), or a flagged candidate in the block
            const bool und = __float_as_uint(s0) - bandLo <= bandSpan || __float_as_uint(s1) - bandLo <= bandSpan ||
                             __float_as_uint(s2) - bandLo <= bandSpan || __float_as_uint(s3) - bandLo <= bandSpan;
            const uint32_t fl = uint32_t(flagged >> k) & 15u;
            if (__any_sync(0xffffffffu, und) || fl)
            {
                carefulOne(s0, k, cnt, J.x, node, fl & 1u);
                carefulOne(s1, k + 1, cnt, J.y, node, fl & 2u);
                carefulOne(s2, k + 2, cnt, J.z, node, fl & 4u);
                carefulOne(s3, k + 3, cnt, J.w, node, fl & 8u);
                continue;
            }
            // `out` runs ahead of the stored entries when a list is full (counts are not truncated, H3)
            appendIf<SELF, GUARD>(out, rowEnd, J.x, s0, thrLo, i);
            appendIf<SELF, GUARD>(out, rowEnd, J.y, s1, thrLo, i);
            appendIf<SELF, GUARD>(out, rowEnd, J.z, s2, thrLo, i);
            appendIf<SELF, GUARD>(out, rowEnd, J.w, s3, thrLo, i);
        }
    };

    //! particles [jb, jb + cnt) of leaf `node`; uncertified: a box on the path failed its containment check
    auto scanLeaf = [&](int node, uint32_t jb, uint32_t cnt, bool uncertified)
    {
        const uint32_t je = jb + cnt;
        const bool self   = jb < grp.y && je > grp.x;
        __syncwarp();
        if (lane < 6) { sh.lbox[lane] = lane < 3 ? centers[3 * node + lane] : sizes[3 * node + lane - 3]; }
        for (uint32_t base = jb; base < je; base += CW_STAGE)
        {
            uint32_t staged  = 0;
            uint64_t flagged = 0;
            __syncwarp();
#pragma unroll
            for (int half = 0; half < CW_STAGE / 32; ++half)
            {
                if (base + half * 32 >= je) { break; }
                const uint32_t j = base + half * 32 + lane;
                bool keep = false, flag = false;
                float c0 = 0, c1 = 0, c2 = 0;
                if (j < je)
                {
                    const T px = x[j], py = y[j], pz = z[j];
                    c0 = float(px - sh.org[0]);
                    c1 = float(py - sh.org[1]);
                    c2 = float(pz - sh.org[2]);
                    float D  = fmaxf(fmaxf(fabsf(c0), fabsf(c1)), fmaxf(fabsf(c2), rb.DwT));
                    float bc = fmaf(D * D * 0x1p-27f, BAND_SB, rb.r2bMax);
                    float ex = fmaxf(fmaxf(rb.lox - c0, c0 - rb.hix), 0.0f);
                    float ey = fmaxf(fmaxf(rb.loy - c1, c1 - rb.hiy), 0.0f);
                    float ez = fmaxf(fmaxf(rb.loz - c2, c2 - rb.hiz), 0.0f);
                    keep     = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > bc);
                    if (keep)
                    {
                        const T tol = sh.org[3];
                        flag = !(rabs(px - sh.lbox[0]) <= sh.lbox[3] + tol && rabs(py - sh.lbox[1]) <= sh.lbox[4] + tol &&
                                 rabs(pz - sh.lbox[2]) <= sh.lbox[5] + tol);
                    }
                }
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                if (keep)
                {
                    const uint32_t pos = staged + __popc(km & ltMask);
                    sh.cx[pos]         = c0;
                    sh.cy[pos]         = c1;
                    sh.cz[pos]         = c2;
                    sh.cj[pos]         = j;
                }
                const unsigned fm = __ballot_sync(0xffffffffu, flag);
                if (fm | unsigned(uncertified))
                {
                    // positions of the flagged candidates among the kept ones (rare)
                    for (unsigned m = uncertified ? km : fm; m; m &= m - 1)
                    {
                        const int l = __ffs(int(m)) - 1;
                        flagged |= uint64_t(1) << (staged + __popc(km & ((1u << l) - 1u)));
                    }
                }
                staged += __popc(km);
            }
            if (staged == 0) { continue; }
            // pad to a multiple of four with candidates at infinity (never inside; index checked on the exact route)
            if (lane < 4)
            {
                const float inf      = __int_as_float(0x7f800000);
                sh.cx[staged + lane] = inf;
                sh.cy[staged + lane] = inf;
                sh.cz[staged + lane] = inf;
                sh.cj[staged + lane] = i;
            }
            __syncwarp();
            const bool guard = __any_sync(0xffffffffu, out + 4ull * staged > rowEnd);
            if (self)
            {
                if (guard) { testStaged(staged, node, flagged, std::true_type{}, std::true_type{}); }
                else { testStaged(staged, node, flagged, std::true_type{}, std::false_type{}); }
            }
            else
            {
                if (guard) { testStaged(staged, node, flagged, std::false_type{}, std::true_type{}); }
                else { testStaged(staged, node, flagged, std::false_type{}, std::false_type{}); }
            }
        }
    };

    /*! the 8 children of `parent`, one per lane: child offsets, reach test against the targets' bounding box,
     *  containment of the child boxes in the parent's box, particle ranges of the leaf children */
    auto enterGroup = [&](int parent, int child0, int depth, bool& bad) -> uint32_t
    {
        bool reach = false, out = false;
        T pv = T(0);
        if (lane >= 8 && lane < 14)
        {
            pv = lane < 11 ? centers[3 * parent + (lane - 8)] : sizes[3 * parent + (lane - 11)];
        }
        const T pcx = __shfl_sync(0xffffffffu, pv, 8), pcy = __shfl_sync(0xffffffffu, pv, 9),
                pcz = __shfl_sync(0xffffffffu, pv, 10), psx = __shfl_sync(0xffffffffu, pv, 11),
                psy = __shfl_sync(0xffffffffu, pv, 12), psz = __shfl_sync(0xffffffffu, pv, 13);
        if (lane < 8)
        {
            const int node = child0 + int(lane);
            const int co   = childOffsets[node];
            const T cx = centers[3 * node], cy = centers[3 * node + 1], cz = centers[3 * node + 2];
            const T sx = sizes[3 * node], sy = sizes[3 * node + 1], sz = sizes[3 * node + 2];
            const T tol = sh.org[3];
            float4 gc, gs;
            gc.x = float(cx - sh.org[0]), gc.y = float(cy - sh.org[1]), gc.z = float(cz - sh.org[2]), gc.w = 0.0f;
            gs.x = float(sx), gs.y = float(sy), gs.z = float(sz);
            float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), rb.DwT)),
                            fmaxf(gs.x, fmaxf(gs.y, gs.z)));
            gs.w    = D * D * 0x1p-27f;
            reach   = boxReachable(rb, gc, gs);
            out     = !(rabs(cx - pcx) + sx <= psx + tol && rabs(cy - pcy) + sy <= psy + tol &&
                    rabs(cz - pcz) + sz <= psz + tol);
            uint32_t jb = 0, cnt = 0;
            if (reach && co == 0)
            {
                const int leafIdx = internalToLeaf[node];
                jb                = layout[leafIdx];
                cnt               = layout[leafIdx + 1] - jb;
            }
            sh.co[depth][lane]  = co;
            sh.jb[depth][lane]  = jb;
            sh.cnt[depth][lane] = cnt;
        }
        const uint32_t reachMask = __ballot_sync(0xffffffffu, reach) & 0xffu;
        bad                      = (__ballot_sync(0xffffffffu, out) & reachMask) != 0;
        if (lane == 0) { sh.wm[depth] = uint8_t(reachMask); }
        __syncwarp();
        return reachMask;
    };

    const int rootChild = childOffsets[0];
    if (rootChild == 0) { scanLeaf(0, layout[0], layout[1] - layout[0], false); }
    else
    {
        int depth    = 1;
        int base     = rootChild;
        int uncDepth = 0; // smallest depth at which a containment check failed on the current path (0: none)
        bool bad;
        uint32_t wm = enterGroup(0, rootChild, 1, bad);
        if (bad) { uncDepth = 1; }
        while (true)
        {
            if (wm == 0)
            {
                if (depth == 1) { break; }
                const int up = parents[(base - 1) >> 3];
                --depth;
                if (depth < uncDepth) { uncDepth = 0; }
                base = ((up - 1) & ~7) + 1;
                wm   = uint32_t(sh.wm[depth]) & ~((2u << ((up - 1) & 7)) - 1u);
                continue;
            }
            const int c = __ffs(int(wm)) - 1;
            wm &= wm - 1;
            const int node  = base + c;
            const int child = sh.co[depth][c];
            if (child == 0)
            {
                const uint32_t cnt = sh.cnt[depth][c];
                if (cnt) { scanLeaf(node, sh.jb[depth][c], cnt, uncDepth != 0 || (!Filt && depth > SG_FLOAT_DEPTH)); }
            }
            else
            {
                ++depth;
                wm   = enterGroup(node, child, depth, bad);
                base = child;
                if (bad && uncDepth == 0) { uncDepth = depth; }
            }
        }
    }
    if (valid) { neighborsCount[i - first] = ngmax - uint32_t((long long)(rowEnd - out) >> 2); }
}

template<class T, bool PBC>
__global__ void __launch_bounds__(SG_TARGETS) nbSearchKernel(const T* __restrict__ x,
                                                             const T* __restrict__ y,
                                                             const T* __restrict__ z,
                                                             const T* __restrict__ h,
                                                             uint32_t first,
                                                             const uint2* __restrict__ superGroups,
                                                             const uint32_t* __restrict__ numSuperGroupsPtr,
                                                             Box<T> box,
                                                             const int* __restrict__ childOffsets,
                                                             const int* __restrict__ parents,
                                                             const int* __restrict__ internalToLeaf,
                                                             const uint32_t* __restrict__ layout,
                                                             const T* __restrict__ centers,
                                                             const T* __restrict__ sizes,
                                                             const uint4* __restrict__ entries,
                                                             const int* __restrict__ header,
                                                             uint32_t ngmax,
                                                             uint32_t* __restrict__ neighbors,
                                                             uint32_t* __restrict__ neighborsCount)
{
    constexpr bool Filt = sizeof(T) == 8;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SearchShared& sh = *reinterpret_cast<SearchShared*>(smemRaw);

    const size_t sg = blockIdx.x;
    if (sg >= size_t(*numSuperGroupsPtr)) { return; }
    const unsigned lane   = threadIdx.x & 31;
    const unsigned warp   = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    const uint2 grp       = superGroups[sg];
    const uint32_t numT   = grp.y - grp.x;
    const uint32_t numSub = (numT + 31) / 32;
    // balanced split of the targets over the warps
    const uint32_t subBegin = grp.x + uint32_t(uint64_t(min(warp, numSub)) * numT / numSub);
    const uint32_t subEnd   = grp.x + uint32_t(uint64_t(min(warp + 1, numSub)) * numT / numSub);
    const int numEntries    = header[sg];

    if (numEntries < 0)
    {
        if (warp < numSub)
        {
            warpSearch<T, PBC>(reinterpret_cast<WarpShared*>(smemRaw)[warp], make_uint2(subBegin, subEnd), x, y, z, h,
                               first, box, childOffsets, parents, internalToLeaf, layout, centers, sizes, ngmax,
                               neighbors, neighborsCount);
        }
        return;
    }

    const bool active = warp < numSub;
    const bool valid  = active && subBegin + lane < subEnd;
    const uint32_t i  = valid ? subBegin + lane : (active ? subEnd - 1 : grp.x);

    Target<T> t;
    t.x        = x[i];
    t.y        = y[i];
    t.z        = z[i];
    const T hi = h[i];
    t.radiusSq = T(4.0) * hi * hi;
    t.usePbc   = false; // super-groups with a periodic target never get here

    // single-precision frame of the CTA: relative to the first target of the super-group for double searches
    const T ox = Filt ? x[grp.x] : T(0), oy = Filt ? y[grp.x] : T(0), oz = Filt ? z[grp.x] : T(0);
    const float txf = float(t.x - ox);
    const float tyf = float(t.y - oy);
    const float tzf = float(t.z - oz);
    const float r2f = float(t.radiusSq);

    // bounding box of the warp's targets, largest radius
    ReachBox rb;
    rb.lox = warpMinF(txf), rb.hix = warpMaxF(txf);
    rb.loy = warpMinF(tyf), rb.hiy = warpMaxF(tyf);
    rb.loz = warpMinF(tzf), rb.hiz = warpMaxF(tzf);
    rb.DwT = fmaxf(fmaxf(fmaxf(fabsf(rb.lox), fabsf(rb.hix)), fmaxf(fabsf(rb.loy), fabsf(rb.hiy))),
                   fmaxf(fabsf(rb.loz), fabsf(rb.hiz)));
    const float r2max = warpMaxF(r2f);
    rb.r2bMax         = (r2max > 1e-30f) ? r2max * BAND_KB : __int_as_float(0x7f800000);

    /* ---- thresholds of the batch test (per lane).
     * sure: the candidate is inside AND every box on the path to its leaf passes the continuation test:
     *       d < r (1 - 2^-20) - Delta, Delta = CertTol::delta * magnitude (see the header of this section)
     * T = double: the test runs on the float sum s of squared relative coordinates; s < pairA certifies d2 < r2sure,
     *       s > pairB certifies d2 >= r2 (bounds as in warpSearch); the shell in between takes the exact route.
     * T = float: the reference's own float expression decides inside / outside; inside but not sure: exactMine. */
    // coordinate magnitude of the super-group (the same value nbRangesKernel used for its checks)
    {
        const float mf = warpMaxF(__double2float_ru(coordMagnitude(box, t.x, t.y, t.z)));
        if (lane == 0) { sh.magW[warp] = mf; }
    }
    __syncthreads();
    const double mag   = double(fmaxf(fmaxf(sh.magW[0], sh.magW[1]), fmaxf(sh.magW[2], sh.magW[3])));
    const double rD    = sqrt(double(t.radiusSq));
    const double vSure = rD * (1.0 - 0x1p-20) - CertTol<T>::delta * mag;
    const double r2sure = vSure > 0.0 ? vSure * vSure * (1.0 - 0x1p-19) : -1.0;
    const float Dpair  = 1.01f * (rb.DwT + sqrtf(r2max));
    const float Epair  = Dpair * Dpair * 0x1p-29f;
    float thrLo, thrHi; // accept surely below thrLo; T = double: reject surely above thrHi; T = float: inside below thrHi
    if (Filt)
    {
        const float r2sf = __double2float_rd(r2sure);
        thrLo            = (valid && r2sf > 1e-30f) ? fmaf(-Epair, BAND_SA, r2sf * BAND_KA) : -1.0f;
        thrHi            = (r2f > 1e-30f) ? fmaf(Epair, BAND_SB, r2f * BAND_KB) : __int_as_float(0x7f800000);
        if (!valid) { thrHi = -1.0f; } // padding lanes: everything is surely outside
    }
    else
    {
        thrLo = valid ? __double2float_rd(r2sure) : -1.0f;
        thrHi = valid ? r2f : -1.0f; // d2 < radiusSq, the reference's comparison
    }
    const uint64_t ntx2 = pack2(-txf, -txf), nty2 = pack2(-tyf, -tyf), ntz2 = pack2(-tzf, -tzf);
    const uint64_t nLo2 = pack2(-thrLo, -thrLo);
    const uint64_t hi2  = Filt ? pack2(thrHi, thrHi) : pack2(-thrHi, -thrHi);
    const uint64_t mOne = pack2(-1.0f, -1.0f), zero2 = pack2(0.0f, 0.0f);

    uint32_t* const row = neighbors + size_t(i - first) * size_t(ngmax);
    uint32_t numFound   = 0;
    SearchWarpShared& pv = sh.w[warp];
    uint32_t pcnt        = 0; // buffered candidates of this warp
    if (lane < SG_PCAP / 32) { pv.unc[lane] = 0; }
    __syncwarp();

    /*! tests buffer entries [base, base + 32) against this lane's target; returns the accepted ones, first entry in
     *  bit 31 (`m = (m << 1) | sign` per candidate; the sign bit of (s - bound) is the outcome of s < bound for non-NaN
     *  operands; arithmetic NaNs are the canonical positive NaN: "not smaller", as the comparison says; for T = double
     *  they end up in the shell that takes the exact route) */
    auto evalWord = [&](int base, uint32_t numValid) -> uint32_t
    {
        uint32_t lo = 0, up = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
        {
            const float4 X = *reinterpret_cast<const float4*>(&pv.cx[base + 4 * q]);
            const float4 Y = *reinterpret_cast<const float4*>(&pv.cy[base + 4 * q]);
            const float4 Z = *reinterpret_cast<const float4*>(&pv.cz[base + 4 * q]);
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                const uint64_t x2 = half ? pack2(X.z, X.w) : pack2(X.x, X.y);
                const uint64_t y2 = half ? pack2(Y.z, Y.w) : pack2(Y.x, Y.y);
                const uint64_t z2 = half ? pack2(Z.z, Z.w) : pack2(Z.x, Z.y);
                const uint64_t dx = add2(x2, ntx2), dy = add2(y2, nty2), dz = add2(z2, ntz2);
                uint64_t s;
                if (Filt)
                {
                    s = mul2(dz, dz);
                    s = fma2(dy, dy, s);
                    s = fma2(dx, dx, s);
                }
                else
                {
                    // the reference's float expression (dx*dx + dy*dy) + dz*dz (findneighbors.hpp:33-60).  The products
                    // are formed as fma(d, d, +0) = RN(d*d): separate mul.rn/add.rn.f32x2 pairs get contracted to
                    // FFMA2 by ptxas even under --fmad=false
                    s = add2(add2(fma2(dx, dx, zero2), fma2(dy, dy, zero2)), fma2(dz, dz, zero2));
                }
                uint32_t a0, a1, b0, b1;
                unpack2(add2(s, nLo2), a0, a1); // sign set: s < thrLo
                lo = __funnelshift_l(a0, lo, 1);
                lo = __funnelshift_l(a1, lo, 1);
                // T = double: sign set: s > thrHi (surely outside); T = float: sign set: d2 < radiusSq
                unpack2(Filt ? fma2(s, mOne, hi2) : add2(s, hi2), b0, b1);
                up = __funnelshift_l(b0, up, 1);
                up = __funnelshift_l(b1, up, 1);
            }
        }
        const uint32_t validMask = numValid >= 32 ? 0xffffffffu : ~(0xffffffffu >> numValid);
        const uint32_t uncBits   = __brev(pv.unc[base >> 5]) & validMask;
        uint32_t in  = lo & validMask & ~uncBits;
        uint32_t amb = (Filt ? ~(lo | up) : (up & ~lo)) & validMask; // undecided by the thresholds
        amb |= (Filt ? ~up : up) & uncBits;                          // unchecked particles: never "surely"
        if (!valid) { amb = 0; }
        if (__any_sync(0xffffffffu, amb != 0))
        {
            while (amb)
            {
                const int k        = __clz(int(amb));
                const uint32_t bit = 0x80000000u >> k;
                amb ^= bit;
                bool ok = !Filt || exactInside(x, y, z, pv.cj[base + k], t.x, t.y, t.z, t.radiusSq);
                ok      = ok && exactMine(t, centers, sizes, parents, pv.cnode[base + k], box);
                if (ok) { in |= bit; }
            }
        }
        return in;
    };

    //! evaluate and write out the first min(pcnt, 128) buffered candidates, keep the rest
    auto flush = [&]()
    {
        __syncwarp();
        const uint32_t n = min(pcnt, uint32_t(SG_BATCH));
#pragma unroll 1
        for (uint32_t w = 0; w < SG_BATCH / 32; ++w)
        {
            pv.res[w][lane] = w * 32 < n ? evalWord(int(w * 32), min(n - w * 32, 32u)) : 0u;
        }
        // accepted candidates in buffer order = ascending particle index; lanes run through their words independently
        {
            uint32_t w = 0, m = pv.res[0][lane];
            while (true)
            {
                if (m == 0)
                {
                    if (++w == SG_BATCH / 32) { break; }
                    m = pv.res[w][lane];
                    continue;
                }
                const int p = 31 - __clz(int(m));
                m ^= 1u << p;
                const uint32_t j = pv.cj[w * 32 + 31 - p];
                if (j != i)
                {
                    if (numFound < ngmax) { row[numFound] = j; }
                    ++numFound;
                }
            }
        }
        __syncwarp();
        if (pcnt > SG_BATCH)
        {
            const uint32_t rest = pcnt - SG_BATCH; // <= 31
            float a = 0, b = 0, c = 0;
            uint32_t d = 0;
            int e      = 0;
            if (lane < rest)
            {
                a = pv.cx[SG_BATCH + lane];
                b = pv.cy[SG_BATCH + lane];
                c = pv.cz[SG_BATCH + lane];
                d = pv.cj[SG_BATCH + lane];
                e = pv.cnode[SG_BATCH + lane];
            }
            const uint32_t u = pv.unc[SG_BATCH / 32];
            __syncwarp();
            if (lane < rest)
            {
                pv.cx[lane]    = a;
                pv.cy[lane]    = b;
                pv.cz[lane]    = c;
                pv.cj[lane]    = d;
                pv.cnode[lane] = e;
            }
            if (lane < SG_PCAP / 32) { pv.unc[lane] = lane == 0 ? u : 0u; }
            pcnt = rest;
        }
        else
        {
            if (lane < SG_PCAP / 32) { pv.unc[lane] = 0; }
            pcnt = 0;
        }
        __syncwarp();
    };

    const uint4* const list = entries + sg * size_t(SG_CAP);
    for (int e0 = 0; e0 < numEntries;)
    {
        /* ---- ranges of this round: as many list entries as fit the chunk (at least 16), with their leaf boxes */
        if (warp == 0)
        {
            const int e = e0 + int(lane);
            uint4 ent   = make_uint4(0, 0, 0, 0);
            if (e < numEntries) { ent = list[e]; }
            const uint32_t cnt = ent.y & ~SG_UNCERT;
            uint32_t incl      = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= unsigned(o)) { incl += v; }
            }
            const bool fits = e < numEntries && incl <= uint32_t(SG_CHUNK);
            const int take  = __popc(__ballot_sync(0xffffffffu, fits)); // fits is a prefix property
            if (int(lane) < take)
            {
                sh.rJb[lane]   = ent.x;
                sh.rCnt[lane]  = ent.y;
                sh.rNode[lane] = int(ent.z);
                sh.rOff[lane]  = incl - cnt;
                float4 gc, gs;
                // boxes are compared with the warps' bounding boxes: the error term uses the CTA-wide magnitude bound
                relBox(centers, sizes, int(ent.z), ox, oy, oz, 0.0f, gc, gs);
                sh.rBoxC[lane] = gc;
                sh.rBoxS[lane] = gs;
            }
            if (lane == 0) { sh.numRanges = take; }
        }
        __syncthreads();
        const int nr = sh.numRanges;

        /* ---- stage the particles of the ranges: one load, containment check and conversion per particle and CTA */
        for (int r = int(warp); r < nr; r += SG_WARPS)
        {
            const uint32_t jb  = sh.rJb[r];
            const uint32_t cw  = sh.rCnt[r];
            const uint32_t cnt = cw & ~SG_UNCERT;
            const uint32_t off = sh.rOff[r];
            const int node     = sh.rNode[r];
            const T bcx = centers[3 * node], bcy = centers[3 * node + 1], bcz = centers[3 * node + 2];
            const T bsx = sizes[3 * node], bsy = sizes[3 * node + 1], bsz = sizes[3 * node + 2];
            const T tol = T(CertTol<T>::check * mag);
            for (uint32_t k = lane; k < cnt; k += 32)
            {
                const uint32_t j = jb + k;
                const T px = x[j], py = y[j], pz = z[j];
                const bool inBox = rabs(px - bcx) <= bsx + tol && rabs(py - bcy) <= bsy + tol && rabs(pz - bcz) <= bsz + tol;
                float4 c;
                c.x = float(px - ox);
                c.y = float(py - oy);
                c.z = float(pz - oz);
                c.w = __uint_as_float(j | ((cw & SG_UNCERT) || !inBox ? SG_UNCERT : 0u));
                sh.chunk[off + k] = c;
            }
        }
        __syncthreads();

        /* ---- every warp: leaves within reach of ITS targets, their particles culled against its bounding box */
        if (active)
        {
            bool reach = false;
            if (int(lane) < nr)
            {
                float4 gc = sh.rBoxC[lane], gs = sh.rBoxS[lane];
                // error term with this warp's magnitude
                float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), rb.DwT)),
                                fmaxf(gs.x, fmaxf(gs.y, gs.z)));
                gs.w    = D * D * 0x1p-27f;
                reach   = boxReachable(rb, gc, gs) || (sh.rCnt[lane] & SG_UNCERT);
            }
            uint32_t rm = __ballot_sync(0xffffffffu, reach);
            while (rm)
            {
                const int r = __ffs(int(rm)) - 1;
                rm &= rm - 1;
                const uint32_t cnt = sh.rCnt[r] & ~SG_UNCERT;
                const uint32_t off = sh.rOff[r];
                const int node     = sh.rNode[r];
                for (uint32_t k0 = 0; k0 < cnt; k0 += 32)
                {
                    const uint32_t k = k0 + lane;
                    bool keep        = false;
                    float4 c         = make_float4(0, 0, 0, 0);
                    if (k < cnt)
                    {
                        c        = sh.chunk[off + k];
                        float D  = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), rb.DwT));
                        float bc = fmaf(D * D * 0x1p-27f, BAND_SB, rb.r2bMax);
                        float ex = fmaxf(fmaxf(rb.lox - c.x, c.x - rb.hix), 0.0f);
                        float ey = fmaxf(fmaxf(rb.loy - c.y, c.y - rb.hiy), 0.0f);
                        float ez = fmaxf(fmaxf(rb.loz - c.z, c.z - rb.hiz), 0.0f);
                        keep     = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > bc);
                    }
                    const unsigned km = __ballot_sync(0xffffffffu, keep);
                    if (km == 0) { continue; }
                    const uint32_t jw = __float_as_uint(c.w);
                    if (keep)
                    {
                        const uint32_t pos = pcnt + __popc(km & ltMask);
                        pv.cx[pos]         = c.x;
                        pv.cy[pos]         = c.y;
                        pv.cz[pos]         = c.z;
                        pv.cj[pos]         = jw & ~SG_UNCERT;
                        pv.cnode[pos]      = node;
                        if (jw & SG_UNCERT) { atomicOr(&pv.unc[pos >> 5], 1u << (pos & 31)); }
                    }
                    pcnt += __popc(km);
                    if (pcnt >= uint32_t(SG_BATCH)) { flush(); }
                }
            }
        }
        e0 += nr;
        __syncthreads(); // the chunk and the range table are free again
    }
    if (active && pcnt) { flush(); }
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
    Box<T> box         = makeBox<T>(lim, bnd);
    const bool pbc     = box.pbc(0) || box.pbc(1) || box.pbc(2);
    const int variant  = tuning(TUNE_NB_KERNEL);
    const int numNodes = numLeaves + (numLeaves - 1) / 7;

    // target groups: counts per leaf -> exclusive scan -> fill.  Upper bound on the number of groups is known on the
    // host; the per-warp search leaves the exact number on the device (no synchronisation)
    const size_t limit = variant == 2 ? SG_TARGETS : 32;
    size_t maxGroups   = size_t(numLeaves) + (size_t(last) - first) / limit + 1;
    CSB_SCRATCH(groupOffsets, uint32_t*, s, SCRATCH_A, (size_t(numLeaves) + 1) * sizeof(uint32_t));
    CSB_SCRATCH(groups, uint2*, s, SCRATCH_B, maxGroups * sizeof(uint2));
    CSB_SCRATCH(scanTmp, void*, s, SCRATCH_C, scanTempBytes(size_t(numLeaves) + 1));
    CSB_CHECK(cudaMemsetAsync(groupOffsets, 0, (size_t(numLeaves) + 1) * sizeof(uint32_t), s));

    auto buildGroups = [&](auto countKernel, auto fillKernel) -> int
    {
        countKernel<<<iceil(numNodes, 256), 256, 0, s>>>(childOffsets, internalToLeaf, layout, numNodes, first, last,
                                                         groupOffsets, nullptr, nullptr);
        CSB_LAUNCH_CHECK();
        if (int e = exclusiveScanU32(groupOffsets, groupOffsets, size_t(numLeaves) + 1, scanTmp, s)) { return e; }
        fillKernel<<<iceil(numNodes, 256), 256, 0, s>>>(childOffsets, internalToLeaf, layout, numNodes, first, last,
                                                        nullptr, groupOffsets, groups);
        CSB_LAUNCH_CHECK();
        return 0;
    };

    if (variant == 0 || variant == 1)
    {
        const bool greedy = tuning(TUNE_NB_GROUPS) == 0; // knob value 1 selects the leaf-aligned groups
        if (int e = greedy ? buildGroups(groupBuildKernel<false, 1, 32>, groupBuildKernel<true, 1, 32>)
                           : buildGroups(groupBuildKernel<false, 0, 32>, groupBuildKernel<true, 0, 32>))
        {
            return e;
        }
        unsigned grid = iceil(maxGroups * 32, NB_THREADS);
        auto kernel   = variant == 0 ? (pbc ? findNeighborsKernel<T, true> : findNeighborsKernel<T, false>)
                                     : (pbc ? findNeighborsCertKernel<T, true> : findNeighborsCertKernel<T, false>);
        kernel<<<grid, NB_THREADS, 0, s>>>(x, y, z, h, first, groups, groupOffsets + numLeaves, box, childOffsets,
                                           parents, internalToLeaf, layout, centers, sizes, ngmax, neighbors,
                                           neighborsCount);
        CSB_LAUNCH_CHECK();
        return 0;
    }

    // cooperative search: super-groups, range lists, search.  The number of super-groups sizes the range lists, so it
    // is read back (one 4-byte copy)
    if (int e = buildGroups(groupBuildKernel<false, 1, SG_TARGETS>, groupBuildKernel<true, 1, SG_TARGETS>)) { return e; }
    uint32_t numSuperGroups = 0;
    CSB_CHECK(cudaMemcpyAsync(&numSuperGroups, groupOffsets + numLeaves, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CSB_CHECK(cudaStreamSynchronize(s));
    if (numSuperGroups == 0) { return 0; }
    CSB_SCRATCH(entries, uint4*, s, SCRATCH_D, size_t(numSuperGroups) * SG_CAP * sizeof(uint4));
    CSB_SCRATCH(header, int*, s, SCRATCH_E, size_t(numSuperGroups) * sizeof(int));
    {
        auto kernel = pbc ? nbRangesKernel<T, true> : nbRangesKernel<T, false>;
        kernel<<<iceil(size_t(numSuperGroups) * 32, SG_TARGETS), SG_TARGETS, 0, s>>>(
            x, y, z, h, groups, groupOffsets + numLeaves, box, childOffsets, parents, internalToLeaf, layout, centers,
            sizes, entries, header);
        CSB_LAUNCH_CHECK();
    }
    {
        auto kernel = pbc ? nbSearchKernel<T, true> : nbSearchKernel<T, false>;
        CSB_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(SearchShared))));
        kernel<<<numSuperGroups, SG_TARGETS, sizeof(SearchShared), s>>>(
            x, y, z, h, first, groups, groupOffsets + numLeaves, box, childOffsets, parents, internalToLeaf, layout,
            centers, sizes, entries, header, ngmax, neighbors, neighborsCount);
        CSB_LAUNCH_CHECK();
    }
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
