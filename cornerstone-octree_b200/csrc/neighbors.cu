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
 *  - Trees whose leaves hold few particles (the occupancy of a uniform tree jumps by 8x whenever the particle count
 *    crosses a power of 8 times the bucket size) are searched by a second organisation, the group-steered search
 *    further down: the walk is steered by the bounding box of the warp's targets alone, contiguous particle ranges are
 *    staged whatever the leaf boundaries, and the reference's per-target pruning is reproduced by a certified
 *    distance shell plus the reference's own box tests on the rare candidates inside the shell.
 *  A CTA-cooperative variant with shared staging and a bit-mask batch variant were also built and measured in round 2;
 *  both returned identical lists and both were slower, see profiles/r2_notes.md.
 */
#include <algorithm>
#include <cmath>
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
 *  POLICY 0 (tuning knob TUNE_NB_GROUPS = 1): consecutive sibling leaves are packed greedily into one group while they hold at most 32 targets together
 *  (deep trees have leaves with a handful of particles; a warp per such leaf would run mostly empty), a leaf with more
 *  than 32 targets is cut into ceil(count/32) balanced groups: groups never straddle a leaf boundary unless they hold
 *  whole leaves.
 *  POLICY 1 (default): every maximal run of consecutive leaf siblings (their particles are contiguous) is cut into
 *  ceil(total/32) balanced groups regardless of the leaf boundaries inside the run: groups are full (a uniform tree with
 *  32 particles per leaf gives 8 groups of 32 per parent instead of ~12 of 21), at the price of a slightly larger
 *  bounding box where a group takes particles from two curve-adjacent leaves.  Measured at 64 Mi uniform particles:
 *  58.1 -> 53.6 ms (profiles/r2_notes.md); this is the default.
 *  FILL = false counts the groups that start at each leaf, FILL = true writes them at the scanned offsets. */
template<bool FILL, int POLICY>
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
        uint32_t ng = (c + 31) / 32;
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
    T cellRadiusSq; // radiusSq * searchExtFactor^2: the radius of the continuation test (findneighbors.hpp:100)
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
    return n2 < t.cellRadiusSq;
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
__device__ __noinline__ T exactDistSq(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                      uint32_t j, T tx, T ty, T tz)
{
    T ex = x[j] - tx;
    T ey = y[j] - ty;
    T ez = z[j] - tz;
    return ex * ex + ey * ey + ez * ez;
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
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

/*! if (s < thr [&& j != self] [&& out < end]) *out = j;  if (s < thr [&& j != self]) out += 4
 *  (out = {outLo, outHi}, end: global-space byte addresses).  Predicated, no branches; the address is kept as two
 *  32-bit halves so that the increment is a predicated add-with-carry pair: 4 to 7 SASS instructions per candidate. */
template<bool SELF, bool GUARD>
__device__ __forceinline__ void appendIf(uint32_t& outLo, uint32_t& outHi, unsigned long long end, uint32_t j, float s,
                                         float thr, uint32_t self)
{
    if (SELF && GUARD)
    {
        asm volatile("{\n .reg .pred q, g;\n .reg .b64 a;\n mov.b64 a, {%0, %1};\n setp.lt.f32 q, %3, %4;\n"
                     " setp.ne.and.u32 q, %2, %5, q;\n setp.lt.and.u64 g, a, %6, q;\n @g st.global.u32 [a], %2;\n"
                     " @q add.cc.u32 %0, %0, 4;\n @q addc.u32 %1, %1, 0;\n}"
                     : "+r"(outLo), "+r"(outHi)
                     : "r"(j), "f"(s), "f"(thr), "r"(self), "l"(end)
                     : "memory");
    }
    else if (SELF)
    {
        asm volatile("{\n .reg .pred q;\n .reg .b64 a;\n mov.b64 a, {%0, %1};\n setp.lt.f32 q, %3, %4;\n"
                     " setp.ne.and.u32 q, %2, %5, q;\n @q st.global.u32 [a], %2;\n"
                     " @q add.cc.u32 %0, %0, 4;\n @q addc.u32 %1, %1, 0;\n}"
                     : "+r"(outLo), "+r"(outHi)
                     : "r"(j), "f"(s), "f"(thr), "r"(self)
                     : "memory");
    }
    else if (GUARD)
    {
        asm volatile("{\n .reg .pred q, g;\n .reg .b64 a;\n mov.b64 a, {%0, %1};\n setp.lt.f32 q, %3, %4;\n"
                     " setp.lt.and.u64 g, a, %5, q;\n @g st.global.u32 [a], %2;\n"
                     " @q add.cc.u32 %0, %0, 4;\n @q addc.u32 %1, %1, 0;\n}"
                     : "+r"(outLo), "+r"(outHi)
                     : "r"(j), "f"(s), "f"(thr), "l"(end)
                     : "memory");
    }
    else
    {
        asm volatile("{\n .reg .pred q;\n .reg .b64 a;\n mov.b64 a, {%0, %1};\n setp.lt.f32 q, %3, %4;\n"
                     " @q st.global.u32 [a], %2;\n @q add.cc.u32 %0, %0, 4;\n @q addc.u32 %1, %1, 0;\n}"
                     : "+r"(outLo), "+r"(outHi)
                     : "r"(j), "f"(s), "f"(thr)
                     : "memory");
    }
}

//! the reference's acceptance test with the periodic fold of findneighbors.hpp:33-48, applied if the target's search
//! sphere leaves the box (usePbc, :104-106)
template<class T>
__device__ __noinline__ T exactDistSqPbc(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                         uint32_t j, T tx, T ty, T tz, bool usePbc, const Box<T>& box)
{
    T dx = x[j] - tx;
    T dy = y[j] - ty;
    T dz = z[j] - tz;
    if (usePbc)
    {
        dx = pbcFold(dx, 0, box);
        dy = pbcFold(dy, 1, box);
        dz = pbcFold(dz, 2, box);
    }
    return dx * dx + dy * dy + dz * dz;
}

template<class T>
__device__ __forceinline__ bool exactInside(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                            uint32_t j, T tx, T ty, T tz, T radiusSq)
{
    return exactDistSq(x, y, z, j, tx, ty, tz) < radiusSq;
}
template<class T>
__device__ __forceinline__ bool exactInsidePbc(const T* __restrict__ x, const T* __restrict__ y,
                                               const T* __restrict__ z, uint32_t j, T tx, T ty, T tz, T radiusSq,
                                               bool usePbc, const Box<T>& box)
{
    return exactDistSqPbc(x, y, z, j, tx, ty, tz, usePbc, box) < radiusSq;
}

constexpr int NB_MAX_DEPTH = 23; // >= maxTreeLevel<uint64_t> + 2
constexpr int NB_STAGE     = 64; // staged candidates per round (two half-rounds of 32 loads)

struct alignas(16) LaneWalkShared
{
    // staged candidates: x, y, z (relative floats for T = double, the values themselves for T = float) and particle
    // index, padded to a multiple of four entries
    float cx[NB_STAGE + 4], cy[NB_STAGE + 4], cz[NB_STAGE + 4];
    uint32_t cj[NB_STAGE + 4];
    float4 geoC[8], geoS[8];        // centres / sizes of the 8 children being tested; geoS.w = error term E of the child
    uint8_t mask[NB_MAX_DEPTH][32]; // per tree depth, per lane: which of the 8 siblings this lane's own walk enters
};

/* ================================================================ search with per-lane walks (the default)
 * Every lane keeps the exact pruning state of the reference's own walk (a bit per sibling and tree depth); leaves are
 * staged and tested one at a time for the lanes whose walk enters them. */

/*! The search of ONE warp for the targets [grp.x, grp.y) (at most 32).
 *  PBC = false: the box has no periodic dimension, the fold code is not even compiled in.  PBC = true: whether the
 *  fold is needed is decided per warp (any lane whose search sphere leaves the box); such warps run the reference
 *  expressions directly on broadcast loads, lanes that do not need the fold select the unfolded difference exactly as
 *  the reference picks per particle (findneighbors.hpp:104-106,150-151).  Interior warps take the staged path. */
template<class T, bool PBC, bool FOLD, class Th, bool EXT>
__device__ __forceinline__ void warpSearchLanes(LaneWalkShared& sh, const uint2 grp, const T* __restrict__ x, const T* __restrict__ y,
                           const T* __restrict__ z, const Th* __restrict__ h, uint32_t first, const Box<T>& box,
                           const int* __restrict__ childOffsets, const int* __restrict__ parents,
                           const int* __restrict__ internalToLeaf, const uint32_t* __restrict__ layout,
                           const T* __restrict__ centers, const T* __restrict__ sizes, uint32_t ngmax,
                           uint32_t* __restrict__ neighbors, uint32_t* __restrict__ neighborsCount,
                           float searchExt)
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
    // Th may be float with double coordinates: the radius is formed in Th and promoted (findneighbors.hpp:89-99)
    const Th hi = h[i];
    {
        const Th radiusSq = Th(4.0) * hi * hi;
        t.radiusSq        = T(radiusSq);
        t.cellRadiusSq    = EXT ? T(radiusSq * searchExt * searchExt) : t.radiusSq; // findneighbors.hpp:99-100
    }
    {
        bool anyPbc = box.pbc(0) || box.pbc(1) || box.pbc(2);
        T s         = T(2) * T(hi);
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
    // the same for the radius of the continuation tests (differs from the search radius only if searchExtFactor != 1)
    const float c2f    = EXT ? float(t.cellRadiusSq) : r2f;
    const float c2max  = EXT ? warpMaxF(c2f) : r2max;
    const float c2bMax = EXT ? ((c2max > 1e-30f) ? c2max * BAND_KB : __int_as_float(0x7f800000)) : r2bMax;
    float c2a = r2a, c2b = r2b;
    if (EXT)
    {
        c2a = c2f * BAND_KA, c2b = c2f * BAND_KB;
        if (!(c2f > 1e-30f))
        {
            c2a = -1.0f;
            c2b = __int_as_float(0x7f800000);
        }
    }
    // every staged (un-culled) candidate has |coordinate| <= 1.01 (DwT + sqrt(r2max)), see the cull test
    /* Warps that touch a periodic boundary: if the group and its search spheres are small against the box (they always
     * are unless the box holds only a few leaves), every lane sees the same periodic image of a nearby particle or node,
     * so the fold is applied ONCE while staging, to the coordinates relative to the group origin, and the certified
     * single-precision path runs unchanged; the shell it cannot decide (and nodes that are large against the box) take
     * the reference's own folded expressions per lane.  Otherwise, and for float searches (whose staged operands must be
     * the reference's), the whole warp evaluates the reference expressions directly (slowPbc). */
    bool foldOk = Filt && FOLD;
    if (PBC && Filt && FOLD)
    {
        const float reach = 1.01f * sqrtf(fmaxf(r2max, c2max));
        const float ext[3] = {hix - lox, hiy - loy, hiz - loz};
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (box.pbc(d) && !(ext[d] + reach < 0.124f * float(box.len[d]))) { foldOk = false; }
    }
    const bool slowPbc = warpPbc && !foldOk;
    const bool foldPbc = warpPbc && foldOk;

    const float Dpair = 1.01f * (DwT + sqrtf(r2max));
    const float Epair = Dpair * Dpair * 0x1p-29f;
    const float pairA = fmaf(-Epair, BAND_SA, r2a);
    const float pairB = fmaf(Epair, BAND_SB, r2b);

    // out = address of the next list entry; entries at and beyond rowEnd are counted but not stored
    // (findneighbors.hpp:139-146).  Global-space byte addresses: the stores of the staged path are inline PTX.
    unsigned long long out          = __cvta_generic_to_global(neighbors + size_t(i - first) * size_t(ngmax));
    const unsigned long long rowEnd = out + 4ull * ngmax;

    auto append = [&](uint32_t j)
    {
        if (out < rowEnd) { asm volatile("st.global.u32 [%0], %1;" ::"l"(out), "r"(j) : "memory"); }
        out += 4;
    };

    const uint64_t ntx2 = pack2(-txf, -txf), nty2 = pack2(-tyf, -tyf), ntz2 = pack2(-tzf, -tzf);
    const uint64_t zero2 = pack2(0.0f, 0.0f);

    /*! tests the `cnt` staged candidates four at a time (cnt is padded to a multiple of four with candidates at
     *  infinity).  Accept surely below thrLo; T = double: sums in [thrLo, thrHi] (bit patterns bandLo .. bandLo +
     *  bandSpan) take the careful route, where the reference's own double expression decides.
     *  SELF: the leaf holds targets of this group, so a candidate can be the target itself (excluded by index,
     *  findneighbors.hpp:131).  GUARD: a list may reach ngmax during this call. */
    auto testStaged = [&](uint32_t cnt, float thrLo, float thrHi, uint32_t bandLo, uint32_t bandSpan, auto selfTag,
                          auto guardTag)
    {
        constexpr bool SELF  = decltype(selfTag)::value;
        constexpr bool GUARD = decltype(guardTag)::value;
        for (uint32_t k = 0; k < cnt; k += 4)
        {
            const float4 X = *reinterpret_cast<const float4*>(&sh.cx[k]);
            const float4 Y = *reinterpret_cast<const float4*>(&sh.cy[k]);
            const float4 Z = *reinterpret_cast<const float4*>(&sh.cz[k]);
            const uint4 J  = *reinterpret_cast<const uint4*>(&sh.cj[k]);
            float s[4];
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                const uint64_t x2 = half ? pack2(X.z, X.w) : pack2(X.x, X.y);
                const uint64_t y2 = half ? pack2(Y.z, Y.w) : pack2(Y.x, Y.y);
                const uint64_t z2 = half ? pack2(Z.z, Z.w) : pack2(Z.x, Z.y);
                const uint64_t dx = add2(x2, ntx2), dy = add2(y2, nty2), dz = add2(z2, ntz2);
                uint64_t s2;
                if (Filt) { s2 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))); }
                else
                {
                    // the reference's float expression (dx*dx + dy*dy) + dz*dz (findneighbors.hpp:33-60), bit for bit:
                    // the packed operations round per component like the scalar ones.  The products are formed as
                    // fma(d, d, +0) = RN(d*d): separate mul.rn / add.rn.f32x2 get contracted to FFMA2 by ptxas even
                    // under --fmad=false
                    s2 = add2(add2(fma2(dx, dx, zero2), fma2(dy, dy, zero2)), fma2(dz, dz, zero2));
                }
                unpack2(s2, s[2 * half], s[2 * half + 1]);
            }
            if (Filt)
            {
                // a sum in the uncertainty band (rare, < 1 % of the neighbours; never for lanes that do not own the
                // leaf, their band is empty): the reference's own expression decides.  The branch is warp-uniform.
                const bool amb = __float_as_uint(s[0]) - bandLo <= bandSpan || __float_as_uint(s[1]) - bandLo <= bandSpan ||
                                 __float_as_uint(s[2]) - bandLo <= bandSpan || __float_as_uint(s[3]) - bandLo <= bandSpan;
                if (__any_sync(0xffffffffu, amb))
                {
                    const uint32_t jj[4] = {J.x, J.y, J.z, J.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        bool in = s[c] < thrLo;
                        if (!in && !(s[c] > thrHi) && k + c < cnt)
                        {
                            in = foldPbc ? exactInsidePbc(x, y, z, jj[c], t.x, t.y, t.z, t.radiusSq, t.usePbc, box)
                                         : exactInside(x, y, z, jj[c], t.x, t.y, t.z, t.radiusSq);
                        }
                        if (in && jj[c] != i) { append(jj[c]); }
                    }
                    continue;
                }
            }
            uint32_t outLo = uint32_t(out), outHi = uint32_t(out >> 32);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.x, s[0], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.y, s[1], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.z, s[2], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.w, s[3], thrLo, i);
            out = (unsigned long long)(outHi) << 32 | outLo;
        }
    };

    auto scanLeaf = [&](int node, bool mine)
    {
        int leafIdx = internalToLeaf[node];
        uint32_t jb = layout[leafIdx];
        uint32_t je = layout[leafIdx + 1];
        if (slowPbc)
        {
            // the reference expressions on broadcast loads
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
        // not mine: never inside and always surely outside (sums of squares are >= 0).  T = float: the staged values
        // are the reference's operands and thrLo = radiusSq is the reference's comparison, there is no band
        const float thrLo = mine ? (Filt ? pairA : r2f) : -1.0f;
        const float thrHi = mine && Filt ? pairB : -1.0f;
        // bit patterns of the band [thrLo, thrHi] among the non-negative floats (they order like their patterns; NaN
        // sums lie above +inf and never match: NaN < r2 is false for the reference as well)
        uint32_t bandLo = 0xffffffffu, bandSpan = 0;
        if (Filt && thrHi >= 0.0f)
        {
            const uint32_t lo = __float_as_uint(fmaxf(thrLo, 0.0f)), hi = __float_as_uint(thrHi);
            if (hi >= lo)
            {
                bandLo   = lo;
                bandSpan = hi - lo;
            }
        }
        const bool self = jb < grp.y && je > grp.x;
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
                float c0 = 0, c1 = 0, c2 = 0;
                if (j < je)
                {
                    T rx = x[j] - ox, ry = y[j] - oy, rz = z[j] - oz;
                    if (foldPbc)
                    {
                        rx = pbcFold(rx, 0, box);
                        ry = pbcFold(ry, 1, box);
                        rz = pbcFold(rz, 2, box);
                    }
                    c0 = float(rx);
                    c1 = float(ry);
                    c2 = float(rz);
                    float D  = fmaxf(fmaxf(fabsf(c0), fabsf(c1)), fmaxf(fabsf(c2), DwT));
                    float bc = fmaf(D * D * 0x1p-27f, BAND_SB, r2bMax);
                    float ex = fmaxf(fmaxf(lox - c0, c0 - hix), 0.0f);
                    float ey = fmaxf(fmaxf(loy - c1, c1 - hiy), 0.0f);
                    float ez = fmaxf(fmaxf(loz - c2, c2 - hiz), 0.0f);
                    keep     = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > bc);
                }
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                if (keep)
                {
                    const uint32_t pos = cnt + __popc(km & ltMask);
                    sh.cx[pos]         = c0;
                    sh.cy[pos]         = c1;
                    sh.cz[pos]         = c2;
                    sh.cj[pos]         = j;
                }
                cnt += __popc(km);
            }
            if (cnt == 0) { continue; }
            // pad to a multiple of four with candidates at infinity: s = +inf is never below thrLo, and on the careful
            // route the entry index is checked
            if (lane < 4)
            {
                const float inf   = __int_as_float(0x7f800000);
                sh.cx[cnt + lane] = inf;
                sh.cy[cnt + lane] = inf;
                sh.cz[cnt + lane] = inf;
                sh.cj[cnt + lane] = i;
            }
            __syncwarp();
            const bool guard = __any_sync(0xffffffffu, out + 4ull * cnt > rowEnd);
            if (self)
            {
                if (guard) { testStaged(cnt, thrLo, thrHi, bandLo, bandSpan, std::true_type{}, std::true_type{}); }
                else { testStaged(cnt, thrLo, thrHi, bandLo, bandSpan, std::true_type{}, std::false_type{}); }
            }
            else
            {
                if (guard) { testStaged(cnt, thrLo, thrHi, bandLo, bandSpan, std::false_type{}, std::true_type{}); }
                else { testStaged(cnt, thrLo, thrHi, bandLo, bandSpan, std::false_type{}, std::false_type{}); }
            }
        }
    };

    //! this lane's continuation decisions for the 8 children of an internal node its own walk has entered
    auto testChildren = [&](int child0, bool mine) -> uint32_t
    {
        uint32_t bits = 0;
        if (slowPbc)
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
            T rx = centers[3 * node] - ox, ry = centers[3 * node + 1] - oy, rz = centers[3 * node + 2] - oz;
            gs.x = float(sizes[3 * node]);
            gs.y = float(sizes[3 * node + 1]);
            gs.z = float(sizes[3 * node + 2]);
            bool large = false; // a node that is large against the periodic box: the lanes may see different images
            if (foldPbc)
            {
                rx    = pbcFold(rx, 0, box);
                ry    = pbcFold(ry, 1, box);
                rz    = pbcFold(rz, 2, box);
                large = (box.pbc(0) && !(gs.x < 0.124f * float(box.len[0]))) ||
                        (box.pbc(1) && !(gs.y < 0.124f * float(box.len[1]))) ||
                        (box.pbc(2) && !(gs.z < 0.124f * float(box.len[2])));
            }
            gc.x = float(rx);
            gc.y = float(ry);
            gc.z = float(rz);
            gc.w = 0.0f;
            float D = fmaxf(fmaxf(fmaxf(fabsf(gc.x), fabsf(gc.y)), fmaxf(fabsf(gc.z), DwT)),
                            fmaxf(gs.x, fmaxf(gs.y, gs.z)));
            // an infinite error term sends every lane to the reference's own (folded) test and keeps the node reachable
            gs.w          = large ? __int_as_float(0x7f800000) : D * D * 0x1p-27f;
            sh.geoC[lane] = gc;
            sh.geoS[lane] = gs;
            // warp-level cull: a child whose box is certainly farther from the targets' bounding box than the
            // largest radius fails the continuation test of every lane
            float ex = fmaxf(fmaxf(lox - (gc.x + gs.x), (gc.x - gs.x) - hix), 0.0f);
            float ey = fmaxf(fmaxf(loy - (gc.y + gs.y), (gc.y - gs.y) - hiy), 0.0f);
            float ez = fmaxf(fmaxf(loz - (gc.z + gs.z), (gc.z - gs.z) - hiz), 0.0f);
            reach    = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(gs.w, BAND_SB, c2bMax));
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
                    pass     = s2 < fmaf(-gs.w, BAND_SA, c2a);
                    if (!pass && !(s2 > fmaf(gs.w, BAND_SB, c2b)))
                    {
                        pass = cellOverlap<PBC>(t, centers, sizes, child0 + c, box);
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
                    pass = dx * dx + (dy * dy + dz * dz) < t.cellRadiusSq;
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

    if (valid) { neighborsCount[i - first] = ngmax - uint32_t((long long)(rowEnd - out) >> 2); }
}

/*! DEFER (periodic boxes): groups with a target whose search sphere crosses a periodic boundary are not searched here
 *  but appended to `deferred` ([0] = count, then group numbers); this launch then runs the code of the open-box search
 *  for all interior groups, and a second launch (groupList = deferred) handles the few boundary groups with the
 *  periodic code, whose register footprint would otherwise slow down every warp. */
template<class T, bool PBC, bool FOLD, bool DEFER, class Th, bool PERSIST, bool EXT>
__global__ void __launch_bounds__(NB_THREADS) findNeighborsKernel(const T* __restrict__ x,
                                                                  const T* __restrict__ y,
                                                                  const T* __restrict__ z,
                                                                  const Th* __restrict__ h,
                                                                  uint32_t first,
                                                                  const uint2* __restrict__ groups,
                                                                  const uint32_t* __restrict__ numGroupsPtr,
                                                                  const uint32_t* __restrict__ groupList,
                                                                  uint32_t* __restrict__ deferred,
                                                                  Box<T> box,
                                                                  const int* __restrict__ childOffsets,
                                                                  const int* __restrict__ parents,
                                                                  const int* __restrict__ internalToLeaf,
                                                                  const uint32_t* __restrict__ layout,
                                                                  const T* __restrict__ centers,
                                                                  const T* __restrict__ sizes,
                                                                  uint32_t ngmax,
                                                                  uint32_t* __restrict__ neighbors,
                                                                  uint32_t* __restrict__ neighborsCount,
                                                                  const uint32_t* __restrict__ numBad,
                                                                  uint32_t badLimit,
                                                                  uint32_t* __restrict__ workCounter,
                                                                  float searchExt)
{
    // the group-steered search runs instead (it was launched before this kernel with the same criterion)
    if (numBad != nullptr && *numBad <= badLimit) { return; }
    __shared__ LaneWalkShared shAll[NB_THREADS / 32];
    /* PERSIST: groups are handed out through a device counter to one wave of blocks, so the grid does not depend on the
     * (upper bound of the) number of groups and a kernel that has nothing to do costs next to nothing; otherwise one
     * group per warp of the grid.  Which is faster depends on the variant (register allocation), see findNeighbors */
    const unsigned lane  = threadIdx.x & 31;
    const uint32_t limit = groupList ? groupList[0] : *numGroupsPtr;
    size_t w             = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    while (true)
    {
        if (PERSIST)
        {
            uint32_t next = 0;
            if (lane == 0) { next = atomicAdd(workCounter, 1u); }
            w = __shfl_sync(0xffffffffu, next, 0);
        }
        if (w >= limit) { break; }
        const uint32_t g = groupList ? groupList[1 + w] : uint32_t(w);
        const uint2 grp  = groups[g];
        if (DEFER)
        {
            const uint32_t i = min(grp.x + lane, grp.y - 1);
            const T tx = x[i], ty = y[i], tz = z[i], s = T(2) * T(h[i]);
            // insideBox of findneighbors.hpp:104-106
            const bool inside = (tx - s >= box.lim[0]) && (ty - s >= box.lim[2]) && (tz - s >= box.lim[4]) &&
                                (tx + s <= box.lim[1]) && (ty + s <= box.lim[3]) && (tz + s <= box.lim[5]);
            if (__any_sync(0xffffffffu, !inside))
            {
                if (lane == 0) { deferred[1 + atomicAdd(&deferred[0], 1u)] = g; }
                if (PERSIST) { continue; }
                break;
            }
        }
        warpSearchLanes<T, PBC, FOLD, Th, EXT>(shAll[threadIdx.x >> 5], grp, x, y, z, h, first, box, childOffsets, parents, internalToLeaf,
                                 layout, centers, sizes, ngmax, neighbors, neighborsCount, searchExt);
        if (!PERSIST) { break; }
        __syncwarp();
    }
}

/* ================================================================ group-steered search (trees with small leaves)
 * The cost of the search above grows with the number of leaves a warp visits: staging and the per-lane tests of the
 * nodes are paid per leaf, and the leaf occupancy of a tree jumps by 8x whenever the particle count crosses a power of
 * 8 times the bucket size (64 Mi particles with bucket 64: 32 per leaf, 128 Mi: ~8, 256 Mi: 16).  This variant keeps no
 * per-lane walk state at all and stages contiguous particle ranges whatever the leaf boundaries; its run time hardly
 * depends on the leaf occupancy (64 Mi particles, ng ~ 100: 53.7 / 61.2 / 64 ms at 32 / 8 / 4 particles per leaf, where
 * the per-lane search takes 48.0 / 106 / 160 ms; 32 Mi particles at 16 per leaf: 36.5 vs 37.2 ms, in a periodic box 41.2
 * vs 50 ms).  findNeighbors picks by the mean leaf occupancy (NB_SMALL_LEAVES). */

constexpr int NB_CAP            = 64;  // staged candidates per test round (filled in rounds of up to 32 loads)
constexpr uint32_t NB_COARSE    = 64;  // subtrees with at most this many particles are staged whole (lower limit)
constexpr double NB_SMALL_LEAVES = 20;  // mean particles per leaf below which the group-steered search is used
constexpr uint32_t NB_BAD_LEAF  = 0x80000000u; // in NodeRange::y: a particle of this leaf lies outside the leaf's box

struct alignas(16) GroupWalkShared
{
    // staged candidates: x, y, z (relative floats for T = double, the values themselves for T = float) and particle
    // index, padded to a multiple of four entries
    float cx[NB_CAP + 4], cy[NB_CAP + 4], cz[NB_CAP + 4];
    uint32_t cj[NB_CAP + 4];
    uint8_t wm[NB_MAX_DEPTH + 1]; // per tree depth: the siblings the walk of the warp enters
};

/* ---- preparation: particle range of every node, leaves with stray particles ---- */

//! particles [x, y) below every node (leaves: their layout range)
__global__ void nodeRangeKernel(const int* __restrict__ childOffsets, const int* __restrict__ internalToLeaf,
                                const uint32_t* __restrict__ layout, int numNodes, uint2* __restrict__ nodeRange)
{
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= numNodes) { return; }
    int a = node, b = node;
    for (int c = childOffsets[a]; c != 0; c = childOffsets[a])
        a = c;
    for (int c = childOffsets[b]; c != 0; c = childOffsets[b])
        b = c + 7;
    nodeRange[node] = make_uint2(layout[internalToLeaf[a]], layout[internalToLeaf[b] + 1]);
}

/*! Marks the leaves (and all their ancestors) whose particles are not all inside the leaf's box, or whose box is not
 *  inside the box of every ancestor up to tolNest; 8 lanes per leaf.  Keys are computed from the coordinates, so with
 *  consistent input only particles within a rounding error of a cell face stray (none in double precision, about one
 *  leaf in 200 in single), and the rounded boxes of a tree nest up to 3 ulp of the largest coordinate; particles beyond
 *  the domain box (clamped keys) or arrays that do not belong to the tree are caught as well.  The search evaluates the
 *  reference expressions directly on marked leaves and does not take marked subtrees whole. */
template<class T>
__global__ void leafContainmentKernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                      const int* __restrict__ childOffsets, uint2* __restrict__ nodeRange,
                                      const T* __restrict__ centers, const T* __restrict__ sizes, int numNodes,
                                      T tolNest, const int* __restrict__ parents, uint32_t* __restrict__ numBad)
{
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int node   = int(tid >> 3);
    if (node >= numNodes || childOffsets[node] != 0) { return; }
    const uint2 r = nodeRange[node];
    const T cx = centers[3 * node], cy = centers[3 * node + 1], cz = centers[3 * node + 2];
    const T sx = sizes[3 * node], sy = sizes[3 * node + 1], sz = sizes[3 * node + 2];
    bool bad = false;
    for (uint32_t j = r.x + uint32_t(tid & 7); j < (r.y & ~NB_BAD_LEAF); j += 8)
        bad |= !(rabs(x[j] - cx) <= sx && rabs(y[j] - cy) <= sy && rabs(z[j] - cz) <= sz);
    if ((tid & 7) == 0)
    {
        for (int a = node; a != 0;)
        {
            a = parents[(a - 1) >> 3];
            bad |= !(rabs(cx - centers[3 * a]) + sx <= sizes[3 * a] + tolNest &&
                     rabs(cy - centers[3 * a + 1]) + sy <= sizes[3 * a + 1] + tolNest &&
                     rabs(cz - centers[3 * a + 2]) + sz <= sizes[3 * a + 2] + tolNest);
        }
    }
    if (bad && !(atomicOr(&nodeRange[node].y, NB_BAD_LEAF) & NB_BAD_LEAF))
    {
        atomicAdd(numBad, 1u);
        for (int a = node; a != 0;)
        {
            a = parents[(a - 1) >> 3];
            atomicOr(&nodeRange[a].y, NB_BAD_LEAF);
        }
    }
}

/*! The reference's walk for one lane, restricted to the root-to-leaf path of particle j: true if the lane's own
 *  traversal (findneighbors.hpp:108-112 at every node of the path) reaches the leaf that holds j.  The root has been
 *  tested by the caller.  Rare: only for candidates in the thin shell below the search radius. */
template<bool PBC, class T>
__device__ __forceinline__ bool walkReachesLeafOf(uint32_t j, const Target<T>& t, const int* __restrict__ childOffsets,
                                               const uint2* __restrict__ nodeRange, const T* __restrict__ centers,
                                               const T* __restrict__ sizes, const Box<T>& box)
{
    int node = 0;
    while (true)
    {
        const int child = childOffsets[node];
        if (child == 0) { return true; }
        // the child whose particle range holds j (empty children share their start with the next sibling)
        int k = 0;
#pragma unroll
        for (int m = 1; m < 8; ++m)
            k += int(nodeRange[child + m].x <= j);
        node = child + k;
        if (!cellOverlap<PBC>(t, centers, sizes, node, box)) { return false; }
    }
}

/*! Search of one warp in a periodic box whose fold cannot be applied once per warp (the group or its search spheres are
 *  large against the box, or T = float): every lane keeps the exact state of the reference's own walk (a bit per
 *  sibling and tree depth) and evaluates the reference expressions on broadcast loads. */
template<class T, class Th>
__device__ __noinline__ void warpSearchDirect(uint8_t (*mask)[32], const uint2 grp, const T* __restrict__ x,
                                              const T* __restrict__ y, const T* __restrict__ z,
                                              const Th* __restrict__ h, uint32_t first, const Box<T>& box,
                                              const int* __restrict__ childOffsets, const int* __restrict__ parents,
                                              const uint2* __restrict__ nodeRange, const T* __restrict__ centers,
                                              const T* __restrict__ sizes, uint32_t ngmax,
                                              uint32_t* __restrict__ neighbors, uint32_t* __restrict__ neighborsCount,
                                              float searchExt)
{
    const unsigned lane = threadIdx.x & 31;
    const bool valid    = grp.x + lane < grp.y;
    const uint32_t i    = valid ? grp.x + lane : grp.y - 1;
    Target<T> t;
    t.x         = x[i];
    t.y         = y[i];
    t.z         = z[i];
    const Th hi = h[i];
    {
        const Th radiusSq = Th(4.0) * hi * hi;
        t.radiusSq        = T(radiusSq);
        t.cellRadiusSq    = searchExt == 1.0f ? t.radiusSq : T(radiusSq * searchExt * searchExt);
    }
    {
        T s         = T(2) * T(hi);
        bool inside = (t.x - s >= box.lim[0]) && (t.y - s >= box.lim[2]) && (t.z - s >= box.lim[4]) &&
                      (t.x + s <= box.lim[1]) && (t.y + s <= box.lim[3]) && (t.z + s <= box.lim[5]);
        t.usePbc    = !inside;
    }
    uint32_t* row  = neighbors + size_t(i - first) * size_t(ngmax);
    uint32_t count = 0;

    auto scanLeaf = [&](int node, bool mine)
    {
        const uint2 r     = nodeRange[node];
        const uint32_t je = r.y & ~NB_BAD_LEAF;
        for (uint32_t j = r.x; j < je; ++j)
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
            if (mine && j != i && d2 < t.radiusSq)
            {
                if (count < ngmax) { row[count] = j; }
                ++count;
            }
        }
    };
    auto testChildren = [&](int child0, bool mine) -> uint32_t
    {
        uint32_t bits = 0;
        if (mine)
        {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                bits |= uint32_t(cellOverlap<true>(t, centers, sizes, child0 + c, box)) << c;
        }
        return bits;
    };

    const bool rootMine = valid && cellOverlap<true>(t, centers, sizes, 0, box);
    if (__any_sync(0xffffffffu, rootMine))
    {
        int rootChild = childOffsets[0];
        if (rootChild == 0) { scanLeaf(0, rootMine); }
        else
        {
            int depth     = 1;
            int base      = rootChild;
            uint32_t lm   = testChildren(rootChild, rootMine);
            mask[1][lane] = uint8_t(lm);
            uint32_t wm   = __reduce_or_sync(0xffffffffu, lm);
            while (true)
            {
                if (wm == 0)
                {
                    if (depth == 1) { break; }
                    const int up = parents[(base - 1) >> 3];
                    --depth;
                    base = ((up - 1) & ~7) + 1;
                    lm   = mask[depth][lane];
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
                    lm                = testChildren(child, mine);
                    mask[depth][lane] = uint8_t(lm);
                    wm                = __reduce_or_sync(0xffffffffu, lm);
                    base              = child;
                }
            }
        }
    }
    if (valid) { neighborsCount[i - first] = count; }
}

/*! The search of ONE warp for the targets [grp.x, grp.y) (at most 32), lanes = targets.
 *
 *  The walk is steered by the bounding box of the warp's targets alone: 8 lanes test the 8 children of a node against
 *  it (certified: a child that fails cannot pass the continuation test of any lane), subtrees with few particles are
 *  taken whole, and the particle ranges of the visited leaves - merged where they are contiguous, whatever the leaf
 *  boundaries inside - are staged with coalesced loads, culled per candidate against the same bounding box and tested
 *  by all lanes.  Ranges come in SFC order, so every lane appends in ascending particle index like the CPU walk.
 *
 *  What the reference computes for target i is { j : the walk of i reaches the leaf of j, and d2(i,j) < r2 } with its
 *  own rounded expressions for both conditions.  Each staged pair gets one single-precision sum of squares s:
 *    s > thrHi        certainly d2 >= r2 in the reference's arithmetic                          -> rejected
 *    s < thrLo        certainly d2 < (1 - mu) r2, which implies BOTH conditions: the particle lies inside the box of its
 *                     leaf and of every ancestor up to the tolerance the preparation kernel has checked, so each of
 *                     those boxes is closer to the target than sqrt(1-mu) r + tau0 < r by more than the rounding of the
 *                     reference's box test                                                         -> accepted
 *    in between       (a shell of relative width ~1e-3 below the radius, < 0.5 % of the neighbours) the reference's
 *                     distance expression and the reference's box tests along the root-to-leaf path of j decide.
 *  mu = 4 tau0 / r + 2^-18 with tau0 = 64 ulp(largest coordinate of the box) covers the containment tolerance
 *  (16 ulp), the nesting of the rounded ancestor boxes and the rounding of the box test; lanes whose radius is not
 *  large against tau0 (or not a normal number) take the exact route for every candidate inside the radius.  Leaves
 *  with a stray particle (NB_BAD_LEAF) are never merged or staged: the lanes evaluate the reference expressions on
 *  them directly, and subtrees are not taken whole while such leaves exist.
 *
 *  PBC = false: the box has no periodic dimension, the fold code is not even compiled in.  PBC = true: if the group
 *  and its search spheres are small against the box, every lane sees the same periodic image of a nearby particle or
 *  node, the fold is applied ONCE while staging (to the coordinates relative to the group origin) and everything above
 *  runs unchanged (the exact route uses the reference's folded expressions); otherwise, and for float searches (whose
 *  staged operands must be the reference's), the warp runs warpSearchDirect. */
template<class T, bool PBC, bool FOLD, class Th, bool EXT>
__device__ __forceinline__ void warpSearchGroup(GroupWalkShared& sh, const uint2 grp, const T* __restrict__ x, const T* __restrict__ y,
                           const T* __restrict__ z, const Th* __restrict__ h, uint32_t first, const Box<T>& box,
                           const int* __restrict__ childOffsets, const int* __restrict__ parents,
                           const uint2* __restrict__ nodeRange, float tau0, uint32_t coarse,
                           const T* __restrict__ centers, const T* __restrict__ sizes, uint32_t ngmax,
                           uint32_t* __restrict__ neighbors, uint32_t* __restrict__ neighborsCount,
                           float searchExt)
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
    // Th may be float with double coordinates: the radius is formed in Th and promoted (findneighbors.hpp:89-99)
    const Th hi = h[i];
    {
        const Th radiusSq = Th(4.0) * hi * hi;
        t.radiusSq        = T(radiusSq);
        t.cellRadiusSq    = EXT ? T(radiusSq * searchExt * searchExt) : t.radiusSq; // findneighbors.hpp:99-100
    }
    {
        bool anyPbc = box.pbc(0) || box.pbc(1) || box.pbc(2);
        T s         = T(2) * T(hi);
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
    // the radius of the continuation tests (differs from the search radius only if searchExtFactor != 1)
    const float c2f    = EXT ? float(t.cellRadiusSq) : r2f;
    const float c2max  = EXT ? warpMaxF(c2f) : r2max;
    const float c2bMax = EXT ? ((c2max > 1e-30f) ? c2max * BAND_KB : __int_as_float(0x7f800000)) : r2bMax;

    bool foldOk = Filt && FOLD;
    if (PBC && Filt && FOLD)
    {
        const float reach = 1.01f * sqrtf(fmaxf(r2max, c2max));
        const float ext[3] = {hix - lox, hiy - loy, hiz - loz};
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (box.pbc(d) && !(ext[d] + reach < 0.124f * float(box.len[d]))) { foldOk = false; }
    }
    if (PBC && warpPbc && !foldOk)
    {
        warpSearchDirect<T, Th>(reinterpret_cast<uint8_t(*)[32]>(&sh), grp, x, y, z, h, first, box, childOffsets,
                                parents, nodeRange, centers, sizes, ngmax, neighbors, neighborsCount, searchExt);
        return;
    }
    const bool foldPbc = warpPbc;

    // every staged (un-culled) candidate has |coordinate| <= 1.01 (DwT + sqrt(r2max)), see the cull test
    const float Dpair = 1.01f * (DwT + sqrtf(r2max));
    const float Epair = Dpair * Dpair * 0x1p-29f;

    /* per-lane thresholds on the single-precision sum (see above).  A lane whose own walk does not even enter the root
     * (or that holds no target) accepts nothing. */
    const bool active = valid && cellOverlap<PBC>(t, centers, sizes, 0, box);
    float thrLo = -1.0f, thrHi = -1.0f;
    T sureT     = T(0); // exact route, T = double: d2 < sureT * r2 implies that the walk reaches the leaf
    if (active)
    {
        // the walk is certain to reach the leaf if the particle is inside BOTH radii by the margin mu
        const float own2f = EXT ? fminf(r2f, c2f) : r2f;
        const float mu    = fmaf(4.0f * tau0, rsqrtf(own2f), 0x1p-18f);
        // (tau0 = 12 ulp of the largest box coordinate also bounds the rounding of the reference's box tests only for
        // targets that are not far outside the box themselves: others take the exact route)
        const float tmax  = float(fmax(fmax(fabs(t.x), fabs(t.y)), fabs(t.z)));
        const bool normal = own2f > 1e-30f && mu < 0x1p-6f && tmax * 12.0f * float(sizeof(T) == 8 ? 0x1p-52 : 0x1p-23) <=
                                                                   2.0f * tau0; // false for NaN
        if (normal)
        {
            sureT = T(1.0f - 2.0f * mu);
            if (EXT && t.cellRadiusSq < t.radiusSq) { sureT *= t.cellRadiusSq / t.radiusSq * T(1.0 - 0x1p-20); }
        }
        if (Filt)
        {
            thrHi = r2f > 1e-30f ? fmaf(Epair, BAND_SB, r2f * BAND_KB) : __int_as_float(0x7f800000);
            if (normal) { thrLo = fmaf(-Epair, BAND_SA, (own2f * (1.0f - mu)) * BAND_KA); }
        }
        else
        {
            // the staged values are the reference's operands and s is the reference's expression: s < r2f decides the
            // distance, the largest float below r2f is the upper end of the shell
            if (r2f > 0.0f) { thrHi = __uint_as_float(__float_as_uint(r2f) - 1u); } // +inf -> FLT_MAX; NaN -> stays -1
            if (normal) { thrLo = (own2f * (1.0f - mu)) * (1.0f - 0x1p-22f); }
        }
        if (!(thrHi >= 0.0f)) { thrHi = -1.0f; }
    }
    // bit patterns of the shell [thrLo, thrHi] among the non-negative floats (they order like their patterns; NaN sums
    // lie above +inf and never match: NaN < r2 is false for the reference as well)
    uint32_t bandLo = 0xffffffffu, bandSpan = 0;
    if (thrHi >= 0.0f)
    {
        const uint32_t lo = __float_as_uint(fmaxf(thrLo, 0.0f)), hi2 = __float_as_uint(thrHi);
        if (hi2 >= lo)
        {
            bandLo   = lo;
            bandSpan = hi2 - lo;
        }
    }

    // out = address of the next list entry; entries at and beyond rowEnd are counted but not stored
    // (findneighbors.hpp:139-146).  Global-space byte addresses: the stores of the staged path are inline PTX.
    unsigned long long out          = __cvta_generic_to_global(neighbors + size_t(i - first) * size_t(ngmax));
    const unsigned long long rowEnd = out + 4ull * ngmax;

    auto append = [&](uint32_t j)
    {
        if (out < rowEnd) { asm volatile("st.global.u32 [%0], %1;" ::"l"(out), "r"(j) : "memory"); }
        out += 4;
    };

    // through a shuffle with the own lane: the values become opaque to ptxas, which otherwise re-derives them from the
    // double coordinates (three FP64 subtractions and conversions) in every iteration of the test loop
    const float txs = __shfl_sync(0xffffffffu, -txf, lane), tys = __shfl_sync(0xffffffffu, -tyf, lane),
                tzs = __shfl_sync(0xffffffffu, -tzf, lane);
    const uint64_t ntx2 = pack2(txs, txs), nty2 = pack2(tys, tys), ntz2 = pack2(tzs, tzs);
    const uint64_t zero2 = pack2(0.0f, 0.0f);

    /*! tests the `cnt` staged candidates four at a time (cnt is padded to a multiple of four with candidates at
     *  infinity).  SELF: the batch holds targets of this group, so a candidate can be the target itself (excluded by
     *  index, findneighbors.hpp:131).  GUARD: a list may reach ngmax during this call. */
    auto testStaged = [&](uint32_t cnt, auto selfTag, auto guardTag)
    {
        constexpr bool SELF  = decltype(selfTag)::value;
        constexpr bool GUARD = decltype(guardTag)::value;
        for (uint32_t k = 0; k < cnt; k += 4)
        {
            const float4 X = *reinterpret_cast<const float4*>(&sh.cx[k]);
            const float4 Y = *reinterpret_cast<const float4*>(&sh.cy[k]);
            const float4 Z = *reinterpret_cast<const float4*>(&sh.cz[k]);
            const uint4 J  = *reinterpret_cast<const uint4*>(&sh.cj[k]);
            float s[4];
#pragma unroll
            for (int half = 0; half < 2; ++half)
            {
                const uint64_t x2 = half ? pack2(X.z, X.w) : pack2(X.x, X.y);
                const uint64_t y2 = half ? pack2(Y.z, Y.w) : pack2(Y.x, Y.y);
                const uint64_t z2 = half ? pack2(Z.z, Z.w) : pack2(Z.x, Z.y);
                const uint64_t dx = add2(x2, ntx2), dy = add2(y2, nty2), dz = add2(z2, ntz2);
                uint64_t s2;
                if (Filt) { s2 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz))); }
                else
                {
                    // the reference's float expression (dx*dx + dy*dy) + dz*dz (findneighbors.hpp:33-60), bit for bit:
                    // the packed operations round per component like the scalar ones.  The products are formed as
                    // fma(d, d, +0) = RN(d*d): separate mul.rn / add.rn.f32x2 get contracted to FFMA2 by ptxas even
                    // under --fmad=false
                    s2 = add2(add2(fma2(dx, dx, zero2), fma2(dy, dy, zero2)), fma2(dz, dz, zero2));
                }
                unpack2(s2, s[2 * half], s[2 * half + 1]);
            }
            // a sum in the shell: the reference's own expressions decide.  The branch is warp-uniform.
            const bool amb = __float_as_uint(s[0]) - bandLo <= bandSpan || __float_as_uint(s[1]) - bandLo <= bandSpan ||
                             __float_as_uint(s[2]) - bandLo <= bandSpan || __float_as_uint(s[3]) - bandLo <= bandSpan;
            if (__any_sync(0xffffffffu, amb))
            {
#pragma unroll 1
                for (int c = 0; c < 4; ++c)
                {
                    // (selects instead of indexed arrays: the loop is kept rolled, the rare route stays small)
                    const uint32_t jc = c == 0 ? J.x : c == 1 ? J.y : c == 2 ? J.z : J.w;
                    const float sc    = c == 0 ? s[0] : c == 1 ? s[1] : c == 2 ? s[2] : s[3];
                    bool in           = sc < thrLo;
                    if (!in && sc <= thrHi && k + c < cnt)
                    {
                        bool sure = false; // inside by more than the margin the walk needs
                        if (Filt)
                        {
                            // the reference's expression (findneighbors.hpp:33-60,134), folded if the target's
                            // search sphere leaves the box (:104-106)
                            T ex = x[jc] - t.x, ey = y[jc] - t.y, ez = z[jc] - t.z;
                            if (PBC && foldPbc && t.usePbc)
                            {
                                ex = pbcFold(ex, 0, box);
                                ey = pbcFold(ey, 1, box);
                                ez = pbcFold(ez, 2, box);
                            }
                            const T d2 = ex * ex + ey * ey + ez * ez;
                            in   = d2 < t.radiusSq;
                            sure = d2 < t.radiusSq * sureT;
                        }
                        else { in = true; }
                        if (in && !sure)
                        {
                            in = walkReachesLeafOf<PBC>(jc, t, childOffsets, nodeRange, centers, sizes, box);
                        }
                    }
                    if (in && jc != i) { append(jc); }
                }
                continue;
            }
            uint32_t outLo = uint32_t(out), outHi = uint32_t(out >> 32);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.x, s[0], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.y, s[1], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.z, s[2], thrLo, i);
            appendIf<SELF, GUARD>(outLo, outHi, rowEnd, J.w, s[3], thrLo, i);
            out = (unsigned long long)(outHi) << 32 | outLo;
        }
    };

    uint32_t bcnt = 0;     // candidates staged
    bool bself    = false; // the staged ranges hold targets of this group

    auto flush = [&]()
    {
        // pad to a multiple of four with candidates at infinity: s = +inf is never below thrLo, and on the exact
        // route the entry index is checked
        if (lane < 4)
        {
            const float inf    = __int_as_float(0x7f800000);
            sh.cx[bcnt + lane] = inf;
            sh.cy[bcnt + lane] = inf;
            sh.cz[bcnt + lane] = inf;
            sh.cj[bcnt + lane] = i;
        }
        __syncwarp();
        const bool guard = __any_sync(0xffffffffu, out + 4ull * bcnt > rowEnd);
        if (guard) { testStaged(bcnt, std::true_type{}, std::true_type{}); }
        else if (bself) { testStaged(bcnt, std::true_type{}, std::false_type{}); }
        else { testStaged(bcnt, std::false_type{}, std::false_type{}); }
        __syncwarp();
        bcnt  = 0;
        bself = false;
    };

    //! stage the particles [pb, pe) in rounds of 32 coalesced loads, dropping those no lane can reach; the staged
    //! candidates are tested whenever another round might not fit and, if `drain`, at the end
    auto stageRange = [&](uint32_t pb, uint32_t pe, bool drain)
    {
        for (uint32_t base = pb;; base += 32)
        {
            const bool more = base < pe;
            if (bcnt && (more ? bcnt + 32 > NB_CAP : drain)) { flush(); }
            if (!more) { break; }
            const uint32_t j = base + lane;
            bool keep        = false;
            float c0 = 0, c1 = 0, c2 = 0;
            if (j < pe)
            {
                T qx = x[j] - ox, qy = y[j] - oy, qz = z[j] - oz;
                if (PBC && foldPbc)
                {
                    qx = pbcFold(qx, 0, box);
                    qy = pbcFold(qy, 1, box);
                    qz = pbcFold(qz, 2, box);
                }
                c0 = float(qx);
                c1 = float(qy);
                c2 = float(qz);
                float D  = fmaxf(fmaxf(fabsf(c0), fabsf(c1)), fmaxf(fabsf(c2), DwT));
                float bc = fmaf(D * D * 0x1p-27f, BAND_SB, r2bMax);
                float ex = fmaxf(fmaxf(lox - c0, c0 - hix), 0.0f);
                float ey = fmaxf(fmaxf(loy - c1, c1 - hiy), 0.0f);
                float ez = fmaxf(fmaxf(loz - c2, c2 - hiz), 0.0f);
                keep     = !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > bc);
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (keep)
            {
                const uint32_t pos = bcnt + __popc(km & ltMask);
                sh.cx[pos]         = c0;
                sh.cy[pos]         = c1;
                sh.cz[pos]         = c2;
                sh.cj[pos]         = j;
            }
            bcnt += __popc(km);
            bself |= base < grp.y && base + 32 > grp.x;
        }
    };

    //! a leaf with stray particles: the reference expressions on broadcast loads for the lanes whose walk reaches it
    auto scanLeafDirect = [&](uint32_t pb, uint32_t pe, bool mine)
    {
        for (uint32_t j = pb; j < pe; ++j)
        {
            T dx = x[j] - t.x;
            T dy = y[j] - t.y;
            T dz = z[j] - t.z;
            if (PBC)
            {
                T fx = pbcFold(dx, 0, box);
                T fy = pbcFold(dy, 1, box);
                T fz = pbcFold(dz, 2, box);
                dx   = t.usePbc ? fx : dx;
                dy   = t.usePbc ? fy : dy;
                dz   = t.usePbc ? fz : dz;
            }
            T d2 = dx * dx + dy * dy + dz * dz;
            if (mine && j != i && d2 < t.radiusSq) { append(j); }
        }
    };

    //! which of the 8 children of a node can hold a neighbour of any target of the warp (lanes 0-7 test one each)
    auto reachableChildren = [&](int child0) -> uint32_t
    {
        bool reach = false;
        if (lane < 8)
        {
            const int node = child0 + int(lane);
            T rx = centers[3 * node] - ox, ry = centers[3 * node + 1] - oy, rz = centers[3 * node + 2] - oz;
            const float gx = float(sizes[3 * node]), gy = float(sizes[3 * node + 1]), gz = float(sizes[3 * node + 2]);
            bool large = false; // a node that is large against the periodic box: the lanes may see different images
            if (PBC && foldPbc)
            {
                rx    = pbcFold(rx, 0, box);
                ry    = pbcFold(ry, 1, box);
                rz    = pbcFold(rz, 2, box);
                large = (box.pbc(0) && !(gx < 0.124f * float(box.len[0]))) ||
                        (box.pbc(1) && !(gy < 0.124f * float(box.len[1]))) ||
                        (box.pbc(2) && !(gz < 0.124f * float(box.len[2])));
            }
            const float cx = float(rx), cy = float(ry), cz = float(rz);
            const float D  = fmaxf(fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fmaxf(fabsf(cz), DwT)), fmaxf(gx, fmaxf(gy, gz)));
            const float E  = D * D * 0x1p-27f;
            // a child whose box is certainly farther from the targets' bounding box than the largest radius fails the
            // continuation test of every lane
            float ex = fmaxf(fmaxf(lox - (cx + gx), (cx - gx) - hix), 0.0f);
            float ey = fmaxf(fmaxf(loy - (cy + gy), (cy - gy) - hiy), 0.0f);
            float ez = fmaxf(fmaxf(loz - (cz + gz), (cz - gz) - hiz), 0.0f);
            reach    = large || !(fmaf(ex, ex, fmaf(ey, ey, ez * ez)) > fmaf(E, BAND_SB, c2bMax));
        }
        return __ballot_sync(0xffffffffu, reach) & 0xffu;
    };

    if (__any_sync(0xffffffffu, active))
    {
        const int rootChild = childOffsets[0];
        if (rootChild == 0)
        {
            const uint2 r = nodeRange[0];
            scanLeafDirect(r.x, r.y & ~NB_BAD_LEAF, active);
        }
        else
        {
            // depth-first walk in SFC order over the children the warp enters; `wm` = siblings still to visit at the
            // current depth (the full set is kept per depth in shared memory for the way back up); [pb, pe) = the
            // contiguous particle range collected from the last leaves, not staged yet
            uint32_t pb = 0, pe = 0;
            int depth   = 1;
            int base    = rootChild;
            uint32_t wm = reachableChildren(rootChild);
            if (lane == 0) { sh.wm[1] = uint8_t(wm); }
            while (true)
            {
                uint32_t nb = 0, ne = 0; // the next range, if this step finds one that does not extend [pb, pe)
                bool done = false, bad = false;
                if (wm == 0)
                {
                    if (depth == 1) { done = true; }
                    else
                    {
                        const int up = parents[(base - 1) >> 3];
                        --depth;
                        base = ((up - 1) & ~7) + 1;
                        __syncwarp();
                        wm = uint32_t(sh.wm[depth]) & ~((2u << ((up - 1) & 7)) - 1u);
                        continue;
                    }
                }
                else
                {
                    const int c = __ffs(int(wm)) - 1;
                    wm &= wm - 1;
                    const int node  = base + c;
                    const int child = childOffsets[node];
                    const uint2 r   = nodeRange[node];
                    ne              = r.y & ~NB_BAD_LEAF;
                    nb              = r.x;
                    if (child != 0 && ((r.y & NB_BAD_LEAF) || ne - nb > coarse))
                    {
                        ++depth;
                        wm = reachableChildren(child);
                        if (lane == 0) { sh.wm[depth] = uint8_t(wm); }
                        base = child;
                        continue;
                    }
                    if (ne == nb) { continue; }
                    bad = (r.y & NB_BAD_LEAF) != 0;
                    if (!bad && pe == nb)
                    {
                        pe = ne;
                        continue;
                    }
                }
                // the collected range ends here: a range that is not contiguous with it, a leaf with stray particles,
                // or the end of the walk
                stageRange(pb, pe, done || bad);
                if (done) { break; }
                if (bad)
                {
                    const bool mine = active && walkReachesLeafOf<PBC>(nb, t, childOffsets, nodeRange, centers, sizes, box);
                    scanLeafDirect(nb, ne, mine);
                    pb = pe = 0;
                }
                else
                {
                    pb = nb;
                    pe = ne;
                }
            }
        }
    }

    if (valid) { neighborsCount[i - first] = ngmax - uint32_t((long long)(rowEnd - out) >> 2); }
}

//! the kernel of the group-steered search; DEFER as in findNeighborsKernel.  Does nothing if too many leaves hold stray
//! particles (trees over 32-bit keys: the key grid is coarse against the search radius), findNeighborsKernel runs then
template<class T, bool PBC, bool FOLD, bool DEFER, class Th, bool PERSIST, bool EXT>
__global__ void __launch_bounds__(NB_THREADS, 8) findNeighborsGroupKernel(const T* __restrict__ x,
                                                                  const T* __restrict__ y,
                                                                  const T* __restrict__ z,
                                                                  const Th* __restrict__ h,
                                                                  uint32_t first,
                                                                  const uint2* __restrict__ groups,
                                                                  const uint32_t* __restrict__ numGroupsPtr,
                                                                  const uint32_t* __restrict__ groupList,
                                                                  uint32_t* __restrict__ deferred,
                                                                  Box<T> box,
                                                                  const int* __restrict__ childOffsets,
                                                                  const int* __restrict__ parents,
                                                                  const uint2* __restrict__ nodeRange,
                                                                  float tau0,
                                                                  uint32_t coarse,
                                                                  const T* __restrict__ centers,
                                                                  const T* __restrict__ sizes,
                                                                  uint32_t ngmax,
                                                                  uint32_t* __restrict__ neighbors,
                                                                  uint32_t* __restrict__ neighborsCount,
                                                                  const uint32_t* __restrict__ numBad,
                                                                  uint32_t badLimit,
                                                                  uint32_t* __restrict__ workCounter,
                                                                  float searchExt)
{
    if (*numBad > badLimit) { return; }
    __shared__ GroupWalkShared shAll[NB_THREADS / 32];
    /* PERSIST: groups are handed out through a device counter to one wave of blocks, so the grid does not depend on the
     * (upper bound of the) number of groups and a kernel that has nothing to do costs next to nothing; otherwise one
     * group per warp of the grid.  Which is faster depends on the variant (register allocation), see findNeighbors */
    const unsigned lane  = threadIdx.x & 31;
    const uint32_t limit = groupList ? groupList[0] : *numGroupsPtr;
    size_t w             = (size_t(blockIdx.x) * NB_THREADS + threadIdx.x) >> 5;
    while (true)
    {
        if (PERSIST)
        {
            uint32_t next = 0;
            if (lane == 0) { next = atomicAdd(workCounter, 1u); }
            w = __shfl_sync(0xffffffffu, next, 0);
        }
        if (w >= limit) { break; }
        const uint32_t g = groupList ? groupList[1 + w] : uint32_t(w);
        const uint2 grp  = groups[g];
        if (DEFER)
        {
            const uint32_t i = min(grp.x + lane, grp.y - 1);
            const T tx = x[i], ty = y[i], tz = z[i], s = T(2) * T(h[i]);
            // insideBox of findneighbors.hpp:104-106
            const bool inside = (tx - s >= box.lim[0]) && (ty - s >= box.lim[2]) && (tz - s >= box.lim[4]) &&
                                (tx + s <= box.lim[1]) && (ty + s <= box.lim[3]) && (tz + s <= box.lim[5]);
            if (__any_sync(0xffffffffu, !inside))
            {
                if (lane == 0) { deferred[1 + atomicAdd(&deferred[0], 1u)] = g; }
                if (PERSIST) { continue; }
                break;
            }
        }
        warpSearchGroup<T, PBC, FOLD, Th, EXT>(shAll[threadIdx.x >> 5], grp, x, y, z, h, first, box, childOffsets, parents, nodeRange,
                                     tau0, coarse, centers, sizes, ngmax, neighbors, neighborsCount, searchExt);
        if (!PERSIST) { break; }
        __syncwarp();
    }
}

/*! periodic boxes: groups whose targets all keep their search spheres inside the box (findneighbors.hpp:104-106) go to
 *  `interior`, the others to `boundary` ([0] = count, then group numbers): the interior groups are then searched by the
 *  open-box kernel, the few boundary groups by the periodic one */
template<class T, class Th>
__global__ void classifyGroupsKernel(const T* __restrict__ x, const T* __restrict__ y, const T* __restrict__ z,
                                     const Th* __restrict__ h, const uint2* __restrict__ groups,
                                     const uint32_t* __restrict__ numGroupsPtr, Box<T> box,
                                     uint32_t* __restrict__ interior, uint32_t* __restrict__ boundary)
{
    const size_t w = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (w >= size_t(*numGroupsPtr)) { return; }
    const unsigned lane = threadIdx.x & 31;
    const uint2 grp     = groups[w];
    const uint32_t i    = min(grp.x + lane, grp.y - 1);
    const T tx = x[i], ty = y[i], tz = z[i], s = T(2) * T(h[i]);
    const bool inside = (tx - s >= box.lim[0]) && (ty - s >= box.lim[2]) && (tz - s >= box.lim[4]) &&
                        (tx + s <= box.lim[1]) && (ty + s <= box.lim[3]) && (tz + s <= box.lim[5]);
    const bool atBoundary = __any_sync(0xffffffffu, !inside);
    if (lane == 0)
    {
        uint32_t* list                 = atBoundary ? boundary : interior;
        list[1 + atomicAdd(&list[0], 1u)] = uint32_t(w);
    }
}

} // namespace

template<class T, class Th, bool EXT>
int findNeighborsImpl(const T* x, const T* y, const T* z, const Th* h, uint32_t first, uint32_t last, const double* lim,
                      const int* bnd, int numLeaves, const int* childOffsets, const int* parents,
                      const int* internalToLeaf, const uint32_t* layout, const T* centers, const T* sizes,
                      uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount, cudaStream_t s, float searchExt)
{
    CSB_REQUIRE(last >= first, "invalid particle range");
    CSB_REQUIRE(numLeaves >= 1, "empty tree");
    if (last == first) { return 0; }
    Box<T> box         = makeBox<T>(lim, bnd);
    const bool pbc     = box.pbc(0) || box.pbc(1) || box.pbc(2);
    const int numNodes = numLeaves + (numLeaves - 1) / 7;

    // target groups: counts per leaf -> exclusive scan -> fill.  Upper bound on the number of groups is known on the
    // host, the exact number stays on the device (no synchronisation)
    size_t maxGroups = size_t(numLeaves) + (size_t(last) - first) / 32 + 1;
    CSB_SCRATCH(groupOffsets, uint32_t*, s, SCRATCH_A, (size_t(numLeaves) + 1) * sizeof(uint32_t));
    CSB_SCRATCH(groups, uint2*, s, SCRATCH_B, maxGroups * sizeof(uint2));
    CSB_SCRATCH(scanTmp, void*, s, SCRATCH_C, scanTempBytes(size_t(numLeaves) + 1));
    CSB_CHECK(cudaMemsetAsync(groupOffsets, 0, (size_t(numLeaves) + 1) * sizeof(uint32_t), s));

    const bool leafAligned = tuning(TUNE_NB_GROUPS) != 0;
    auto countKernel       = leafAligned ? groupBuildKernel<false, 0> : groupBuildKernel<false, 1>;
    auto fillKernel        = leafAligned ? groupBuildKernel<true, 0> : groupBuildKernel<true, 1>;
    countKernel<<<iceil(numNodes, 256), 256, 0, s>>>(childOffsets, internalToLeaf, layout, numNodes, first, last,
                                                     groupOffsets, nullptr, nullptr);
    CSB_LAUNCH_CHECK();
    if (int e = exclusiveScanU32(groupOffsets, groupOffsets, size_t(numLeaves) + 1, scanTmp, s)) { return e; }
    fillKernel<<<iceil(numNodes, 256), 256, 0, s>>>(childOffsets, internalToLeaf, layout, numNodes, first, last, nullptr,
                                                    groupOffsets, groups);
    CSB_LAUNCH_CHECK();

    // persistent grids (one wave of blocks per kernel); work[k] = next group of the k-th search launch; the list of the
    // groups deferred to the periodic kernel follows
    CSB_SCRATCH(work, uint32_t*, s, SCRATCH_E, (2 * maxGroups + 8) * sizeof(uint32_t));
    CSB_CHECK(cudaMemsetAsync(work, 0, 5 * sizeof(uint32_t), s));
    uint32_t* deferred = work + 4;
    uint32_t* interior = deferred + maxGroups + 1; // used when the groups are classified up front (see below)
    int device = 0, numSms = 0;
    CSB_CHECK(cudaGetDevice(&device));
    CSB_CHECK(cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, device));
    const unsigned fullGrid = iceil(maxGroups * 32, NB_THREADS);
    auto waveOf = [&](auto kernel) -> unsigned
    {
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, NB_THREADS, 0) != cudaSuccess || perSm < 1)
        {
            perSm = 4;
        }
        return unsigned(std::min<size_t>(size_t(numSms) * perSm, iceil(maxGroups * 32, NB_THREADS)));
    };

    /* Trees with few particles per leaf (T = double): the group-steered search, preceded by its preparation - the
     * particle range below every node and the leaves with stray particles (the tolerances scale with the rounding unit
     * of the largest coordinate of the box, see warpSearchGroup).  It only runs if at most one leaf in 32 holds stray
     * particles, otherwise the per-lane search below does; both decide from the same device counter, no host
     * synchronisation.  TUNE_NB_SEARCH: 0 = by leaf occupancy, 1 = per-lane walks always, 2 = group-steered always. */
    const int searchKnob   = tuning(TUNE_NB_SEARCH);
    const bool groupSearch = sizeof(T) == 8 && searchKnob != 1 &&
                             (searchKnob == 2 || double(last - first) < NB_SMALL_LEAVES * double(numLeaves));
    const uint32_t* numBadGate = nullptr;
    const uint32_t badLimit    = uint32_t(numLeaves) / 32;
    if constexpr (sizeof(T) == 8)
    {
        if (groupSearch)
        {
            double cabs = 0;
            for (int d = 0; d < 6; ++d)
                cabs = std::max(cabs, std::fabs(lim[d]));
            const double eps = 0x1p-52;
            const T tolNest  = T(4 * eps * cabs);
            const float tau0 = float(12 * eps * cabs) * (1.0f + 0x1p-20f);
            CSB_SCRATCH(prep, char*, s, SCRATCH_D, 16 + size_t(numNodes) * sizeof(uint2));
            uint32_t* numBad = reinterpret_cast<uint32_t*>(prep);
            uint2* nodeRange = reinterpret_cast<uint2*>(prep + 16);
            CSB_CHECK(cudaMemsetAsync(numBad, 0, sizeof(uint32_t), s));
            nodeRangeKernel<<<iceil(numNodes, 256), 256, 0, s>>>(childOffsets, internalToLeaf, layout, numNodes,
                                                                 nodeRange);
            CSB_LAUNCH_CHECK();
            leafContainmentKernel<T><<<iceil(size_t(numNodes) * 8, 256), 256, 0, s>>>(
                x, y, z, childOffsets, nodeRange, centers, sizes, numNodes, tolNest, parents, numBad);
            CSB_LAUNCH_CHECK();
            numBadGate = numBad;
            // subtrees up to this size are staged whole: about one level above the leaves (measured: 32 Mi particles at
            // 16 per leaf in a periodic box 41.2 / 38.6 / 36.3 ms for 64 / 128 / 256, 64 Mi at 8 per leaf 61.3 / 61.2 /
            // 63.0 ms)
            const double meanLeaf = double(last - first) / double(numLeaves);
            const uint32_t coarse = uint32_t(std::min(std::max(12.0 * meanLeaf, double(NB_COARSE)), 256.0));
            if (!pbc)
            {
                auto k0 = findNeighborsGroupKernel<T, false, false, false, Th, false, EXT>;
                k0<<<fullGrid, NB_THREADS, 0, s>>>(
                    x, y, z, h, first, groups, groupOffsets + numLeaves, nullptr, nullptr, box, childOffsets, parents,
                    nodeRange, tau0, coarse, centers, sizes, ngmax, neighbors, neighborsCount, numBad, badLimit, work + 0, searchExt);
            }
            else
            {
                auto k0 = findNeighborsGroupKernel<T, false, false, true, Th, false, EXT>;
                k0<<<fullGrid, NB_THREADS, 0, s>>>(
                    x, y, z, h, first, groups, groupOffsets + numLeaves, nullptr, deferred, box, childOffsets, parents,
                    nodeRange, tau0, coarse, centers, sizes, ngmax, neighbors, neighborsCount, numBad, badLimit, work + 0, searchExt);
                CSB_LAUNCH_CHECK();
                auto k1 = findNeighborsGroupKernel<T, true, true, false, Th, false, EXT>;
                k1<<<fullGrid, NB_THREADS, 0, s>>>(
                    x, y, z, h, first, groups, groupOffsets + numLeaves, deferred, nullptr, box, childOffsets, parents,
                    nodeRange, tau0, coarse, centers, sizes, ngmax, neighbors, neighborsCount, numBad, badLimit, work + 1, searchExt);
            }
            CSB_LAUNCH_CHECK();
        }
    }

    /* The per-lane search.  Measured at 64 Mi particles: the open-box kernel for double is faster with persistent warps
     * (48.0 vs 51.5 ms), the float and periodic variants are faster with one group per warp of the grid (28.5 vs 33.9,
     * 56.9 vs 62.7 ms); kernels that only run if the group-steered search declined are launched persistent, which makes
     * the launch that finds nothing to do free. */
    auto launchLanes = [&](auto kernel, bool persist, const uint32_t* groupList, uint32_t* deferredOut, uint32_t* counter)
    {
        kernel<<<persist ? waveOf(kernel) : fullGrid, NB_THREADS, 0, s>>>(
            x, y, z, h, first, groups, groupOffsets + numLeaves, groupList, deferredOut, box, childOffsets, parents,
            internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount, numBadGate, badLimit, counter,
            searchExt);
    };
    const bool gated = numBadGate != nullptr;
    if (!pbc)
    {
        if (gated || sizeof(T) == 8)
        {
            launchLanes(findNeighborsKernel<T, false, false, false, Th, true, EXT>, true, nullptr, nullptr, work + 2);
        }
        else { launchLanes(findNeighborsKernel<T, false, false, false, Th, false, EXT>, false, nullptr, nullptr, work + 2); }
    }
    else if (gated)
    {
        launchLanes(findNeighborsKernel<T, false, false, true, Th, true, EXT>, true, nullptr, deferred, work + 2);
        CSB_LAUNCH_CHECK();
        launchLanes(findNeighborsKernel<T, true, true, false, Th, true, EXT>, true, deferred, nullptr, work + 3);
    }
    else if (sizeof(T) == 8)
    {
        // the groups are classified up front, so that the interior ones can take the persistent open-box kernel
        CSB_CHECK(cudaMemsetAsync(interior, 0, sizeof(uint32_t), s));
        classifyGroupsKernel<T, Th><<<fullGrid, NB_THREADS, 0, s>>>(x, y, z, h, groups, groupOffsets + numLeaves, box,
                                                                    interior, deferred);
        CSB_LAUNCH_CHECK();
        launchLanes(findNeighborsKernel<T, false, false, false, Th, true, EXT>, true, interior, nullptr, work + 2);
        CSB_LAUNCH_CHECK();
        launchLanes(findNeighborsKernel<T, true, true, false, Th, false, EXT>, false, deferred, nullptr, work + 3);
    }
    else
    {
        // interior groups with the open-box code, then the groups at the periodic boundaries
        launchLanes(findNeighborsKernel<T, false, false, true, Th, false, EXT>, false, nullptr, deferred, work + 2);
        CSB_LAUNCH_CHECK();
        launchLanes(findNeighborsKernel<T, true, true, false, Th, false, EXT>, false, deferred, nullptr, work + 3);
    }
    CSB_LAUNCH_CHECK();
    return 0;
}

/*! findNeighbors (findneighbors.hpp:156-177).  searchExtFactor (OctreeNsView, tree/octree.hpp:279-282): the
 *  continuation tests of the walk use the radius 2h * searchExtFactor (findneighbors.hpp:100), acceptance stays at 2h;
 *  the code for factors other than 1 is a separate instantiation so that the default path is unchanged */
template<class T, class Th>
int findNeighbors(const T* x, const T* y, const T* z, const Th* h, uint32_t first, uint32_t last, const double* lim,
                  const int* bnd, int numLeaves, const int* childOffsets, const int* parents, const int* internalToLeaf,
                  const uint32_t* layout, const T* centers, const T* sizes, uint32_t ngmax, uint32_t* neighbors,
                  uint32_t* neighborsCount, cudaStream_t s, float searchExtFactor)
{
    CSB_REQUIRE(searchExtFactor > 0, "findNeighbors: searchExtFactor must be positive");
    if (searchExtFactor == 1.0f)
    {
        return findNeighborsImpl<T, Th, false>(x, y, z, h, first, last, lim, bnd, numLeaves, childOffsets, parents,
                                               internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                               s, 1.0f);
    }
    return findNeighborsImpl<T, Th, true>(x, y, z, h, first, last, lim, bnd, numLeaves, childOffsets, parents,
                                          internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount, s,
                                          searchExtFactor);
}

template int findNeighbors<float, float>(const float*, const float*, const float*, const float*, uint32_t, uint32_t,
                                         const double*, const int*, int, const int*, const int*, const int*,
                                         const uint32_t*, const float*, const float*, uint32_t, uint32_t*, uint32_t*,
                                         cudaStream_t, float);
template int findNeighbors<double, double>(const double*, const double*, const double*, const double*, uint32_t,
                                           uint32_t, const double*, const int*, int, const int*, const int*, const int*,
                                           const uint32_t*, const double*, const double*, uint32_t, uint32_t*,
                                           uint32_t*, cudaStream_t, float);
template int findNeighbors<double, float>(const double*, const double*, const double*, const float*, uint32_t, uint32_t,
                                          const double*, const int*, int, const int*, const int*, const int*,
                                          const uint32_t*, const double*, const double*, uint32_t, uint32_t*, uint32_t*,
                                          cudaStream_t, float);

} // namespace csb

extern "C"
{

int cs_find_neighbors_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                        const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount, void* stream)
{
    return csb::findNeighbors<float, float>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                     internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                     cudaStream_t(stream), 1.0f);
}

int cs_find_neighbors_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                        uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                        const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                        const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                        void* stream)
{
    return csb::findNeighbors<double, double>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                      internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                      cudaStream_t(stream), 1.0f);
}

/* double coordinates, float smoothing lengths (Th != Tc, findneighbors.hpp:89-99): radiusSq is formed in float */
int cs_find_neighbors_df(const double* x, const double* y, const double* z, const float* h, uint32_t firstId,
                         uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                         const int* parents, const int* internalToLeaf, const uint32_t* layout, const double* centers,
                         const double* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                         void* stream)
{
    return csb::findNeighbors<double, float>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                             internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                             cudaStream_t(stream), 1.0f);
}

/* the same with OctreeNsView::searchExtFactor (tree/octree.hpp:279-282): node continuation tests with the radius
 * 2h * searchExtFactor (findneighbors.hpp:100) */
int cs_find_neighbors_ext_f(const float* x, const float* y, const float* z, const float* h, uint32_t firstId,
                            uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                            const int* parents, const int* internalToLeaf, const uint32_t* layout, const float* centers,
                            const float* sizes, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                            float searchExtFactor, void* stream)
{
    return csb::findNeighbors<float, float>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                            internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                            cudaStream_t(stream), searchExtFactor);
}

int cs_find_neighbors_ext_d(const double* x, const double* y, const double* z, const double* h, uint32_t firstId,
                            uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                            const int* parents, const int* internalToLeaf, const uint32_t* layout,
                            const double* centers, const double* sizes, uint32_t ngmax, uint32_t* neighbors,
                            uint32_t* neighborsCount, float searchExtFactor, void* stream)
{
    return csb::findNeighbors<double, double>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                              internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                              cudaStream_t(stream), searchExtFactor);
}

int cs_find_neighbors_ext_df(const double* x, const double* y, const double* z, const float* h, uint32_t firstId,
                             uint32_t lastId, const double* lim, const int* bnd, int numLeaves, const int* childOffsets,
                             const int* parents, const int* internalToLeaf, const uint32_t* layout,
                             const double* centers, const double* sizes, uint32_t ngmax, uint32_t* neighbors,
                             uint32_t* neighborsCount, float searchExtFactor, void* stream)
{
    return csb::findNeighbors<double, float>(x, y, z, h, firstId, lastId, lim, bnd, numLeaves, childOffsets, parents,
                                             internalToLeaf, layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                             cudaStream_t(stream), searchExtFactor);
}

} // extern "C"
