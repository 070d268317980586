/* TEST INFRASTRUCTURE ONLY — not part of the product; never linked into libcstone_b200.so.
 *
 * Thin extern "C" shim over the UNMODIFIED reference headers in /root/reference/include (execution::Cpu path),
 * compiled by oracle/Makefile into oracle/_ref/libcstone_ref.so.  Used by tests/ to (a) pin the C restatement in
 * oracle/cstone_oracle.c and (b) act as the strongest available oracle for GPU parity, and by bench.py's
 * cpu_baseline / --impl reference leg.  No reference source is copied: this file only #includes it.
 *
 * MPI is provided by oracle/mpi_shim/mpi.h (threads as ranks).
 */
#include <mpi.h>

#include <omp.h>

#include <chrono>
#include <cstdint>
#include <string>
#include <cstring>
#include <functional>
#include <numeric>
#include <span>
#include <thread>
#include <tuple>
#include <vector>

#include "cstone/domain/domain.hpp"
#include "cstone/findneighbors.hpp"
#include "cstone/focus/source_center.hpp"
#include "cstone/primitives/gather.hpp"
#include "cstone/sfc/sfc.hpp"
#include "cstone/traversal/collisions.hpp"
#include "cstone/tree/csarray.hpp"
#include "cstone/tree/octree.hpp"

#include "coord_samples/plummer.hpp"

using namespace cstone;

namespace
{

template<class T>
Box<T> makeBox(const double* lim, const int* bnd)
{
    return Box<T>(T(lim[0]), T(lim[1]), T(lim[2]), T(lim[3]), T(lim[4]), T(lim[5]), BoundaryType(bnd[0]),
                  BoundaryType(bnd[1]), BoundaryType(bnd[2]));
}

template<class T>
void storeBox(const Box<T>& b, double* lim)
{
    lim[0] = b.xmin(), lim[1] = b.xmax(), lim[2] = b.ymin(), lim[3] = b.ymax(), lim[4] = b.zmin(), lim[5] = b.zmax();
}

template<class KeyType, class T>
void sfcKeys(int kind, const T* x, const T* y, const T* z, KeyType* keys, size_t n, const double* lim, const int* bnd)
{
    auto box = makeBox<T>(lim, bnd);
    if (kind == 0) { computeSfcKeys(x, y, z, reinterpret_cast<HilbertKey<KeyType>*>(keys), n, box); }
    else { computeSfcKeys(x, y, z, reinterpret_cast<MortonKey<KeyType>*>(keys), n, box); }
}

template<class KeyType>
long computeTree(const KeyType* keys, size_t n, unsigned bucket, KeyType* leaves, unsigned* counts, long cap)
{
    auto [t, c] = computeOctree<KeyType>(std::span<const KeyType>(keys, n), bucket);
    long nl     = long(nNodes(t));
    if (nl > cap) return -nl;
    std::copy(t.begin(), t.end(), leaves);
    std::copy(c.begin(), c.end(), counts);
    return nl;
}

template<class KeyType>
long updateTree(const KeyType* keys,
                size_t n,
                unsigned bucket,
                KeyType* leaves,
                unsigned* counts,
                long nLeaves,
                long cap,
                int* converged)
{
    std::vector<KeyType> t(leaves, leaves + nLeaves + 1);
    std::vector<unsigned> c(counts, counts + nLeaves);
    *converged = updateOctree<KeyType>(std::span<const KeyType>(keys, n), bucket, t, c);
    long nl    = long(nNodes(t));
    if (nl > cap) return -nl;
    std::copy(t.begin(), t.end(), leaves);
    std::copy(c.begin(), c.end(), counts);
    return nl;
}

template<class KeyType>
void linkTree(const KeyType* leaves,
              int nLeaves,
              KeyType* prefixes,
              int* childOffsets,
              int* parents,
              int* levelRange,
              int* internalToLeaf,
              int* leafToInternal)
{
    int numInternal = (nLeaves - 1) / 7;
    std::vector<TreeNodeIndex> co(nLeaves + numInternal + 1, 0);
    std::vector<TreeNodeIndex> par(std::max(1, (nLeaves + numInternal - 1) / 8), 0);
    buildOctreeCpu(leaves, nLeaves, numInternal, prefixes, co.data(), par.data(), levelRange, internalToLeaf,
                   leafToInternal);
    std::copy_n(co.begin(), nLeaves + numInternal, childOffsets);
    std::copy_n(par.begin(), (nLeaves + numInternal - 1) / 8, parents);
}

template<class KeyType, class T>
void fpCenters(const KeyType* prefixes, size_t n, T* centers, T* sizes, const double* lim, const int* bnd)
{
    auto box = makeBox<T>(lim, bnd);
    nodeFpCenters<KeyType>(std::span<const KeyType>(prefixes, n), reinterpret_cast<Vec3<T>*>(centers),
                           reinterpret_cast<Vec3<T>*>(sizes), box);
}

//! OctreeNsView::searchExtFactor of the standalone neighbour searches below
static float g_searchExtFactor = 1.0f;
extern "C" void ref_set_search_ext_factor(float f) { g_searchExtFactor = f; }

template<class KeyType, class T, class Th = T>
void neighbors(const T* x,
               const T* y,
               const T* z,
               const Th* h,
               unsigned first,
               unsigned last,
               const double* lim,
               const int* bnd,
               int numLeaves,
               int numNodes,
               const KeyType* prefixes,
               const int* childOffsets,
               const int* parents,
               const int* internalToLeaf,
               const int* leafToInternal,
               const int* levelRange,
               const KeyType* leaves,
               const unsigned* layout,
               const T* centers,
               const T* sizes,
               unsigned ngmax,
               unsigned* nb,
               unsigned* nc)
{
    auto box = makeBox<T>(lim, bnd);
    OctreeNsView<T, KeyType> view{numLeaves,
                                  numNodes,
                                  prefixes,
                                  childOffsets,
                                  parents,
                                  internalToLeaf,
                                  leafToInternal,
                                  levelRange,
                                  leaves,
                                  layout,
                                  reinterpret_cast<const Vec3<T>*>(centers),
                                  reinterpret_cast<const Vec3<T>*>(sizes)};
    view.searchExtFactor = g_searchExtFactor; // tree/octree.hpp:279-282; 1 unless ref_set_search_ext_factor
    findNeighbors(x, y, z, h, first, last, box, view, ngmax, nb, nc);
}

template<class KeyType, class T>
void halos(const KeyType* prefixes,
           const int* childOffsets,
           const int* parents,
           const T* centers,
           const T* sizes,
           const KeyType* leaves,
           const T* searchCenters,
           const T* searchSizes,
           const double* lim,
           const int* bnd,
           int firstNode,
           int lastNode,
           uint8_t* flags)
{
    auto box = makeBox<T>(lim, bnd);
    findHalos(prefixes, childOffsets, parents, reinterpret_cast<const Vec3<T>*>(centers),
              reinterpret_cast<const Vec3<T>*>(sizes), leaves, reinterpret_cast<const Vec3<T>*>(searchCenters),
              reinterpret_cast<const Vec3<T>*>(searchSizes), box, firstNode, lastNode, flags);
}

template<class T, class Th = T>
void boundingBoxes(const T* x,
                   const T* y,
                   const T* z,
                   const Th* h,
                   const unsigned* layout,
                   int firstLeaf,
                   int lastLeaf,
                   Th scale,
                   T* searchCenters,
                   T* searchSizes)
{
    auto* c = reinterpret_cast<Vec3<T>*>(searchCenters);
    auto* s = reinterpret_cast<Vec3<T>*>(searchSizes);
#pragma omp parallel for schedule(static)
    for (int i = firstLeaf; i < lastLeaf; ++i)
    {
        std::tie(c[i], s[i]) = computeBoundingBox(x, y, z, h, layout[i], layout[i + 1], scale, c[i]);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Domain driver: P ranks as threads, each runs numSyncs x Domain::sync on its slice, optional findNeighbors.
 * ---------------------------------------------------------------------------------------------------------------- */

struct RankOut
{
    std::vector<uint64_t> keys; // widened
    std::vector<double> x, y, z, h;
    unsigned start{0}, end{0};
    std::vector<uint64_t> focusLeaves, globalLeaves;
    std::vector<unsigned> focusCounts, globalCounts, layout;
    std::vector<uint64_t> prefixes;
    std::vector<int> childOffsets, parents, internalToLeaf, leafToInternal, levelRange;
    std::vector<double> centers, sizes;
    std::vector<unsigned> nb, nc;
    std::vector<uint8_t> flags;
    double box[6];
    double tSync[8]{};
    double tNeighbors{0};
};

std::vector<RankOut> g_out;

//! halo search radius factor of the Domains built by the drivers below (Domain::setHaloFactor)
static float g_haloFactor = 1.0f;
extern "C" void ref_set_halo_factor(float f) { g_haloFactor = f; }

template<class KeyType, class T>
void domainRank(int rank,
                int P,
                unsigned bucket,
                unsigned bucketFocus,
                float theta,
                const double* lim,
                const int* bnd,
                const T* x0,
                const T* y0,
                const T* z0,
                const T* h0,
                size_t n,
                int numSyncs,
                unsigned ngmax,
                const T* moves)
{
    mpishim::rankRef() = rank;
    RankOut& o         = g_out[rank];

    Domain<KeyType, T> domain(execution::cpu, rank, P, bucket, bucketFocus, theta, MPI_COMM_WORLD,
                              makeBox<T>(lim, bnd));
    domain.setHaloFactor(g_haloFactor); // Domain::setHaloFactor (domain/domain.hpp:365); 1 unless ref_set_halo_factor
    std::vector<T> x(x0, x0 + n), y(y0, y0 + n), z(z0, z0 + n), h(h0, h0 + n);
    std::vector<KeyType> keys(n);
    std::vector<T> s1, s2;
    std::vector<LocalIndex> s3;

    for (int s = 0; s < numSyncs; ++s)
    {
        if (s > 0 && moves)
        {
            // deterministic per-sync drift of the assigned particles so that later syncs exchange particles
            T d = moves[s - 1];
            for (size_t i = domain.startIndex(); i < domain.endIndex(); ++i)
            {
                x[i] += d * (T(0.5) - T((keys[i] >> 3) & 7u) / T(7));
                y[i] += d * (T(0.5) - T((keys[i] >> 6) & 7u) / T(7));
                z[i] += d * (T(0.5) - T((keys[i] >> 9) & 7u) / T(7));
                auto b = domain.box();
                x[i]   = std::min(std::max(x[i], b.xmin()), std::nextafter(b.xmax(), b.xmin()));
                y[i]   = std::min(std::max(y[i], b.ymin()), std::nextafter(b.ymax(), b.ymin()));
                z[i]   = std::min(std::max(z[i], b.zmin()), std::nextafter(b.zmax(), b.zmin()));
            }
        }
        MPI_Barrier(MPI_COMM_WORLD);
        auto t0 = std::chrono::steady_clock::now();
        domain.sync(keys, x, y, z, h, std::tuple{}, std::tie(s1, s2, s3));
        auto t1 = std::chrono::steady_clock::now();
        if (s < 8) o.tSync[s] = std::chrono::duration<double>(t1 - t0).count();
    }

    o.start = domain.startIndex();
    o.end   = domain.endIndex();
    o.keys.assign(keys.begin(), keys.end());
    o.x.assign(x.begin(), x.end());
    o.y.assign(y.begin(), y.end());
    o.z.assign(z.begin(), z.end());
    o.h.assign(h.begin(), h.end());
    storeBox(domain.box(), o.box);

    auto fl = domain.focusTree().treeLeaves();
    o.focusLeaves.assign(fl.begin(), fl.end());
    auto fc = domain.focusTree().leafCountsAcc();
    o.focusCounts.assign(fc.begin(), fc.end());
    auto lay = domain.layout();
    o.layout.assign(lay.begin(), lay.end());
    auto gt = domain.globalTree();
    o.globalLeaves.assign(gt.leaves, gt.leaves + gt.numLeafNodes + 1);

    auto ft = domain.focusTree().octreeViewAcc();
    o.prefixes.assign(ft.prefixes, ft.prefixes + ft.numNodes);
    o.childOffsets.assign(ft.childOffsets, ft.childOffsets + ft.numNodes);
    o.parents.assign(ft.parents, ft.parents + (ft.numNodes - 1) / 8);
    o.internalToLeaf.assign(ft.internalToLeaf, ft.internalToLeaf + ft.numNodes);
    o.leafToInternal.assign(ft.leafToInternal, ft.leafToInternal + ft.numNodes);
    o.levelRange.assign(ft.levelRange, ft.levelRange + maxTreeLevel<KeyType>{} + 2);
    auto gc = domain.focusTree().geoCentersAcc();
    auto gs = domain.focusTree().geoSizesAcc();
    o.centers.resize(3 * ft.numNodes);
    o.sizes.resize(3 * ft.numNodes);
    for (int i = 0; i < ft.numNodes; ++i)
        for (int d = 0; d < 3; ++d)
        {
            o.centers[3 * i + d] = gc[i][d];
            o.sizes[3 * i + d]   = gs[i][d];
        }
    auto fg = domain.focusTree().flags();
    o.flags.assign(fg.begin(), fg.begin() + ft.numNodes);

    if (ngmax)
    {
        size_t nLoc = o.end - o.start;
        o.nb.resize(nLoc * ngmax);
        o.nc.resize(nLoc);
        auto t0 = std::chrono::steady_clock::now();
        findNeighbors(x.data(), y.data(), z.data(), h.data(), domain.startIndex(), domain.endIndex(), domain.box(),
                      domain.octreeProperties(), ngmax, o.nb.data(), o.nc.data());
        auto t1      = std::chrono::steady_clock::now();
        o.tNeighbors = std::chrono::duration<double>(t1 - t0).count();
    }
}

template<class KeyType, class T>
int domainRun(int P,
              unsigned bucket,
              unsigned bucketFocus,
              float theta,
              const double* lim,
              const int* bnd,
              const T* x,
              const T* y,
              const T* z,
              const T* h,
              const uint64_t* offsets,
              int numSyncs,
              unsigned ngmax,
              const T* moves)
{
    g_out.clear();
    g_out.resize(P);
    mpishim::world().reset(P);
    std::vector<std::thread> threads;
    std::vector<std::string> errors(P);
    for (int r = 0; r < P; ++r)
    {
        threads.emplace_back(
            [&, r]()
            {
                try
                {
                    domainRank<KeyType, T>(r, P, bucket, bucketFocus, theta, lim, bnd, x + offsets[r], y + offsets[r],
                                           z + offsets[r], h + offsets[r], offsets[r + 1] - offsets[r], numSyncs, ngmax,
                                           moves);
                }
                catch (std::exception& e)
                {
                    errors[r] = e.what();
                    fprintf(stderr, "ref domain rank %d: %s\n", r, e.what());
                    std::abort();
                }
            });
    }
    for (auto& t : threads)
        t.join();
    mpishim::world().reset(1);
    mpishim::rankRef() = 0;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Bench / full-size parity drivers: the same calls as domainRank, but nothing is copied out.  What comes back are the
 * wall times of the reference calls and position-weighted wrapping sums ("digests") of every result array, which the
 * GPU side recomputes from its own arrays: equal digests <=> equal arrays (up to 2^-64 collisions) at sizes where the
 * arrays themselves (40 GB of neighbour lists at 64 Mi particles) cannot be held twice.  findNeighbors runs in chunks
 * of `chunk` targets over one reused list buffer.
 * ---------------------------------------------------------------------------------------------------------------- */

constexpr uint64_t DIGEST_C1 = 0x9E3779B97F4A7C15ull, DIGEST_C2 = 0xC2B2AE3D27D4EB4Full;

template<class E>
uint64_t bitsOf(E v)
{
    if constexpr (sizeof(E) == 8)
    {
        uint64_t b;
        std::memcpy(&b, &v, 8);
        return b;
    }
    else if constexpr (sizeof(E) == 4)
    {
        uint32_t b;
        std::memcpy(&b, &v, 4);
        return b;
    }
    else
    {
        static_assert(sizeof(E) == 1);
        return uint64_t(uint8_t(v));
    }
}

//! sum_i bits(a[i]) * (i + 1) mod 2^64
template<class E>
uint64_t weightedSum(const E* a, size_t n)
{
    uint64_t s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i)
        s += bitsOf(a[i]) * uint64_t(i + 1);
    return s;
}

struct BenchOut
{
    double tSync[4]{};
    double tNeighbors{0};
    double tHalos{0};
    uint64_t digest[24]{};
};
std::vector<BenchOut> g_bench;

//! digest slots
enum
{
    DG_KEYS = 0, DG_X, DG_Y, DG_Z, DG_H, DG_LEAVES, DG_NUM_LEAVES, DG_LAYOUT, DG_NC_SUM, DG_LISTS, DG_START, DG_END,
    DG_SIZE, DG_NUM_NODES, DG_PREFIXES, DG_CHILD_OFFSETS, DG_CENTERS, DG_SIZES, DG_HALO_FLAGS, DG_HALO_COUNT,
    DG_LEAF_COUNTS
};

template<class T, class KeyType, class View>
void neighborsChunked(const T* x, const T* y, const T* z, const T* h, LocalIndex first, LocalIndex last,
                      const Box<T>& box, const View& view, unsigned ngmax, size_t chunk, double* seconds,
                      uint64_t* ncSum, uint64_t* listDigest)
{
    chunk = std::max<size_t>(1, std::min<size_t>(chunk, last - first));
    std::vector<LocalIndex> nb(chunk * ngmax);
    std::vector<unsigned> nc(chunk);
    double t      = 0;
    uint64_t sNc  = 0, sList = 0;
    for (size_t c0 = first; c0 < last; c0 += chunk)
    {
        size_t c1 = std::min<size_t>(c0 + chunk, last);
        auto t0   = std::chrono::steady_clock::now();
        findNeighbors(x, y, z, h, LocalIndex(c0), LocalIndex(c1), box, view, ngmax, nb.data(), nc.data());
        auto t1 = std::chrono::steady_clock::now();
        t += std::chrono::duration<double>(t1 - t0).count();
#pragma omp parallel for reduction(+ : sNc, sList) schedule(static)
        for (size_t i = c0; i < c1; ++i)
        {
            unsigned cnt = nc[i - c0];
            sNc += cnt;
            uint64_t row = uint64_t(i - first + 1) * DIGEST_C1;
            unsigned m   = std::min(cnt, ngmax);
            for (unsigned k = 0; k < m; ++k)
                sList += (uint64_t(nb[(i - c0) * ngmax + k]) + 1) * (row + uint64_t(k + 1) * DIGEST_C2);
        }
    }
    *seconds    = t;
    *ncSum      = sNc;
    *listDigest = sList;
}

template<class KeyType, class T>
void benchRank(int rank, int P, unsigned bucket, unsigned bucketFocus, float theta, const double* lim, const int* bnd,
               const T* x0, const T* y0, const T* z0, const T* h0, size_t n, int numSyncs, unsigned ngmax, size_t chunk,
               int numThreads, int haloQuarter)
{
    mpishim::rankRef() = rank;
    if (numThreads > 0) { omp_set_num_threads(numThreads); }
    BenchOut& o = g_bench[rank];

    Domain<KeyType, T> domain(execution::cpu, rank, P, bucket, bucketFocus, theta, MPI_COMM_WORLD,
                              makeBox<T>(lim, bnd));
    std::vector<T> x(x0, x0 + n), y(y0, y0 + n), z(z0, z0 + n), h(h0, h0 + n);
    std::vector<KeyType> keys(n);
    std::vector<T> s1, s2;
    std::vector<LocalIndex> s3;
    for (int s = 0; s < numSyncs; ++s)
    {
        MPI_Barrier(MPI_COMM_WORLD);
        auto t0 = std::chrono::steady_clock::now();
        domain.sync(keys, x, y, z, h, std::tuple{}, std::tie(s1, s2, s3));
        auto t1 = std::chrono::steady_clock::now();
        if (s < 4) o.tSync[s] = std::chrono::duration<double>(t1 - t0).count();
    }
    uint64_t* d      = o.digest;
    d[DG_KEYS]       = weightedSum(keys.data(), keys.size());
    d[DG_X]          = weightedSum(x.data(), x.size());
    d[DG_Y]          = weightedSum(y.data(), y.size());
    d[DG_Z]          = weightedSum(z.data(), z.size());
    d[DG_H]          = weightedSum(h.data(), h.size());
    auto fl          = domain.focusTree().treeLeaves();
    d[DG_LEAVES]     = weightedSum(fl.data(), fl.size());
    d[DG_NUM_LEAVES] = fl.size() - 1;
    auto lay         = domain.layout();
    d[DG_LAYOUT]     = weightedSum(lay.data(), lay.size());
    d[DG_START]      = domain.startIndex();
    d[DG_END]        = domain.endIndex();
    d[DG_SIZE]       = x.size();
    auto ft          = domain.focusTree().octreeViewAcc();
    d[DG_NUM_NODES]  = ft.numNodes;
    d[DG_PREFIXES]   = weightedSum(ft.prefixes, ft.numNodes);
    d[DG_CHILD_OFFSETS] = weightedSum(ft.childOffsets, ft.numNodes);
    auto gc             = domain.focusTree().geoCentersAcc();
    auto gs             = domain.focusTree().geoSizesAcc();
    d[DG_CENTERS]       = weightedSum(reinterpret_cast<const T*>(gc.data()), size_t(3) * ft.numNodes);
    d[DG_SIZES]         = weightedSum(reinterpret_cast<const T*>(gs.data()), size_t(3) * ft.numNodes);
    auto lc             = domain.focusTree().leafCountsAcc();
    d[DG_LEAF_COUNTS]   = weightedSum(lc.data(), lc.size());
    if (haloQuarter)
    {
        // standalone halo discovery like test/performance/octree.cu:110-143: the first quarter of the leaves plays
        // the assigned range (a single rank has no foreign leaves of its own)
        TreeNodeIndex nLeaves = TreeNodeIndex(fl.size()) - 1, firstLeaf = 0, lastLeaf = nLeaves / 4;
        std::vector<Vec3<T>> sc(nLeaves), ss(nLeaves);
        std::vector<uint8_t> flags(ft.numNodes, 0);
        auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(static)
        for (TreeNodeIndex i = firstLeaf; i < lastLeaf; ++i)
        {
            Vec3<T> init = gc[ft.leafToInternal[ft.numInternalNodes + i]];
            std::tie(sc[i], ss[i]) =
                computeBoundingBox(x.data(), y.data(), z.data(), h.data(), lay[i], lay[i + 1], T(2), init);
        }
        findHalos(ft.prefixes, ft.childOffsets, ft.parents, gc.data(), gs.data(), fl.data(), sc.data(), ss.data(),
                  domain.box(), firstLeaf, lastLeaf, flags.data());
        auto t1           = std::chrono::steady_clock::now();
        o.tHalos          = std::chrono::duration<double>(t1 - t0).count();
        d[DG_HALO_FLAGS]  = weightedSum(flags.data(), flags.size());
        d[DG_HALO_COUNT]  = std::accumulate(flags.begin(), flags.end(), uint64_t(0));
    }
    if (ngmax)
    {
        neighborsChunked<T, KeyType>(x.data(), y.data(), z.data(), h.data(), domain.startIndex(), domain.endIndex(),
                                     domain.box(), domain.octreeProperties(), ngmax, chunk, &o.tNeighbors,
                                     &d[DG_NC_SUM], &d[DG_LISTS]);
    }
}

template<class KeyType, class T>
int benchRun(int P, unsigned bucket, unsigned bucketFocus, float theta, const double* lim, const int* bnd, const T* x,
             const T* y, const T* z, const T* h, const uint64_t* offsets, int numSyncs, unsigned ngmax, size_t chunk,
             int numThreads, int haloQuarter)
{
    g_bench.clear();
    g_bench.resize(P);
    mpishim::world().reset(P);
    std::vector<std::thread> threads;
    for (int r = 0; r < P; ++r)
    {
        threads.emplace_back(
            [&, r]()
            {
                try
                {
                    benchRank<KeyType, T>(r, P, bucket, bucketFocus, theta, lim, bnd, x + offsets[r], y + offsets[r],
                                          z + offsets[r], h + offsets[r], offsets[r + 1] - offsets[r], numSyncs, ngmax,
                                          chunk, numThreads, haloQuarter);
                }
                catch (std::exception& e)
                {
                    fprintf(stderr, "ref bench rank %d: %s\n", r, e.what());
                    std::abort();
                }
            });
    }
    for (auto& t : threads)
        t.join();
    mpishim::world().reset(1);
    mpishim::rankRef() = 0;
    return 0;
}

/*! standalone tree build + neighbour search (BASELINE configs[0] / configs[4]): keys of either curve, sort_by_key,
 *  gather, computeOctree, buildOctreeCpu, node centres decoded with the same curve, findNeighbors in chunks.
 *  times: [0] keys+sort+gather+tree+link+centres, [1] findNeighbors */
template<class KeyType, class T>
int benchTreeNeighbors(int kind, const T* x0, const T* y0, const T* z0, const T* h0, size_t n, unsigned bucket,
                       const double* lim, const int* bnd, unsigned ngmax, size_t chunk, double* times, uint64_t* d)
{
    auto box = makeBox<T>(lim, bnd);
    std::vector<T> x(x0, x0 + n), y(y0, y0 + n), z(z0, z0 + n), h(h0, h0 + n);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<KeyType> keys(n);
    sfcKeys<KeyType, T>(kind, x.data(), y.data(), z.data(), keys.data(), n, lim, bnd);
    std::vector<LocalIndex> order(n);
    std::iota(order.begin(), order.end(), LocalIndex(0));
    sort_by_key(keys.begin(), keys.end(), order.begin());
    {
        std::vector<T> tmp(n);
        gatherArrays(execution::cpu, order, 0, std::tie(x, y, z, h), std::tie(tmp));
    }
    auto [leaves, counts] = computeOctree<KeyType>(std::span<const KeyType>(keys), bucket);
    int nLeaves           = int(nNodes(leaves));
    int numInternal       = (nLeaves - 1) / 7;
    int numNodes          = nLeaves + numInternal;
    std::vector<KeyType> prefixes(numNodes);
    std::vector<TreeNodeIndex> co(numNodes + 1, 0), par(std::max(1, (numNodes - 1) / 8), 0), i2l(numNodes),
        l2i(numNodes), levelRange(maxTreeLevel<KeyType>{} + 2);
    buildOctreeCpu(leaves.data(), nLeaves, numInternal, prefixes.data(), co.data(), par.data(), levelRange.data(),
                   i2l.data(), l2i.data());
    std::vector<Vec3<T>> centers(numNodes), sizes(numNodes);
    if (kind == 0) { nodeFpCenters<KeyType>(std::span<const KeyType>(prefixes), centers.data(), sizes.data(), box); }
    else
    {
        // nodeFpCenters decodes node boxes as Hilbert keys whatever curve produced the tree; for a
        // Morton tree the same two reference functions are applied with the Morton decode
#pragma omp parallel for schedule(static)
        for (int i = 0; i < numNodes; ++i)
        {
            KeyType startKey                = decodePlaceholderBit(prefixes[i]);
            unsigned level                  = decodePrefixLength(prefixes[i]) / 3;
            auto nodeBox                    = sfcIBox(MortonKey<KeyType>(startKey), level);
            util::tie(centers[i], sizes[i]) = centerAndSize<KeyType>(nodeBox, box);
        }
    }
    std::vector<LocalIndex> layout(nLeaves + 1, 0);
    std::exclusive_scan(counts.begin(), counts.end(), layout.begin(), LocalIndex(0));
    layout[nLeaves] = LocalIndex(n);
    auto t1         = std::chrono::steady_clock::now();
    times[0]        = std::chrono::duration<double>(t1 - t0).count();

    OctreeNsView<T, KeyType> view{nLeaves,      numNodes,   prefixes.data(),   co.data(),     par.data(),
                                  i2l.data(),   l2i.data(), levelRange.data(), leaves.data(), layout.data(),
                                  centers.data(), sizes.data()};
    d[DG_KEYS]       = weightedSum(keys.data(), n);
    d[DG_X]          = weightedSum(x.data(), n);
    d[DG_Y]          = weightedSum(y.data(), n);
    d[DG_Z]          = weightedSum(z.data(), n);
    d[DG_H]          = weightedSum(h.data(), n);
    d[DG_LEAVES]     = weightedSum(leaves.data(), leaves.size());
    d[DG_NUM_LEAVES] = nLeaves;
    d[DG_LAYOUT]     = weightedSum(layout.data(), layout.size());
    d[DG_START]      = 0;
    d[DG_END]        = n;
    d[DG_SIZE]       = n;
    d[DG_NUM_NODES]  = numNodes;
    d[DG_PREFIXES]   = weightedSum(prefixes.data(), numNodes);
    d[DG_CHILD_OFFSETS] = weightedSum(co.data(), numNodes);
    d[DG_CENTERS]       = weightedSum(reinterpret_cast<const T*>(centers.data()), size_t(3) * numNodes);
    d[DG_SIZES]         = weightedSum(reinterpret_cast<const T*>(sizes.data()), size_t(3) * numNodes);
    d[DG_LEAF_COUNTS]   = weightedSum(counts.data(), counts.size());
    neighborsChunked<T, KeyType>(x.data(), y.data(), z.data(), h.data(), 0, LocalIndex(n), box, view, ngmax, chunk,
                                 &times[1], &d[DG_NC_SUM], &d[DG_LISTS]);
    return 0;
}

template<class V, class O>
long fetch(const V& v, O* out, long cap)
{
    if (out == nullptr) return long(v.size());
    long n = std::min<long>(cap, long(v.size()));
    for (long i = 0; i < n; ++i)
        out[i] = O(v[i]);
    return long(v.size());
}

} // namespace

#define CS_INST_KT(SUFFIX, KeyType, T)                                                                                 \
    extern "C" void ref_sfc_keys_##SUFFIX(int kind, const T* x, const T* y, const T* z, KeyType* keys, size_t n,       \
                                          const double* lim, const int* bnd)                                           \
    {                                                                                                                  \
        sfcKeys<KeyType, T>(kind, x, y, z, keys, n, lim, bnd);                                                         \
    }                                                                                                                  \
    extern "C" void ref_node_fp_centers_##SUFFIX(const KeyType* prefixes, size_t n, T* centers, T* sizes,              \
                                                 const double* lim, const int* bnd)                                    \
    {                                                                                                                  \
        fpCenters<KeyType, T>(prefixes, n, centers, sizes, lim, bnd);                                                  \
    }                                                                                                                  \
    extern "C" void ref_find_neighbors_##SUFFIX(                                                                       \
        const T* x, const T* y, const T* z, const T* h, unsigned first, unsigned last, const double* lim,              \
        const int* bnd, int numLeaves, int numNodes, const KeyType* prefixes, const int* childOffsets,                 \
        const int* parents, const int* internalToLeaf, const int* leafToInternal, const int* levelRange,               \
        const KeyType* leaves, const unsigned* layout, const T* centers, const T* sizes, unsigned ngmax, unsigned* nb, \
        unsigned* nc)                                                                                                  \
    {                                                                                                                  \
        neighbors<KeyType, T>(x, y, z, h, first, last, lim, bnd, numLeaves, numNodes, prefixes, childOffsets, parents, \
                              internalToLeaf, leafToInternal, levelRange, leaves, layout, centers, sizes, ngmax, nb,   \
                              nc);                                                                                     \
    }                                                                                                                  \
    extern "C" void ref_find_halos_##SUFFIX(const KeyType* prefixes, const int* childOffsets, const int* parents,      \
                                            const T* centers, const T* sizes, const KeyType* leaves,                   \
                                            const T* searchCenters, const T* searchSizes, const double* lim,           \
                                            const int* bnd, int firstNode, int lastNode, uint8_t* flags)               \
    {                                                                                                                  \
        halos<KeyType, T>(prefixes, childOffsets, parents, centers, sizes, leaves, searchCenters, searchSizes, lim,    \
                          bnd, firstNode, lastNode, flags);                                                            \
    }                                                                                                                  \
    extern "C" int ref_domain_run_##SUFFIX(int P, unsigned bucket, unsigned bucketFocus, float theta,                  \
                                           const double* lim, const int* bnd, const T* x, const T* y, const T* z,      \
                                           const T* h, const uint64_t* offsets, int numSyncs, unsigned ngmax,          \
                                           const T* moves)                                                             \
    {                                                                                                                  \
        return domainRun<KeyType, T>(P, bucket, bucketFocus, theta, lim, bnd, x, y, z, h, offsets, numSyncs, ngmax,    \
                                     moves);                                                                           \
    }

extern "C" void ref_sfc_keys_u32d(int kind, const double* x, const double* y, const double* z, uint32_t* keys, size_t n,
                                  const double* lim, const int* bnd)
{
    sfcKeys<uint32_t, double>(kind, x, y, z, keys, n, lim, bnd);
}

//! mixed precision (H4): double coordinates, float smoothing lengths
extern "C" void ref_find_neighbors_u64df(const double* x, const double* y, const double* z, const float* h,
                                         unsigned first, unsigned last, const double* lim, const int* bnd,
                                         int numLeaves, int numNodes, const uint64_t* prefixes, const int* childOffsets,
                                         const int* parents, const int* internalToLeaf, const int* leafToInternal,
                                         const int* levelRange, const uint64_t* leaves, const unsigned* layout,
                                         const double* centers, const double* sizes, unsigned ngmax, unsigned* nb,
                                         unsigned* nc)
{
    neighbors<uint64_t, double, float>(x, y, z, h, first, last, lim, bnd, numLeaves, numNodes, prefixes, childOffsets,
                                       parents, internalToLeaf, leafToInternal, levelRange, leaves, layout, centers,
                                       sizes, ngmax, nb, nc);
}

CS_INST_KT(u32f, uint32_t, float)
CS_INST_KT(u64f, uint64_t, float)
CS_INST_KT(u64d, uint64_t, double)

#define CS_INST_BENCH(SUFFIX, KeyType, T)                                                                              \
    extern "C" int ref_bench_run_##SUFFIX(int P, unsigned bucket, unsigned bucketFocus, float theta,                   \
                                          const double* lim, const int* bnd, const T* x, const T* y, const T* z,       \
                                          const T* h, const uint64_t* offsets, int numSyncs, unsigned ngmax,           \
                                          size_t chunk, int numThreads, int haloQuarter)                               \
    {                                                                                                                  \
        return benchRun<KeyType, T>(P, bucket, bucketFocus, theta, lim, bnd, x, y, z, h, offsets, numSyncs, ngmax,     \
                                    chunk, numThreads, haloQuarter);                                                   \
    }                                                                                                                  \
    extern "C" int ref_bench_tree_neighbors_##SUFFIX(int kind, const T* x, const T* y, const T* z, const T* h,         \
                                                     size_t n, unsigned bucket, const double* lim, const int* bnd,     \
                                                     unsigned ngmax, size_t chunk, double* times, uint64_t* digest)    \
    {                                                                                                                  \
        return benchTreeNeighbors<KeyType, T>(kind, x, y, z, h, n, bucket, lim, bnd, ngmax, chunk, times, digest);     \
    }

CS_INST_BENCH(u32f, uint32_t, float)
CS_INST_BENCH(u64f, uint64_t, float)
CS_INST_BENCH(u64d, uint64_t, double)

//! results of the last ref_bench_run_*: times = {sync[0..3], neighbors, halo discovery}, digest[24]
extern "C" void ref_bench_get(int rank, double* times6, uint64_t* digest24)
{
    std::memcpy(times6, g_bench[rank].tSync, sizeof(double) * 4);
    times6[4] = g_bench[rank].tNeighbors;
    times6[5] = g_bench[rank].tHalos;
    std::memcpy(digest24, g_bench[rank].digest, sizeof(uint64_t) * 24);
}

//! the reference's Plummer sphere (test/coord_samples/plummer.hpp:15-78, srand48(42))
extern "C" void ref_plummer_d(size_t n, double* x, double* y, double* z)
{
    auto pos = plummer<double>(n);
    std::copy(pos[0].begin(), pos[0].end(), x);
    std::copy(pos[1].begin(), pos[1].end(), y);
    std::copy(pos[2].begin(), pos[2].end(), z);
}
extern "C" void ref_plummer_f(size_t n, float* x, float* y, float* z)
{
    auto pos = plummer<float>(n);
    std::copy(pos[0].begin(), pos[0].end(), x);
    std::copy(pos[1].begin(), pos[1].end(), y);
    std::copy(pos[2].begin(), pos[2].end(), z);
}

#define CS_INST_K(SUFFIX, KeyType)                                                                                     \
    extern "C" void ref_sort_by_key_##SUFFIX(KeyType* keys, unsigned* values, size_t n)                                \
    {                                                                                                                  \
        sort_by_key(keys, keys + n, values);                                                                           \
    }                                                                                                                  \
    extern "C" long ref_compute_octree_##SUFFIX(const KeyType* keys, size_t n, unsigned bucket, KeyType* leaves,       \
                                                unsigned* counts, long cap)                                            \
    {                                                                                                                  \
        return computeTree<KeyType>(keys, n, bucket, leaves, counts, cap);                                             \
    }                                                                                                                  \
    extern "C" long ref_update_octree_##SUFFIX(const KeyType* keys, size_t n, unsigned bucket, KeyType* leaves,        \
                                               unsigned* counts, long nLeaves, long cap, int* converged)               \
    {                                                                                                                  \
        return updateTree<KeyType>(keys, n, bucket, leaves, counts, nLeaves, cap, converged);                          \
    }                                                                                                                  \
    extern "C" void ref_compute_node_counts_##SUFFIX(const KeyType* leaves, unsigned* counts, int nLeaves,             \
                                                     const KeyType* keys, size_t n, unsigned maxCount)                 \
    {                                                                                                                  \
        computeNodeCounts<KeyType>(leaves, counts, nLeaves, std::span<const KeyType>(keys, n), maxCount, false);       \
    }                                                                                                                  \
    extern "C" int ref_rebalance_decision_##SUFFIX(const KeyType* leaves, const unsigned* counts, int nLeaves,         \
                                                   unsigned bucket, int* ops)                                          \
    {                                                                                                                  \
        return rebalanceDecision(leaves, counts, nLeaves, bucket, ops);                                                \
    }                                                                                                                  \
    extern "C" void ref_build_octree_##SUFFIX(const KeyType* leaves, int nLeaves, KeyType* prefixes,                   \
                                              int* childOffsets, int* parents, int* levelRange, int* internalToLeaf,   \
                                              int* leafToInternal)                                                     \
    {                                                                                                                  \
        linkTree<KeyType>(leaves, nLeaves, prefixes, childOffsets, parents, levelRange, internalToLeaf,                \
                          leafToInternal);                                                                             \
    }

CS_INST_K(u32, uint32_t)
CS_INST_K(u64, uint64_t)

extern "C" void ref_bounding_boxes_f(const float* x, const float* y, const float* z, const float* h,
                                     const unsigned* layout, int firstLeaf, int lastLeaf, float scale, float* sc,
                                     float* ss)
{
    boundingBoxes<float>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, sc, ss);
}
extern "C" void ref_bounding_boxes_df(const double* x, const double* y, const double* z, const float* h,
                                      const unsigned* layout, int firstLeaf, int lastLeaf, float scale, double* sc,
                                      double* ss)
{
    boundingBoxes<double, float>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, sc, ss);
}
extern "C" void ref_bounding_boxes_d(const double* x, const double* y, const double* z, const double* h,
                                     const unsigned* layout, int firstLeaf, int lastLeaf, double scale, double* sc,
                                     double* ss)
{
    boundingBoxes<double>(x, y, z, h, layout, firstLeaf, lastLeaf, scale, sc, ss);
}

/* ---- accessors for the results of the last ref_domain_run_* (out == NULL returns the length) ---- */
#define CS_FETCH(NAME, FIELD, OT)                                                                                      \
    extern "C" long ref_domain_get_##NAME(int rank, OT* out, long cap) { return fetch(g_out[rank].FIELD, out, cap); }

CS_FETCH(keys, keys, uint64_t)
CS_FETCH(x, x, double)
CS_FETCH(y, y, double)
CS_FETCH(z, z, double)
CS_FETCH(h, h, double)
CS_FETCH(focus_leaves, focusLeaves, uint64_t)
CS_FETCH(global_leaves, globalLeaves, uint64_t)
CS_FETCH(focus_counts, focusCounts, unsigned)
CS_FETCH(layout, layout, unsigned)
CS_FETCH(prefixes, prefixes, uint64_t)
CS_FETCH(child_offsets, childOffsets, int)
CS_FETCH(parents, parents, int)
CS_FETCH(internal_to_leaf, internalToLeaf, int)
CS_FETCH(leaf_to_internal, leafToInternal, int)
CS_FETCH(level_range, levelRange, int)
CS_FETCH(centers, centers, double)
CS_FETCH(sizes, sizes, double)
CS_FETCH(neighbors, nb, unsigned)
CS_FETCH(neighbors_count, nc, unsigned)
CS_FETCH(flags, flags, uint8_t)

extern "C" void ref_domain_get_info(int rank, unsigned* startEnd, double* box, double* tSync, double* tNeighbors)
{
    startEnd[0] = g_out[rank].start;
    startEnd[1] = g_out[rank].end;
    std::memcpy(box, g_out[rank].box, sizeof(double) * 6);
    std::memcpy(tSync, g_out[rank].tSync, sizeof(double) * 8);
    *tNeighbors = g_out[rank].tNeighbors;
}

extern "C" int ref_num_threads() { return omp_get_max_threads(); }
