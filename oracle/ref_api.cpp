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

template<class KeyType, class T>
void neighbors(const T* x,
               const T* y,
               const T* z,
               const T* h,
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

template<class T>
void boundingBoxes(const T* x,
                   const T* y,
                   const T* z,
                   const T* h,
                   const unsigned* layout,
                   int firstLeaf,
                   int lastLeaf,
                   T scale,
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

CS_INST_KT(u32f, uint32_t, float)
CS_INST_KT(u64f, uint64_t, float)
CS_INST_KT(u64d, uint64_t, double)

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
