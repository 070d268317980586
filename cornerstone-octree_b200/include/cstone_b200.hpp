/* C++20 host-side mirror of the reference's GPU interface for the domain-sync hot path, implemented as thin forwarders
 * to the C ABI (include/cstone_b200.h).  Names, argument order and error behaviour follow the reference
 * (sekelle/cornerstone-octree, include/cstone/...); file:line citations next to each function.
 *
 * The forwarders are templates over the reference's own vocabulary types, so this header needs none of the reference's
 * headers: `Exec` is anything convertible to cudaStream_t (cstone::execution::Gpu, execution.hpp:37-56), `BoxT` is
 * anything with xmin()..zmax() and boundaryX/Y/Z() (cstone::Box<T>, sfc/box.hpp:86-174), key pointers may be plain
 * integers or the reference's strong types (sfc/sfc.hpp:30-40; same size and layout).
 */
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

#include "cstone_b200.h"

namespace cstone_b200
{

/*! status -> the reference's two error behaviours: CUDA failures print and exit (cuda/errorcheck.cuh:15-27),
 *  contract violations throw std::runtime_error (primitives/primitives_gpu.cu:338, domain/domain.hpp:81-85) */
inline void csCheck(int status, const char* what)
{
    if (status == 0) { return; }
    if (status == 1)
    {
        std::fprintf(stderr, "%s: CUDA error: %s\n", what, cs_last_error());
        std::exit(EXIT_FAILURE);
    }
    throw std::runtime_error(std::string(what) + ": " + cs_last_error());
}

template<class BoxT>
struct BoxArgs
{
    double lim[6];
    int bnd[3];
    explicit BoxArgs(const BoxT& b)
        : lim{double(b.xmin()), double(b.xmax()), double(b.ymin()), double(b.ymax()), double(b.zmin()), double(b.zmax())}
        , bnd{int(b.boundaryX()), int(b.boundaryY()), int(b.boundaryZ())}
    {
    }
};

template<class Exec>
inline void* streamOf(const Exec& exec)
{
    if constexpr (std::is_pointer_v<Exec> || std::is_null_pointer_v<Exec>) { return (void*)exec; }
    else { return (void*)(exec.stream()); }
}

//! @brief sfc kind of a key type: the reference's MortonKey<> strong type carries a nested tag; plain integers = Hilbert
enum class SfcKind : int
{
    hilbert = 0,
    morton  = 1
};

/* computeSfcKeys(Gpu, x, y, z, keys, n, box)   sfc/sfc_gpu.h:24-26 */
template<class Exec, class T, class KeyType, class BoxT>
void computeSfcKeys(
    const Exec& exec, const T* x, const T* y, const T* z, KeyType* keys, size_t n, const BoxT& box, SfcKind kind = SfcKind::hilbert)
{
    BoxArgs<BoxT> b(box);
    static_assert(sizeof(KeyType) == 4 || sizeof(KeyType) == 8);
    if constexpr (sizeof(KeyType) == 4 && std::is_same_v<T, float>)
    {
        csCheck(cs_compute_sfc_keys_u32f(int(kind), x, y, z, reinterpret_cast<uint32_t*>(keys), n, b.lim, b.bnd, streamOf(exec)),
                "computeSfcKeys");
    }
    else if constexpr (sizeof(KeyType) == 8 && std::is_same_v<T, float>)
    {
        csCheck(cs_compute_sfc_keys_u64f(int(kind), x, y, z, reinterpret_cast<uint64_t*>(keys), n, b.lim, b.bnd, streamOf(exec)),
                "computeSfcKeys");
    }
    else
    {
        static_assert(sizeof(KeyType) == 8 && std::is_same_v<T, double>, "unsupported (KeyType, T) combination");
        csCheck(cs_compute_sfc_keys_u64d(int(kind), x, y, z, reinterpret_cast<uint64_t*>(keys), n, b.lim, b.bnd, streamOf(exec)),
                "computeSfcKeys");
    }
}

/* sortByKeyTempStorage<K,V>(n), sortByKey(Gpu, first, last, values, keyBuf, valueBuf, tmp, tmpBytes)
 * primitives/primitives_gpu.h:115-130 */
template<class KeyType>
uint64_t sortByKeyTempStorage(uint64_t n)
{
    if constexpr (sizeof(KeyType) == 8) { return cs_sort_by_key_temp_bytes_u64(n); }
    else { return cs_sort_by_key_temp_bytes_u32(n); }
}

template<class Exec, class KeyType, class ValueType>
void sortByKey(const Exec& exec,
               KeyType* first,
               KeyType* last,
               ValueType* values,
               KeyType* keyBuf,
               ValueType* valueBuf,
               void* tmp,
               uint64_t tmpBytes)
{
    static_assert(sizeof(ValueType) == 4, "values are LocalIndex / TreeNodeIndex");
    size_t n = size_t(last - first);
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_sort_by_key_u64(reinterpret_cast<uint64_t*>(first), reinterpret_cast<uint32_t*>(values), n,
                                   reinterpret_cast<uint64_t*>(keyBuf), reinterpret_cast<uint32_t*>(valueBuf), tmp,
                                   tmpBytes, streamOf(exec)),
                "sortByKey");
    }
    else
    {
        csCheck(cs_sort_by_key_u32(reinterpret_cast<uint32_t*>(first), reinterpret_cast<uint32_t*>(values), n,
                                   reinterpret_cast<uint32_t*>(keyBuf), reinterpret_cast<uint32_t*>(valueBuf), tmp,
                                   tmpBytes, streamOf(exec)),
                "sortByKey");
    }
}

/* gather(Gpu, ordering, src, dst)   primitives/primitives_gpu.h:30-42 */
template<class Exec, class T>
void gather(const Exec& exec, const uint32_t* ordering, size_t n, const T* src, T* dst)
{
    static_assert(sizeof(T) == 4 || sizeof(T) == 8);
    csCheck(cs_gather(ordering, n, src, dst, int(sizeof(T)), streamOf(exec)), "gather");
}

/* computeNodeCountsGpu(tree, counts, numNodes, keys, maxCount, useCountsAsGuess)   tree/csarray_gpu.h:41-52 */
template<class Exec, class KeyType>
void computeNodeCountsGpu(const Exec& exec,
                          const KeyType* tree,
                          unsigned* counts,
                          int numNodes,
                          const KeyType* keysBegin,
                          const KeyType* keysEnd,
                          unsigned maxCount,
                          bool /*useCountsAsGuess: same result either way*/ = false)
{
    size_t n = size_t(keysEnd - keysBegin);
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_compute_node_counts_u64(reinterpret_cast<const uint64_t*>(tree), counts, numNodes,
                                           reinterpret_cast<const uint64_t*>(keysBegin), n, maxCount, streamOf(exec)),
                "computeNodeCountsGpu");
    }
    else
    {
        csCheck(cs_compute_node_counts_u32(reinterpret_cast<const uint32_t*>(tree), counts, numNodes,
                                           reinterpret_cast<const uint32_t*>(keysBegin), n, maxCount, streamOf(exec)),
                "computeNodeCountsGpu");
    }
}

/* buildOctreeGpu(cstoneTree, OctreeView, keyBuf, valueBuf, tmp)   tree/octree_gpu.h:35-52.
 * OctreeViewT: tree/octree.hpp:234-256 */
template<class Exec, class KeyType, class OctreeViewT>
void buildOctreeGpu(const Exec& exec, const KeyType* cstoneTree, OctreeViewT o, void* tmp, size_t tmpBytes)
{
    if constexpr (sizeof(KeyType) == 8)
    {
        csCheck(cs_build_octree_u64(reinterpret_cast<const uint64_t*>(cstoneTree), o.numLeafNodes,
                                    reinterpret_cast<uint64_t*>(o.prefixes), o.childOffsets, o.parents, o.d_levelRange,
                                    o.internalToLeaf, o.leafToInternal, tmp, tmpBytes, streamOf(exec)),
                "buildOctreeGpu");
    }
    else
    {
        csCheck(cs_build_octree_u32(reinterpret_cast<const uint32_t*>(cstoneTree), o.numLeafNodes,
                                    reinterpret_cast<uint32_t*>(o.prefixes), o.childOffsets, o.parents, o.d_levelRange,
                                    o.internalToLeaf, o.leafToInternal, tmp, tmpBytes, streamOf(exec)),
                "buildOctreeGpu");
    }
}

/* findNeighbors(x, y, z, h, firstId, lastId, box, OctreeNsView, ngmax, neighbors, neighborsCount)
 * findneighbors.hpp:156-177, OctreeNsView: tree/octree.hpp:259-283 */
template<class Exec, class T, class BoxT, class NsViewT>
void findNeighborsGpu(const Exec& exec,
                      const T* x,
                      const T* y,
                      const T* z,
                      const T* h,
                      uint32_t firstId,
                      uint32_t lastId,
                      const BoxT& box,
                      const NsViewT& tree,
                      unsigned ngmax,
                      uint32_t* neighbors,
                      unsigned* neighborsCount)
{
    BoxArgs<BoxT> b(box);
    auto* centers = reinterpret_cast<const T*>(tree.centers);
    auto* sizes   = reinterpret_cast<const T*>(tree.sizes);
    if constexpr (std::is_same_v<T, double>)
    {
        csCheck(cs_find_neighbors_d(x, y, z, h, firstId, lastId, b.lim, b.bnd, tree.numLeafNodes, tree.childOffsets,
                                    tree.parents, tree.internalToLeaf, tree.layout, centers, sizes, ngmax, neighbors,
                                    neighborsCount, streamOf(exec)),
                "findNeighbors");
    }
    else
    {
        csCheck(cs_find_neighbors_f(x, y, z, h, firstId, lastId, b.lim, b.bnd, tree.numLeafNodes, tree.childOffsets,
                                    tree.parents, tree.internalToLeaf, tree.layout, centers, sizes, ngmax, neighbors, neighborsCount,
                                    streamOf(exec)),
                "findNeighbors");
    }
}

/*! Domain<KeyType, T, Gpu> (domain/domain.hpp:38-664) over cs_domain_*.  The domain owns the particle arrays in HBM;
 *  sync() takes the caller's arrays (device pointers) and the accessors return the synchronised ones. */
template<class KeyType, class T>
class Domain
{
public:
    template<class BoxT>
    Domain(int rank, int nRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta, const BoxT& box)
    {
        BoxArgs<BoxT> b(box);
        static_assert(sizeof(KeyType) == 8 || std::is_same_v<T, float>,
                      "32-bit keys come with float coordinates only (the instantiations of the reference's Domain)");
        if constexpr (sizeof(KeyType) == 4) { d_ = cs_domain_create_u32f(rank, nRanks, bucketSize, bucketSizeFocus, theta, b.lim, b.bnd); }
        else if constexpr (std::is_same_v<T, float>) { d_ = cs_domain_create_u64f(rank, nRanks, bucketSize, bucketSizeFocus, theta, b.lim, b.bnd); }
        else { d_ = cs_domain_create_u64d(rank, nRanks, bucketSize, bucketSizeFocus, theta, b.lim, b.bnd); }
        if (!d_) { throw std::runtime_error(cs_last_error()); }
    }
    /*! multi-rank domain: the communicator takes the place of the reference's MPI_Comm argument (domain.hpp:63-86).
     *  Create it with cs_comm_create_nccl (one process per GPU; rank 0 distributes cs_nccl_unique_id) or
     *  cs_comm_create_local (ranks as threads); it must outlive the domain. */
    template<class BoxT>
    Domain(int rank, int nRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta, cs_comm_t* comm,
           const BoxT& box)
        : Domain(rank, nRanks, bucketSize, bucketSizeFocus, theta, box)
    {
        if (comm) { csCheck(cs_domain_attach_comm(d_, comm), "Domain: attach communicator"); }
    }
    Domain(const Domain&)            = delete;
    Domain& operator=(const Domain&) = delete;
    ~Domain() { cs_domain_destroy(d_); }

    //! sync(keys, x, y, z, h, ...) domain.hpp:169-218; pass nullptr arrays to re-sync the domain-owned arrays in place
    void sync(const T* x, const T* y, const T* z, const T* h, const KeyType* keys, size_t n, void* stream = nullptr)
    {
        csCheck(cs_domain_sync(d_, x, y, z, h, keys, n, 0, stream), "Domain::sync");
        cs_domain_info(d_, info_, box_);
    }

    uint32_t startIndex() const { return uint32_t(info_[0]); }
    uint32_t endIndex() const { return uint32_t(info_[1]); }
    uint32_t nParticles() const { return endIndex() - startIndex(); }
    uint32_t nParticlesWithHalos() const { return uint32_t(info_[2]); }
    const double* boxLimits() const { return box_; }

    T* x() { return static_cast<T*>(cs_domain_ptr(d_, CS_FIELD_X)); }
    T* y() { return static_cast<T*>(cs_domain_ptr(d_, CS_FIELD_Y)); }
    T* z() { return static_cast<T*>(cs_domain_ptr(d_, CS_FIELD_Z)); }
    T* h() { return static_cast<T*>(cs_domain_ptr(d_, CS_FIELD_H)); }
    KeyType* keys() { return static_cast<KeyType*>(cs_domain_ptr(d_, CS_FIELD_KEYS)); }

    void findNeighbors(unsigned ngmax, uint32_t* neighbors, unsigned* neighborsCount, void* stream = nullptr)
    {
        csCheck(cs_domain_find_neighbors(d_, ngmax, neighbors, neighborsCount, stream), "findNeighbors");
    }

    /*! exchangeHalos(std::tie(fields...), ...) domain.hpp:332-337: device arrays with nParticlesWithHalos() elements;
     *  the halo elements are overwritten with the owners' values.  Element sizes must be multiples of 4 bytes. */
    template<class... Arrays>
    void exchangeHalos(void* stream, Arrays*... arrays)
    {
        void* ptrs[]  = {static_cast<void*>(arrays)...};
        int sizes[]   = {int(sizeof(Arrays))...};
        csCheck(cs_domain_exchange_halos(d_, ptrs, sizes, int(sizeof...(Arrays)), stream), "Domain::exchangeHalos");
    }

    /*! reapplySync(std::tie(fields...), ...) domain.hpp:297-329: replay the particle exchange and reordering of the last
     *  sync for further fields.  Pass (before, after) device pointer pairs: `before` in the order the sync consumed
     *  (replaySizeBefore() elements), `after` with nParticlesWithHalos() elements. */
    template<class... Arrays>
    void reapplySync(void* stream, std::pair<const Arrays*, Arrays*>... fields)
    {
        const void* src[] = {static_cast<const void*>(fields.first)...};
        void* dst[]       = {static_cast<void*>(fields.second)...};
        int sizes[]       = {int(sizeof(Arrays))...};
        csCheck(cs_domain_reapply_sync(d_, src, dst, sizes, int(sizeof...(Arrays)), stream), "Domain::reapplySync");
    }

    uint64_t replaySizeBefore() const
    {
        uint64_t info[4];
        csCheck(cs_domain_replay_info(d_, info), "Domain::replaySizeBefore");
        return info[0];
    }

    //! setHaloFactor domain.hpp:365
    void setHaloFactor(float factor) { csCheck(cs_domain_set_halo_factor(d_, factor), "Domain::setHaloFactor"); }

    cs_domain_t* handle() { return d_; }

private:
    cs_domain_t* d_{nullptr};
    uint64_t info_[8]{};
    double box_[6]{};
};

} // namespace cstone_b200
