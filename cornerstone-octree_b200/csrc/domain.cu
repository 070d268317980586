/* cs_domain_*: device-resident orchestration of the domain-sync hot path, the B200 counterpart of
 * cstone::Domain<KeyType,T,Gpu>::sync (reference domain/domain.hpp:169-218) with its GlobalAssignment
 * (domain/assignment.hpp:92-203) and FocusedOctree (focus/octree_focus_mpi.hpp:100-252,511-603,
 * focus/octree_focus.hpp:65-212) collaborators.
 *
 * The update cadence of the reference is kept exactly (the trees lag the particles by design): one global-tree
 * rebalance per sync (a loop only on the first call or after large count changes, assignment.hpp:115-123), a focus
 * tree that is converged on the first call and then updated once per sync, node counts / layout recomputed after
 * every tree update.  What changes is where things run: every array lives in HBM, each step is one of the kernels in
 * this library, and the host only reads back the handful of scalars the control flow needs.
 *
 * One rank: nothing is exchanged, every leaf is in focus (so MAC flags cannot change any rebalance decision,
 * focus/rebalance.hpp:31-73) and no search box can leave the assigned SFC range (traversal/collisions.hpp:81-90),
 * therefore halo flags are all zero; those two stages are skipped.  Several ranks (a communicator attached with
 * cs_domain_attach_comm): global assignment, exchangeParticles, LET and halo exchange as described in DESIGN.md 4.
 *
 * Errors in collective phases: a rank whose local precondition fails returns its error; its peers are released by
 * cs_local_world_abort (thread ranks) or by the job launcher tearing down the process group (NCCL), as with an MPI
 * abort in the reference.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "assignment.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "cstone_b200.h"
#include "focus.cuh"

namespace csb
{

namespace
{

/* ---------------------------------------------------------------- small device buffer */
template<class E>
struct DevBuf
{
    E* p{nullptr};
    size_t cap{0};
    size_t n{0};

    DevBuf() = default;
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf()
    {
        cudaFree(p);
        for (void* q : retired)
            cudaFree(q);
    }

    /*! set once the allocation has been mapped by peer ranks (Comm::sharePointers): a later reallocation must not free
     *  memory that peers may still have mapped, so the old allocation is parked until the buffer is destroyed (and the
     *  buffer grows geometrically from then on, which bounds what can be parked) */
    bool sharedWithPeers{false};
    std::vector<void*> retired;

    //! grow to m elements; keeps the first min(n, m) elements when keep is set; new elements are zeroed when zeroNew
    int resize(size_t m, cudaStream_t s, bool keep = false, bool zeroNew = false, double growth = 1.05)
    {
        if (m > cap)
        {
            if (sharedWithPeers) { growth = std::max(growth, 1.5); }
            size_t newCap = std::max<size_t>(size_t(double(m) * growth), 64);
            E* q          = nullptr;
            CSB_CHECK(cudaMalloc(&q, newCap * sizeof(E)));
            if (keep && n) { CSB_CHECK(cudaMemcpyAsync(q, p, n * sizeof(E), cudaMemcpyDeviceToDevice, s)); }
            if (p)
            {
                CSB_CHECK(cudaStreamSynchronize(s));
                if (sharedWithPeers) { retired.push_back(p); }
                else { CSB_CHECK(cudaFree(p)); }
            }
            p   = q;
            cap = newCap;
        }
        if (zeroNew && m > n) { CSB_CHECK(cudaMemsetAsync(p + n, 0, (m - n) * sizeof(E), s)); }
        n = m;
        return 0;
    }
    void swap(DevBuf& o)
    {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
        std::swap(n, o.n);
        std::swap(sharedWithPeers, o.sharedWithPeers);
        retired.swap(o.retired);
    }
};

#define CSB_TRY(expr)                                                                                                  \
    do                                                                                                                 \
    {                                                                                                                  \
        if (int e__ = (expr)) { return e__; }                                                                          \
    } while (0)

template<class K>
struct OctreeBufs
{
    int numLeaves{0}, numInternal{0}, numNodes{0};
    DevBuf<K> prefixes;
    DevBuf<int> childOffsets, parents, levelRange, internalToLeaf, leafToInternal;
    std::vector<int> levelRangeHost;

    int resize(int nLeaves, cudaStream_t s)
    {
        numLeaves   = nLeaves;
        numInternal = (nLeaves - 1) / 7;
        numNodes    = numLeaves + numInternal;
        CSB_TRY(prefixes.resize(numNodes, s));
        CSB_TRY(childOffsets.resize(numNodes + 1, s));
        CSB_TRY(parents.resize(std::max(1, (numNodes - 1) / 8), s));
        CSB_TRY(levelRange.resize(KeyTraits<K>::maxLevel + 2, s));
        CSB_TRY(internalToLeaf.resize(numNodes, s));
        CSB_TRY(leafToInternal.resize(numNodes, s));
        levelRangeHost.resize(KeyTraits<K>::maxLevel + 2);
        return 0;
    }
    const int* leafToInternalLeaves() const { return leafToInternal.p + numInternal; }
};

inline int keysDispatch(int kind, const float* x, const float* y, const float* z, uint32_t* k, size_t n,
                        const double* lim, const int* bnd, cudaStream_t s)
{
    return cs_compute_sfc_keys_u32f(kind, x, y, z, k, n, lim, bnd, s);
}
inline int keysDispatch(int kind, const float* x, const float* y, const float* z, uint64_t* k, size_t n,
                        const double* lim, const int* bnd, cudaStream_t s)
{
    return cs_compute_sfc_keys_u64f(kind, x, y, z, k, n, lim, bnd, s);
}
inline int keysDispatch(int kind, const double* x, const double* y, const double* z, uint64_t* k, size_t n,
                        const double* lim, const int* bnd, cudaStream_t s)
{
    return cs_compute_sfc_keys_u64d(kind, x, y, z, k, n, lim, bnd, s);
}
inline int sortDispatch(uint64_t* k, uint32_t* v, size_t n, uint64_t* kb, uint32_t* vb, void* t, size_t tb,
                        cudaStream_t s)
{
    return sortByKeyU64(k, v, n, kb, vb, t, tb, s);
}
inline int sortDispatch(uint32_t* k, uint32_t* v, size_t n, uint32_t* kb, uint32_t* vb, void* t, size_t tb,
                        cudaStream_t s)
{
    return sortByKeyU32(k, v, n, kb, vb, t, tb, s);
}
inline int sortIotaDispatch(uint64_t* k, uint32_t* v, uint32_t first, size_t n, uint64_t* kb, uint32_t* vb, void* t,
                            size_t tb, cudaStream_t s)
{
    return sortByKeyIotaU64(k, v, first, n, kb, vb, t, tb, s);
}
inline int sortIotaDispatch(uint32_t* k, uint32_t* v, uint32_t first, size_t n, uint32_t* kb, uint32_t* vb, void* t,
                            size_t tb, cudaStream_t s)
{
    return sortByKeyIotaU32(k, v, first, n, kb, vb, t, tb, s);
}
template<class K>
size_t sortTempBytesT(size_t n)
{
    if constexpr (sizeof(K) == 8) { return sortTempBytesU64(n); }
    else { return sortTempBytesU32(n); }
}
template<class K>
size_t linkTempBytesT(int numLeaves)
{
    if constexpr (sizeof(K) == 8) { return buildOctreeTempBytesU64(numLeaves); }
    else { return buildOctreeTempBytesU32(numLeaves); }
}

struct DomainBase
{
    virtual ~DomainBase()                                                                         = default;
    virtual int sync(const void* x, const void* y, const void* z, const void* h, const void* keys, size_t n,
                     bool hostInput, cudaStream_t s)                                               = 0;
    virtual int info(uint64_t* out, double* box) const                                             = 0;
    virtual void* ptr(int field)                                                                   = 0;
    virtual int neighbors(uint32_t ngmax, uint32_t* nb, uint32_t* nc, cudaStream_t s)              = 0;
    virtual int download(void* x, void* y, void* z, void* h, void* keys, cudaStream_t s)           = 0;
    virtual int reset(cudaStream_t s)                                                              = 0;
    virtual int attachComm(Comm* c)                                                                = 0;
    virtual int exchangeHaloFields(void* const* arrays, const int* elemBytes, int numArrays, cudaStream_t s) = 0;
    virtual int reapplySync(const void* const* before, void* const* after, const int* elemBytes, int numArrays,
                            cudaStream_t s)                                                                  = 0;
    virtual int replayInfo(uint64_t* info) const                                                             = 0;
    //! Domain::setHaloFactor (domain/domain.hpp:365): enlarges the halo search radius of the following syncs
    float haloFactor_{1.0f};
    int keyBytes{0}, realBytes{0};
};

template<class K, class T>
class DomainImpl : public DomainBase
{
public:
    DomainImpl(int rank, int numRanks, unsigned bucket, unsigned bucketFocus, float theta, const double* lim,
               const int* bnd)
        : rank_(rank)
        , numRanks_(numRanks)
        , bucket_(bucket)
        , bucketFocus_(bucketFocus)
        , theta_(theta)
    {
        std::copy(lim, lim + 6, lim_);
        std::copy(lim, lim + 6, lim0_);
        std::copy(bnd, bnd + 3, bnd_);
        keyBytes  = sizeof(K);
        realBytes = sizeof(T);
    }

    ~DomainImpl() override
    {
        for (auto ps : pushStreams_)
        {
            cudaStreamSynchronize(ps);
            cudaStreamDestroy(ps);
        }
        for (auto e : pushDone_)
            cudaEventDestroy(e);
        if (pushReady_) { cudaEventDestroy(pushReady_); }
        if (copyStream_)
        {
            cudaStreamSynchronize(copyStream_);
            cudaEventDestroy(copyReady_);
            cudaEventDestroy(copyDone_);
            cudaStreamDestroy(copyStream_);
        }
    }

    //! back to the freshly constructed state (next sync is a "first call"); device buffers are kept
    int reset(cudaStream_t s) override
    {
        firstCall_ = true;
        log_.valid = false;
        start_ = end_ = bufSize_ = 0;
        std::copy(lim0_, lim0_ + 6, lim_);
        const double unitBox[6] = {0, 1, 0, 1, 0, 1};
        std::copy(unitBox, unitBox + 6, focusLim_);
        fLower_.clear();
        globDispl_.clear();
        geoCenters_.n = geoSizes_.n = 0;
        macs_.n                     = 0;
        return init(s);
    }

    int init(cudaStream_t s)
    {
        // GlobalAssignment ctor (assignment.hpp:53-74): spanning tree of the initial rank splits, counts start at
        // bucketSize - 1 (for one rank this is the root node)
        K root[2] = {0, nodeRange<K>(0)};
        {
            std::vector<K> initial = initialGlobalTree<K>(numRanks_);
            numGlobalLeaves_       = int(initial.size()) - 1;
            std::vector<uint32_t> c0(numGlobalLeaves_, bucket_ - 1);
            CSB_TRY(gLeaves_.resize(initial.size(), s));
            CSB_TRY(gCounts_.resize(numGlobalLeaves_, s));
            CSB_CHECK(cudaMemcpyAsync(gLeaves_.p, initial.data(), initial.size() * sizeof(K), cudaMemcpyHostToDevice, s));
            CSB_CHECK(cudaMemcpyAsync(gCounts_.p, c0.data(), c0.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            CSB_CHECK(cudaStreamSynchronize(s));
        }

        // FocusedOctree ctor (octree_focus_mpi.hpp:56-83): root leaf with count bucketSizeFocus + 1
        uint32_t c1 = bucketFocus_ + 1;
        CSB_TRY(fLeaves_.resize(2, s));
        CSB_TRY(fLeafCounts_.resize(1, s));
        CSB_TRY(fCounts_.resize(1, s));
        CSB_TRY(macs_.resize(1, s, false, true));
        CSB_CHECK(cudaMemcpyAsync(fLeaves_.p, root, sizeof(root), cudaMemcpyHostToDevice, s));
        CSB_CHECK(cudaMemcpyAsync(fLeafCounts_.p, &c1, sizeof(c1), cudaMemcpyHostToDevice, s));
        CSB_CHECK(cudaMemcpyAsync(fCounts_.p, &c1, sizeof(c1), cudaMemcpyHostToDevice, s));
        CSB_TRY(scalars_.resize(64, s, false, true));
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_TRY(linkTree(fLeaves_, 1, fTree_, s));
        CSB_TRY(linkTree(gLeaves_, numGlobalLeaves_, gTree_, s));
        return 0;
    }

    /* ------------------------------------------------------------ sync */
    int sync(const void* xin, const void* yin, const void* zin, const void* hin, const void* keysIn, size_t nIn,
             bool hostInput, cudaStream_t s) override
    {
        auto kindOfCopy = hostInput ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        if (xin)
        {
            CSB_REQUIRE(yin && zin && hin, "x, y, z, h must be given together");
            if (!firstCall_)
            {
                CSB_REQUIRE(nIn == size_t(bufSize_), "Domain sync: input array sizes are inconsistent");
            }
            CSB_TRY(x_.resize(nIn, s));
            CSB_TRY(y_.resize(nIn, s));
            CSB_TRY(z_.resize(nIn, s));
            CSB_TRY(h_.resize(nIn, s));
            CSB_CHECK(cudaMemcpyAsync(x_.p, xin, nIn * sizeof(T), kindOfCopy, s));
            CSB_CHECK(cudaMemcpyAsync(y_.p, yin, nIn * sizeof(T), kindOfCopy, s));
            CSB_CHECK(cudaMemcpyAsync(z_.p, zin, nIn * sizeof(T), kindOfCopy, s));
            if (hostInput)
            {
                // h is not needed before the particle exchange / gatherArrays: its upload runs on a second stream and
                // overlaps the key generation, the sort and the global tree update
                if (!copyStream_)
                {
                    CSB_CHECK(cudaStreamCreateWithFlags(&copyStream_, cudaStreamNonBlocking));
                    CSB_CHECK(cudaEventCreateWithFlags(&copyReady_, cudaEventDisableTiming));
                    CSB_CHECK(cudaEventCreateWithFlags(&copyDone_, cudaEventDisableTiming));
                }
                CSB_CHECK(cudaEventRecord(copyReady_, s)); // h_ is allocated and no longer in use on s
                CSB_CHECK(cudaStreamWaitEvent(copyStream_, copyReady_, 0));
                CSB_CHECK(cudaMemcpyAsync(h_.p, hin, nIn * sizeof(T), kindOfCopy, copyStream_));
                CSB_CHECK(cudaEventRecord(copyDone_, copyStream_));
                hUploadPending_ = true;
            }
            else { CSB_CHECK(cudaMemcpyAsync(h_.p, hin, nIn * sizeof(T), kindOfCopy, s)); }
            CSB_TRY(keys_.resize(nIn, s));
            if (keysIn) { CSB_CHECK(cudaMemcpyAsync(keys_.p, keysIn, nIn * sizeof(K), kindOfCopy, s)); }
            else { CSB_CHECK(cudaMemsetAsync(keys_.p, 0, nIn * sizeof(K), s)); }
            if (firstCall_)
            {
                start_   = 0;
                end_     = LocalIndex(nIn);
                bufSize_ = LocalIndex(nIn);
            }
        }
        else { CSB_REQUIRE(!firstCall_, "the first sync needs input arrays"); }
        CSB_REQUIRE(size_t(bufSize_) < (size_t(1) << 30), "at most 2^30 - 1 particles per rank");
        CSB_REQUIRE(comm_->size() == numRanks_, "multi-rank domains need cs_domain_attach_comm before the first sync");

        const size_t numPart      = end_ - start_;
        const LocalIndex prevSize = bufSize_;
        Comm& comm                = *comm_;
        const int P               = comm.size();
        const int me              = comm.rank();
        log_.valid                = false;

        phase(nullptr, s);
        /* ---- GlobalAssignment::assign (assignment.hpp:92-144) */
        CSB_TRY(updateBox(numPart, s));
        phase("updateBox", s);
        CSB_TRY(keysDispatch(0, x_.p + start_, y_.p + start_, z_.p + start_, keys_.p + start_, numPart, lim_, bnd_, s));
        // the ordering is indexed by buffer position (primitives_acc.hpp:97-103): ordering[start + i] = start + i
        CSB_TRY(ordering_.resize(std::max<size_t>(bufSize_, 1), s));
        CSB_TRY(sortPairs(keys_.p + start_, ordering_.p + start_, numPart, s, (long long)start_));
        phase("keys+sort", s);

        unsigned maxCount = 0;
        CSB_TRY(updateGlobalTree(keys_.p + start_, numPart, &maxCount, s));
        if (firstCall_ || maxCount >= 8 * bucket_)
        {
            do
            {
                CSB_TRY(updateGlobalTree(keys_.p + start_, numPart, &maxCount, s));
            } while (maxCount > bucket_);
        }
        CSB_TRY(linkTree(gLeaves_, numGlobalLeaves_, gTree_, s));

        phase("globalTree", s);
        // makeSfcAssignment on the host from the replicated leaves and counts (assignment.hpp:125-134)
        if (P == 1)
        {
            // one rank: the assignment is the whole curve whatever the counts are; skip the downloads
            assignment_.boundaries  = {K(0), nodeRange<K>(0)};
            assignment_.treeOffsets = {0, numGlobalLeaves_};
            assignment_.counts      = {0};
        }
        else
        {
            gLeavesHost_.resize(size_t(numGlobalLeaves_) + 1);
            gCountsHost_.resize(numGlobalLeaves_);
            CSB_CHECK(cudaMemcpyAsync(gLeavesHost_.data(), gLeaves_.p, gLeavesHost_.size() * sizeof(K),
                                      cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaMemcpyAsync(gCountsHost_.data(), gCounts_.p, gCountsHost_.size() * sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaStreamSynchronize(s));
            assignment_ = makeSfcAssignment<K>(P, gCountsHost_, gLeavesHost_.data());
        }

        // createSendRangesGpu (domaindecomp.hpp:178-191): particles with key >= 2^(3L) (removeKey) fall outside
        CSB_TRY(boundaryKeys_.resize(size_t(P) + 1, s));
        CSB_CHECK(cudaMemcpyAsync(boundaryKeys_.p, assignment_.boundaries.data(), (size_t(P) + 1) * sizeof(K),
                                  cudaMemcpyHostToDevice, s));
        CSB_TRY(scalars_.resize(std::max<size_t>(64, 32 + size_t(P) + 1), s, true, true));
        uint32_t* sendIdxDev = scalars_.p + 32;
        CSB_TRY(lowerBounds<K>(keys_.p + start_, numPart, boundaryKeys_.p, P + 1, sendIdxDev, s));
        std::vector<uint32_t> sendIdx(size_t(P) + 1);
        CSB_CHECK(cudaMemcpyAsync(sendIdx.data(), sendIdxDev, sendIdx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  s));
        CSB_CHECK(cudaStreamSynchronize(s));
        const LocalIndex numSendDown = sendIdx[me];
        const LocalIndex numPresent  = sendIdx[me + 1] - sendIdx[me];
        const LocalIndex numAssigned = P == 1 ? numPresent : assignment_.counts[me];

        phase("assignment+sendRanges", s);
        if (hUploadPending_)
        {
            CSB_CHECK(cudaStreamWaitEvent(s, copyDone_, 0)); // everything below may read h
            hUploadPending_ = false;
        }
        /* ---- GlobalAssignment::distribute (assignment.hpp:167-203) */
        LocalIndex envStart = start_, envEnd = end_;
        // one rank / nothing received: the present particles are already sorted, the reference's second sort is the
        // identity
        const K* keyView          = keys_.p + start_ + numSendDown;
        const uint32_t* orderView = ordering_.p + start_ + numSendDown;
        if (P > 1)
        {
            CSB_REQUIRE(numAssigned >= numPresent, "global node counts are inconsistent with the local particles");
            const LocalIndex numRecv = numAssigned - numPresent;
            BufferDescription o1{start_, end_, bufSize_};
            const LocalIndex exchangeSize = exchangeBufferSize(o1, numPresent, numAssigned);
            // lowMemReallocate + zero fill of the new key range (domain.hpp:465-468)
            CSB_TRY(x_.resize(exchangeSize, s, true));
            CSB_TRY(y_.resize(exchangeSize, s, true));
            CSB_TRY(z_.resize(exchangeSize, s, true));
            CSB_TRY(h_.resize(exchangeSize, s, true));
            CSB_TRY(keys_.resize(exchangeSize, s, true, true));
            CSB_TRY(ordering_.resize(exchangeSize, s, true));
            BufferDescription o1e{start_, end_, exchangeSize};
            const LocalIndex recvStart = receiveStart(o1e, numRecv);
            /* (x,y,z,h) records of the present particles at their buffer positions: the pack kernels of the exchange
             * and the gather into the new order read one 32-byte record per particle through the ordering instead of
             * four scattered elements (a random 8-byte read costs a 128-byte DRAM fetch); the received particles are
             * added below */
            CSB_TRY(recBuf_.resize(size_t(exchangeSize) * 4 * sizeof(T), s));
            {
                const void* src[4] = {x_.p, y_.p, z_.p, h_.p};
                CSB_TRY(packRecords4(src, start_, numPart, recBuf_.p, int(sizeof(T)), s));
            }
            phase("  pack records", s);
            std::vector<uint32_t> recvCounts;
            CSB_TRY(exchangeParticles(sendIdx, recvStart, numRecv, recvCounts, s));
            if (numRecv)
            {
                const void* src[4] = {x_.p, y_.p, z_.p, h_.p};
                CSB_TRY(packRecords4(src, recvStart, numRecv, recBuf_.p, int(sizeof(T)), s));
            }
            log_.recvStart  = recvStart;
            log_.numRecv    = numRecv;
            log_.recvCounts = recvCounts;
            phase("exchangeParticles", s);
            assignedEnvelope(o1e, numRecv, &envStart, &envEnd);
            bufSize_ = exchangeSize;
            if (numRecv)
            {
                /* The reference sorts the whole envelope - the present particles including those that were just sent
                 * away, plus the received ones (assignment.hpp:197-201) - and then looks at the assigned sub-range.
                 * Only the assigned particles matter afterwards, so the sort runs on them alone: the present-assigned
                 * keys (a sorted, contiguous piece of the first sort) and the received keys, in envelope order so that
                 * the stable sort breaks ties between equal keys exactly as the envelope sort does. */
                const bool recvFirst        = recvStart < start_;
                const LocalIndex offPresent = recvFirst ? numRecv : 0;
                const LocalIndex offRecv    = recvFirst ? 0 : numPresent;
                CSB_TRY(assignedKeys_.resize(numAssigned, s));
                CSB_TRY(assignedOrder_.resize(numAssigned, s));
                CSB_CHECK(cudaMemcpyAsync(assignedKeys_.p + offPresent, keys_.p + start_ + numSendDown,
                                          size_t(numPresent) * sizeof(K), cudaMemcpyDeviceToDevice, s));
                CSB_CHECK(cudaMemcpyAsync(assignedOrder_.p + offPresent, ordering_.p + start_ + numSendDown,
                                          size_t(numPresent) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
                // the key kernel keeps entries that are marked with removeKey: the output starts out as zeros
                CSB_CHECK(cudaMemsetAsync(assignedKeys_.p + offRecv, 0, size_t(numRecv) * sizeof(K), s));
                CSB_TRY(keysDispatch(0, x_.p + recvStart, y_.p + recvStart, z_.p + recvStart,
                                     assignedKeys_.p + offRecv, numRecv, lim_, bnd_, s));
                CSB_TRY(cs_sequence_u32(recvStart, numRecv, assignedOrder_.p + offRecv, s));
                // the present particles and the block of every source rank are sorted runs: a stable merge in buffer
                // order equals the stable sort (merge.cu)
                {
                    std::vector<size_t> runs{0};
                    if (!recvFirst) { runs.push_back(numPresent); }
                    for (int r = 0; r < P; ++r)
                        runs.push_back(runs.back() + recvCounts[r]);
                    if (recvFirst) { runs.push_back(runs.back() + numPresent); }
                    CSB_REQUIRE(runs.back() == size_t(numAssigned), "merge runs do not cover the assigned particles");
                    CSB_TRY(keyBuf_.resize(std::max<size_t>(numAssigned, 1), s));
                    CSB_TRY(valueBuf_.resize(std::max<size_t>(numAssigned, 1), s));
                    bool inBuffers = false;
                    if (numPresent > numRecv)
                    {
                        // steady state, few particles migrate: merge the small received runs among themselves first,
                        // the large present run then takes part in a single round instead of log2(P)
                        const size_t firstRecv = recvFirst ? 0 : 1;
                        CSB_TRY(mergeSortedRuns<K>(assignedKeys_.p, assignedOrder_.p, runs.data() + firstRecv, P,
                                                   keyBuf_.p, valueBuf_.p, s));
                        const size_t two[3] = {0, recvFirst ? size_t(numRecv) : size_t(numPresent), size_t(numAssigned)};
                        CSB_TRY(mergeSortedRuns<K>(assignedKeys_.p, assignedOrder_.p, two, 2, keyBuf_.p, valueBuf_.p, s,
                                                   &inBuffers));
                    }
                    else
                    {
                        CSB_TRY(mergeSortedRuns<K>(assignedKeys_.p, assignedOrder_.p, runs.data(), int(runs.size()) - 1,
                                                   keyBuf_.p, valueBuf_.p, s, &inBuffers));
                    }
                    if (inBuffers) // an odd number of merge rounds: the result sits in the double buffers
                    {
                        assignedKeys_.swap(keyBuf_);
                        assignedOrder_.swap(valueBuf_);
                    }
                }
                keyView   = assignedKeys_.p;
                orderView = assignedOrder_.p;
            }
        }

        // what reapplySync replays (ExchangeLog, domain/index_ranges.hpp:186-210, plus the assigned ordering)
        log_.order       = orderView;
        log_.numAssigned = numAssigned;
        log_.prevStart   = start_;
        log_.prevSize    = prevSize;
        log_.sendIdx     = sendIdx;
        if (P == 1) { log_.recvStart = log_.numRecv = 0, log_.recvCounts.assign(1, 0); }
        log_.valid = true;

        phase("keys+sort received", s);
        /* ---- gatherArrays(x,y,z,h) to offset 0 (domain.hpp:187) */
        CSB_TRY(sx_.resize(std::max<size_t>(numAssigned, 1), s));
        CSB_TRY(sy_.resize(std::max<size_t>(numAssigned, 1), s));
        CSB_TRY(sz_.resize(std::max<size_t>(numAssigned, 1), s));
        CSB_TRY(sh_.resize(std::max<size_t>(numAssigned, 1), s));
        {
            const void* src[4] = {x_.p, y_.p, z_.p, h_.p};
            void* dst[4]       = {sx_.p, sy_.p, sz_.p, sh_.p};
            if (P > 1) { CSB_TRY(gatherFromRecords4(orderView, numAssigned, recBuf_.p, dst, int(sizeof(T)), s)); }
            else { CSB_TRY(cs_gather_arrays4(orderView, numAssigned, bufSize_, src, dst, int(sizeof(T)), s)); }
        }

        phase("gatherArrays", s);
        /* ---- focus tree (domain.hpp:189-213) */
        if (P > 1) { return syncFocusMultiRank(keyView, numAssigned, s); }
        if (firstCall_)
        {
            int converged = 0;
            while (!converged)
            {
                CSB_TRY(updateFocusTree(&converged, s));
                CSB_TRY(updateFocusCounts(keyView, numAssigned, s));
            }
            phase("focus converge", s);
        }
        {
            int converged = 0;
            CSB_TRY(updateFocusTree(&converged, s));
            CSB_TRY(updateFocusCounts(keyView, numAssigned, s));
            // discoverHalos + computeLayout (octree_focus_mpi.hpp:511-582): no foreign leaves on one rank
            CSB_TRY(macs_.resize(fTree_.numNodes, s));
            CSB_CHECK(cudaMemsetAsync(macs_.p, 0, size_t(fTree_.numNodes), s));
            CSB_TRY(layout_.resize(fTree_.numLeaves + 1, s));
            CSB_TRY(layoutCounts(fLeafCounts_.p, macs_.p, fTree_.leafToInternalLeaves(), fTree_.numLeaves, 0,
                                 fTree_.numLeaves, layout_.p, s));
            CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(fTree_.numLeaves) + 1), s));
            CSB_TRY(exclusiveScanU32(layout_.p, layout_.p, size_t(fTree_.numLeaves) + 1, scanTmp_.p, s));
            phase("focus update+layout", s);
        }

        /* ---- updateLayout (domain.hpp:490-537): new buffer = [0, numAssigned) without halos; keys move to offset 0 */
        CSB_TRY(keyBuf_.resize(std::max<size_t>(numAssigned, 1), s));
        CSB_CHECK(cudaMemcpyAsync(keyBuf_.p, keyView, size_t(numAssigned) * sizeof(K), cudaMemcpyDeviceToDevice, s));
        keys_.swap(keyBuf_);
        keys_.n = numAssigned;
        x_.swap(sx_);
        y_.swap(sy_);
        z_.swap(sz_);
        h_.swap(sh_);
        x_.n = y_.n = z_.n = h_.n = numAssigned;

        start_     = 0;
        end_       = numAssigned;
        bufSize_   = numAssigned;
        firstCall_ = false;
        CSB_CHECK(cudaStreamSynchronize(s)); // the caller's input buffers are free again, the results are in place
        return 0;
    }

    /* ============================================================ multi-rank: LET, halos (domain.hpp:189-217) */

    int syncFocusMultiRank(const K* keyView, LocalIndex numAssigned, cudaStream_t s)
    {
        Comm& comm           = *comm_;
        const int P          = comm.size();
        const float invTheta = 1.0f / theta_ + 0.5f; // invThetaMinMac (traversal/macs.hpp:28)

        // 64-bit scan of the replicated global counts for rangeCount
        {
            std::vector<uint64_t> scan(gCountsHost_.size() + 1, 0);
            for (size_t i = 0; i < gCountsHost_.size(); ++i)
                scan[i + 1] = scan[i] + gCountsHost_[i];
            CSB_TRY(gCountScan_.resize(scan.size(), s));
            CSB_CHECK(cudaMemcpyAsync(gCountScan_.p, scan.data(), scan.size() * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                      s));
            CSB_CHECK(cudaStreamSynchronize(s));
        }

        if (firstCall_)
        {
            // FocusedOctree::converge (octree_focus_mpi.hpp:584-602)
            int numConverged = 0;
            while (numConverged != P)
            {
                int converged = 0;
                CSB_TRY(letUpdateMinMac(invTheta, false, s));
                phase("  minMac", s);
                CSB_TRY(updateFocusTree(&converged, s));
                phase("  updateFocusTree(total)", s);
                CSB_TRY(letUpdateCounts(keyView, numAssigned, s));
                phase("  letUpdateCounts", s);
                CSB_TRY(allreduceSumInt(converged, &numConverged, s));
                phase("  allreduce converged", s);
            }
        }

        int fail = 0, maxRep = 10;
        do
        {
            int converged = 0;
            CSB_TRY(letUpdateMinMac(invTheta, true, s));
            phase("minMac", s);
            CSB_TRY(updateFocusTree(&converged, s));
            phase("updateFocusTree(total)", s);
            CSB_TRY(letUpdateCounts(keyView, numAssigned, s));
            phase("letUpdateCounts", s);
            CSB_TRY(letDiscoverHalos(s));
            phase("discoverHalos", s);
            int localFail = 0;
            CSB_TRY(letComputeLayout(&localFail, s));
            CSB_TRY(allreduceSumInt(localFail, &fail, s));
            phase("computeLayout", s);
            CSB_TRY(haloExchangeRequests(s));
            phase("haloExchangeRequests", s);
        } while (fail && maxRep--);

        /* ---- updateLayout (domain.hpp:490-537) */
        const int me              = comm.rank();
        const LocalIndex newStart = layoutAt_[2 * me];
        const LocalIndex newEnd   = layoutAt_[2 * me + 1];
        const LocalIndex newSize  = layoutAt_[2 * P];
        CSB_REQUIRE(newEnd - newStart == numAssigned, "layout of the assigned leaves does not match the particles");
        CSB_TRY(keyBuf_.resize(std::max<size_t>(numAssigned, 1), s));
        CSB_CHECK(cudaMemcpyAsync(keyBuf_.p, keyView, size_t(numAssigned) * sizeof(K), cudaMemcpyDeviceToDevice, s));
        CSB_TRY(keys_.resize(std::max<size_t>(newSize, 1), s));
        CSB_CHECK(cudaMemsetAsync(keys_.p, 0, size_t(newSize) * sizeof(K), s));
        CSB_CHECK(cudaMemcpyAsync(keys_.p + newStart, keyBuf_.p, size_t(numAssigned) * sizeof(K),
                                  cudaMemcpyDeviceToDevice, s));
        DevBuf<T>* dst[4] = {&x_, &y_, &z_, &h_};
        DevBuf<T>* src[4] = {&sx_, &sy_, &sz_, &sh_};
        for (int k = 0; k < 4; ++k)
        {
            CSB_TRY(dst[k]->resize(std::max<size_t>(newSize, 1), s));
            CSB_CHECK(cudaMemcpyAsync(dst[k]->p + newStart, src[k]->p, size_t(numAssigned) * sizeof(T),
                                      cudaMemcpyDeviceToDevice, s));
            dst[k]->n = newSize;
        }
        keys_.n  = newSize;
        start_   = newStart;
        end_     = newEnd;
        bufSize_ = newSize;

        phase("updateLayout", s);
        /* ---- setupHalos (domain.hpp:479-488) */
        CSB_TRY(exchangeHalos(s));
        CSB_TRY(keysDispatch(0, x_.p, y_.p, z_.p, keys_.p, start_, lim_, bnd_, s));
        CSB_TRY(keysDispatch(0, x_.p + end_, y_.p + end_, z_.p + end_, keys_.p + end_, size_t(bufSize_) - end_, lim_,
                             bnd_, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        phase("exchangeHalos+halo keys", s);
        firstCall_ = false;
        return 0;
    }

    int allreduceSumInt(int value, int* sum, cudaStream_t s)
    {
        std::vector<int> all(comm_->size());
        CSB_TRY(comm_->allgatherHost(&value, sizeof(int), all.data(), s));
        *sum = 0;
        for (int v : all)
            *sum += v;
        return 0;
    }

    /*! personalised all-to-all of DEVICE buffers (the Isend/Probe/Recv idiom of focus/exchange_focus.hpp and
     *  domain/exchange_keys.hpp): byte counts go through the host allgather, payloads move device to device.  Message
     *  from rank r lands at recvBuf + recvOff[r] (16-byte aligned), recvBytes[r] bytes. */
    int alltoallvDevice(const std::vector<const void*>& sendPtr, const std::vector<size_t>& sendBytes,
                        DevBuf<char>& recvBuf, std::vector<size_t>& recvOff, std::vector<size_t>& recvBytes,
                        cudaStream_t s)
    {
        Comm& comm   = *comm_;
        const int P  = comm.size();
        const int me = comm.rank();
        std::vector<uint64_t> sizes(P), all(size_t(P) * P);
        for (int r = 0; r < P; ++r)
            sizes[r] = r == me ? 0 : sendBytes[r];
        CSB_TRY(comm.allgatherHost(sizes.data(), P * sizeof(uint64_t), all.data(), s));
        auto pad = [](size_t b) { return (b + 15) & ~size_t(15); };
        recvOff.assign(P, 0);
        recvBytes.assign(P, 0);
        size_t total = 0;
        for (int r = 0; r < P; ++r)
        {
            recvBytes[r] = r == me ? 0 : all[size_t(r) * P + me];
            recvOff[r]   = total;
            total += pad(recvBytes[r]);
        }
        CSB_TRY(recvBuf.resize(std::max<size_t>(total, 16), s));
        std::vector<CommMessage> sends, recvs;
        for (int r = 0; r < P; ++r)
        {
            if (sizes[r]) { sends.push_back({r, const_cast<void*>(sendPtr[r]), size_t(sizes[r])}); }
            if (recvBytes[r]) { recvs.push_back({r, recvBuf.p + recvOff[r], recvBytes[r]}); }
        }
        return comm.exchange(sends, recvs, s);
    }

    /*! translateAssignment (domaindecomp.hpp:129-157) without moving the leaves: findNodeAbove / findNodeBelow of the
     *  P + 1 boundary keys are evaluated on the device, 2 (P + 1) integers come back.  fLower_[r] is also the
     *  findNodeAbove(boundary r) that updateMinMac needs. */
    int translateAssignment(cudaStream_t s)
    {
        const int P = comm_->size();
        CSB_TRY(letInts_.resize(LET_INTS + size_t(4) * (P + 1), s, true, true));
        int* boundsDev = letInts_.p + LET_INTS;
        CSB_TRY(focusBounds<K>(fLeaves_.p, fTree_.numLeaves + 1, boundaryKeys_.p, P + 1, boundsDev, s));
        std::vector<int> b(size_t(2) * (P + 1));
        CSB_CHECK(cudaMemcpyAsync(b.data(), boundsDev, b.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        fLower_.assign(b.begin(), b.begin() + P + 1);
        fAssign_.assign(P, {0, 0});
        // (a member: the asynchronous upload below reads it after this function has returned; the next call writes it
        // only after its own download has synchronised the stream)
        std::vector<int>& flat = fAssignFlat_;
        flat.resize(size_t(2) * P);
        for (int r = 0; r < P; ++r)
        {
            int lo = b[r], hi = b[P + 1 + r + 1];
            if (hi < lo) { hi = lo; }
            fAssign_[r]     = {lo, hi};
            flat[2 * r]     = lo;
            flat[2 * r + 1] = hi;
        }
        // device copy of the focus assignment for the layout / halo request kernels
        CSB_TRY(fAssignDev_.resize(flat.size(), s));
        CSB_CHECK(cudaMemcpyAsync(fAssignDev_.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        return 0;
    }

    //! domaindecomp.hpp:159-168
    void extractPeerRanges()
    {
        peerRanges_.clear();
        for (int p : extPeers_)
            peerRanges_.push_back(fAssign_[p]);
        peerRanges_.push_back(fAssign_[comm_->rank()]);
        std::sort(peerRanges_.begin(), peerRanges_.end());
    }

    /*! the rank-to-rank part of FocusedOctree::updateTree (octree_focus_mpi.hpp:137-165): focus peers, treelet
     *  exchange with rejection of keys the owner does not have (focus/exchange_focus.hpp:60-230), treelet indices.
     *  Leaves, treelets and prefixes stay in HBM; the host sees O(P) scalars per call. */
    int letSyncWithPeers(cudaStream_t s)
    {
        Comm& comm   = *comm_;
        const int P  = comm.size();
        const int me = comm.rank();
        CSB_TRY(translateAssignment(s));
        int* peerFlagsDev = letInts_.p;            // [P]
        int* errorDev     = letInts_.p + LET_ERROR; // [1]

        // focusPeers (focus/peer_flags.hpp:33-57) with the global offsets of the PREVIOUS call (globDispl_)
        std::vector<int> extFlags(P, 0), allFlags(size_t(P) * P);
        if (globDispl_.empty()) { globDispl_.assign(size_t(P) + 1, 0); }
        CSB_CHECK(cudaMemsetAsync(letInts_.p, 0, LET_INTS * sizeof(int), s));
        for (int r = 0; r < P; ++r)
        {
            if (r == me) { continue; }
            CSB_TRY(notIncluded<K>(fLeaves_.p + fAssign_[r].first, fAssign_[r].second - fAssign_[r].first,
                                   gLeaves_.p + globDispl_[r], globDispl_[r + 1] - globDispl_[r], peerFlagsDev + r, s));
        }
        CSB_CHECK(cudaMemcpyAsync(extFlags.data(), peerFlagsDev, P * sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_TRY(comm.allgatherHost(extFlags.data(), P * sizeof(int), allFlags.data(), s));
        extPeers_.clear();
        intPeers_.clear();
        for (int r = 0; r < P; ++r)
        {
            if (extFlags[r]) { extPeers_.push_back(r); }
            if (allFlags[size_t(r) * P + me]) { intPeers_.push_back(r); }
        }

        // exchangeTreelets: my view of the peer's domain goes to the peer, straight out of the leaf array
        std::vector<const void*> sendPtr(P, nullptr);
        std::vector<size_t> sendBytes(P, 0), recvOff, recvBytes;
        for (int p : extPeers_)
        {
            sendPtr[p]   = fLeaves_.p + fAssign_[p].first;
            sendBytes[p] = (size_t(fAssign_[p].second - fAssign_[p].first) + 1) * sizeof(K);
        }
        CSB_TRY(alltoallvDevice(sendPtr, sendBytes, tlRecv_, recvOff, recvBytes, s));

        // checkTreelets + pruneTreelets: keys that are not leaf boundaries here are reported back and dropped
        const int numLeaves = fTree_.numLeaves;
        std::vector<int> recvCount(P, 0), keyOff(size_t(P) + 1, 0);
        for (int r = 0; r < P; ++r)
        {
            recvCount[r]  = int(recvBytes[r] / sizeof(K));
            keyOff[r + 1] = keyOff[r] + recvCount[r];
        }
        const int totalRecv = keyOff[P];
        CSB_TRY(tlValid_.resize(size_t(totalRecv) + 1, s));
        CSB_TRY(tlKeys_.resize(std::max(totalRecv, 1), s));
        CSB_TRY(rejKeys_.resize(std::max(totalRecv, 1), s));
        CSB_CHECK(cudaMemsetAsync(tlValid_.p + totalRecv, 0, sizeof(uint32_t), s));
        for (int p : intPeers_)
            CSB_TRY(checkTreelet<K>(reinterpret_cast<const K*>(tlRecv_.p + recvOff[p]), recvCount[p], fLeaves_.p,
                                    numLeaves, tlValid_.p + keyOff[p], s));
        CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(totalRecv) + 1), s));
        CSB_TRY(exclusiveScanU32(tlValid_.p, tlValid_.p, size_t(totalRecv) + 1, scanTmp_.p, s));
        // the scan runs over the concatenation of all treelets, so accepted / rejected keys of peer p start at
        // validBefore(p) / keyOff[p] - validBefore(p)
        std::vector<uint32_t> validAt(size_t(P) + 1, 0);
        {
            std::vector<int> idx(keyOff.begin(), keyOff.end());
            int* idxDev       = letInts_.p + LET_INTS + 2 * (P + 1);
            uint32_t* pickDev = reinterpret_cast<uint32_t*>(idxDev + (P + 1));
            CSB_CHECK(cudaMemcpyAsync(idxDev, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            CSB_TRY(pickU32(tlValid_.p, idxDev, P + 1, pickDev, s));
            CSB_CHECK(cudaMemcpyAsync(validAt.data(), pickDev, validAt.size() * sizeof(uint32_t),
                                      cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaStreamSynchronize(s));
        }
        for (int p : intPeers_)
            CSB_TRY(splitTreelet<K>(reinterpret_cast<const K*>(tlRecv_.p + recvOff[p]), recvCount[p],
                                    tlValid_.p + keyOff[p], tlKeys_.p + validAt[p],
                                    rejKeys_.p + (keyOff[p] - validAt[p]), s));

        // exchangeRejectedKeys: leaves of mine that the owner does not have are removed (nodeOps = 0)
        std::vector<const void*> rejPtr(P, nullptr);
        std::vector<size_t> rejBytes(P, 0), rejOff, rejRecvBytes;
        for (int p : intPeers_)
        {
            size_t numRejected = size_t(recvCount[p]) - (validAt[p + 1] - validAt[p]);
            rejPtr[p]          = rejKeys_.p + (keyOff[p] - validAt[p]);
            rejBytes[p]        = numRejected * sizeof(K);
        }
        CSB_TRY(alltoallvDevice(rejPtr, rejBytes, rejRecv_, rejOff, rejRecvBytes, s));
        bool changed = false;
        for (int p : extPeers_)
            changed = changed || rejRecvBytes[p] > 0;
        if (changed)
        {
            CSB_TRY(nodeOps_.resize(size_t(numLeaves) + 1, s));
            CSB_TRY(fillInt(nodeOps_.p, numLeaves + 1, 1, s));
            for (int p : extPeers_)
                CSB_TRY(rejectLeaves<K>(reinterpret_cast<const K*>(rejRecv_.p + rejOff[p]),
                                        int(rejRecvBytes[p] / sizeof(K)), fLeaves_.p, numLeaves + 1, nodeOps_.p, s));
            CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(numLeaves) + 1), s));
            CSB_TRY(exclusiveScanU32(reinterpret_cast<uint32_t*>(nodeOps_.p), reinterpret_cast<uint32_t*>(nodeOps_.p),
                                     size_t(numLeaves) + 1, scanTmp_.p, s));
            int newNumLeaves = 0;
            CSB_CHECK(cudaMemcpyAsync(&newNumLeaves, nodeOps_.p + numLeaves, sizeof(int), cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaStreamSynchronize(s));
            CSB_TRY(fLeavesAlt_.resize(size_t(newNumLeaves) + 1, s));
            CSB_TRY(rebalanceTree<K>(fLeaves_.p, numLeaves, newNumLeaves, nodeOps_.p, fLeavesAlt_.p, s));
            CSB_CHECK(cudaStreamSynchronize(s));
            fLeaves_.swap(fLeavesAlt_);
            CSB_TRY(linkTree(fLeaves_, newNumLeaves, fTree_, s));
        }

        // indexTreelets (exchange_focus.hpp:286-308) against the level-sorted prefixes
        treeletOffsets_.assign(size_t(P) + 1, 0);
        int numTreeletNodes = 0;
        for (int p = 0; p < P; ++p)
        {
            treeletOffsets_[p] = numTreeletNodes;
            int numKeys        = int(validAt[p + 1] - validAt[p]);
            if (numKeys > 1) { numTreeletNodes += numKeys - 1; }
        }
        treeletOffsets_[P] = numTreeletNodes;
        CSB_TRY(treeletIdx_.resize(std::max<size_t>(numTreeletNodes, 1), s));
        for (int p : intPeers_)
        {
            int nn = treeletOffsets_[p + 1] - treeletOffsets_[p];
            CSB_TRY(indexTreelet<K>(tlKeys_.p + validAt[p], nn, fTree_.prefixes.p, fTree_.levelRange.p,
                                    treeletIdx_.p + treeletOffsets_[p], errorDev, s));
        }
        int error = 0;
        CSB_CHECK(cudaMemcpyAsync(&error, errorDev, sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_REQUIRE(error == 0, "treelet node of a peer does not exist in the LET");

        if (changed) { CSB_TRY(translateAssignment(s)); }
        extractPeerRanges();
        globDispl_ = assignment_.treeOffsets;
        return 0;
    }

    /*! FocusedOctree::updateMinMac + updateMacs (octree_focus_mpi.hpp:422-499): MAC spheres from the geometric centres
     *  of the last tree update, then mark every node outside the focus that fails the MAC against a focus leaf */
    int letUpdateMinMac(float invTheta, bool accumulate, cudaStream_t s)
    {
        const int me       = comm_->rank();
        const int numNodes = fTree_.numNodes;
        if (geoCenters_.n != size_t(3) * numNodes) { CSB_TRY(updateGeoCenters(s)); } // first call: box (0,1)
        CSB_TRY(translateAssignment(s)); // the boundaries of THIS sync in the leaves of the previous tree update
        CSB_TRY(centers4_.resize(size_t(4) * numNodes, s));
        CSB_TRY(minMacCenters<T>(geoCenters_.p, geoSizes_.p, numNodes, invTheta, centers4_.p, s));
        if (accumulate) { CSB_REQUIRE(macs_.n == size_t(numNodes), "MAC flags not correctly allocated"); }
        CSB_TRY(macs_.resize(numNodes, s, true));
        if (!accumulate) { CSB_CHECK(cudaMemsetAsync(macs_.p, 0, size_t(numNodes), s)); }
        int fStart = fLower_[me];
        int fEnd   = fLower_[me + 1];
        fStart     = std::min(fStart, fTree_.numLeaves);
        fEnd       = std::min(fEnd, fTree_.numLeaves);
        return markMacs<K, T>(fTree_.prefixes.p, fTree_.childOffsets.p, fTree_.parents.p, centers4_.p, focusLim_, bnd_,
                              fLeaves_.p + fStart, fEnd - fStart, macs_.p, s);
    }

    //! FocusedOctree::updateCounts (octree_focus_mpi.hpp:193-252)
    int letUpdateCounts(const K* keyView, size_t numKeys, cudaStream_t s)
    {
        Comm& comm          = *comm_;
        const int numLeaves = fTree_.numLeaves;
        // leaves outside my and my peers' ranges take their counts from the global tree (layout.hpp:59-91)
        // (a member: uploaded asynchronously, rewritten only after later synchronisations of this function's callers)
        std::vector<int>& idxFromGlob = idxFromGlob_;
        idxFromGlob.clear();
        {
            int cur = 0;
            for (auto r : peerRanges_)
            {
                if (r.first == r.second) { continue; }
                for (int i = cur; i < r.first; ++i)
                    idxFromGlob.push_back(i);
                cur = r.second;
            }
            for (int i = cur; i < numLeaves; ++i)
                idxFromGlob.push_back(i);
        }
        CSB_TRY(fLeafCounts_.resize(numLeaves, s));
        CSB_TRY(computeNodeCounts<K>(fLeaves_.p, fLeafCounts_.p, numLeaves, keyView, numKeys,
                                     std::numeric_limits<unsigned>::max(), s));
        CSB_TRY(idxBuf_.resize(std::max<size_t>(idxFromGlob.size(), 1), s));
        CSB_CHECK(cudaMemcpyAsync(idxBuf_.p, idxFromGlob.data(), idxFromGlob.size() * sizeof(int), cudaMemcpyHostToDevice,
                                  s));
        CSB_TRY(rangeCount<K>(gLeaves_.p, numGlobalLeaves_, gCountScan_.p, fLeaves_.p, idxBuf_.p, int(idxFromGlob.size()),
                              fLeafCounts_.p, s));

        CSB_TRY(fCounts_.resize(fTree_.numNodes, s));
        CSB_TRY(scatterCounts(fTree_.leafToInternalLeaves(), numLeaves, fLeafCounts_.p, fCounts_.p, s));
        CSB_TRY(upsweepSum(KeyTraits<K>::maxLevel, fTree_.levelRangeHost.data(), fTree_.childOffsets.p, fCounts_.p, s));

        // peerExchange (exchangeTreeletGeneral, exchange_focus.hpp:310-366): counts of the peers' treelet nodes go out,
        // the owners' counts of my leaves in their domains come in
        size_t sendTotal = size_t(treeletOffsets_.back()), recvTotal = 0;
        for (int p : extPeers_)
            recvTotal += size_t(fAssign_[p].second - fAssign_[p].first);
        CSB_TRY(peerBuf_.resize(std::max<size_t>(sendTotal + recvTotal, 1), s));
        std::vector<CommMessage> sends, recvs;
        for (int p : intPeers_)
        {
            int n = treeletOffsets_[p + 1] - treeletOffsets_[p];
            CSB_TRY(gatherU32(treeletIdx_.p + treeletOffsets_[p], n, fCounts_.p, peerBuf_.p + treeletOffsets_[p], s));
            sends.push_back({p, peerBuf_.p + treeletOffsets_[p], size_t(n) * sizeof(uint32_t)});
        }
        size_t off = sendTotal;
        for (int p : extPeers_)
        {
            size_t n = size_t(fAssign_[p].second - fAssign_[p].first);
            recvs.push_back({p, peerBuf_.p + off, n * sizeof(uint32_t)});
            off += n;
        }
        CSB_TRY(comm.exchange(sends, recvs, s));
        off = sendTotal;
        for (int p : extPeers_)
        {
            int n = fAssign_[p].second - fAssign_[p].first;
            CSB_TRY(scatterU32(fTree_.leafToInternalLeaves() + fAssign_[p].first, n, peerBuf_.p + off, fCounts_.p, s));
            off += size_t(n);
        }
        CSB_TRY(upsweepSum(KeyTraits<K>::maxLevel, fTree_.levelRangeHost.data(), fTree_.childOffsets.p, fCounts_.p, s));
        CSB_TRY(gatherU32(fTree_.leafToInternalLeaves(), numLeaves, fCounts_.p, fLeafCounts_.p, s));
        return 0;
    }

    //! FocusedOctree::discoverHalos (octree_focus_mpi.hpp:511-568); the assigned particles are at offset 0 of sx_..
    int letDiscoverHalos(cudaStream_t s)
    {
        const int me        = comm_->rank();
        const int firstNode = fAssign_[me].first, lastNode = fAssign_[me].second;
        const int numLeaves = fTree_.numLeaves;
        CSB_TRY(macs_.resize(fTree_.numNodes, s));
        CSB_TRY(searchCenters_.resize(size_t(3) * numLeaves, s));
        CSB_TRY(searchSizes_.resize(size_t(3) * numLeaves, s));
        CSB_TRY(gatherVec3<T>(fTree_.leafToInternalLeaves(), numLeaves, geoCenters_.p, searchCenters_.p, s));
        CSB_TRY(layout_.resize(size_t(numLeaves) + 1, s));
        // layout[firstNode] = 0, inclusive scan of the own leaf counts behind it
        size_t cnt = size_t(lastNode - firstNode);
        CSB_CHECK(cudaMemcpyAsync(layout_.p + firstNode, fLeafCounts_.p + firstNode, cnt * sizeof(uint32_t),
                                  cudaMemcpyDeviceToDevice, s));
        CSB_CHECK(cudaMemsetAsync(layout_.p + lastNode, 0, sizeof(uint32_t), s));
        CSB_TRY(scanTmp_.resize(scanTempBytes(cnt + 1), s));
        CSB_TRY(exclusiveScanU32(layout_.p + firstNode, layout_.p + firstNode, cnt + 1, scanTmp_.p, s));
        // Th(2 * searchExtFact) with a float factor (octree_focus_mpi.hpp:542)
        CSB_TRY((computeBoundingBoxes<T, T>(sx_.p, sy_.p, sz_.p, sh_.p, layout_.p, firstNode, lastNode,
                                            T(2 * haloFactor_), searchCenters_.p, searchSizes_.p, s)));
        CSB_CHECK(cudaMemsetAsync(macs_.p, 0, size_t(fTree_.numNodes), s));
        return findHalos<K, T>(fTree_.prefixes.p, fTree_.childOffsets.p, fTree_.parents.p, geoCenters_.p, geoSizes_.p,
                               fLeaves_.p, searchCenters_.p, searchSizes_.p, focusLim_, bnd_, firstNode, lastNode,
                               macs_.p, s);
    }

    /*! FocusedOctree::computeLayout (octree_focus_mpi.hpp:570-582) + checkLayout (domain/layout.hpp:187-219) + the
     *  first half of Halos::exchangeRequests: the layout is scanned, checked and cut into runs of halo leaves on the
     *  device; the host reads back 4 P + 3 integers (layout and run numbers at the rank boundaries, status flags) */
    int letComputeLayout(int* fail, cudaStream_t s)
    {
        const int P         = comm_->size();
        const int me        = comm_->rank();
        const int numLeaves = fTree_.numLeaves;
        CSB_TRY(layoutCounts(fLeafCounts_.p, macs_.p, fTree_.leafToInternalLeaves(), numLeaves, fAssign_[me].first,
                             fAssign_[me].second, layout_.p, s));
        CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(numLeaves) + 1), s));
        CSB_TRY(exclusiveScanU32(layout_.p, layout_.p, size_t(numLeaves) + 1, scanTmp_.p, s));

        int* statusDev = letInts_.p + LET_STATUS;
        CSB_CHECK(cudaMemsetAsync(statusDev, 0, 2 * sizeof(int), s));
        CSB_TRY(runStarts_.resize(size_t(numLeaves) + 1, s));
        CSB_TRY(haloRunStarts(layout_.p, numLeaves, fAssignDev_.p, P, me, 512u * bucketFocus_, runStarts_.p, statusDev,
                              s));
        CSB_TRY(exclusiveScanU32(runStarts_.p, runStarts_.p, size_t(numLeaves) + 1, scanTmp_.p, s));

        // layout and run numbers at every rank's first and last leaf, plus the total
        std::vector<int> idx(size_t(2) * P + 1);
        for (int r = 0; r < P; ++r)
        {
            idx[2 * r]     = fAssign_[r].first;
            idx[2 * r + 1] = fAssign_[r].second;
        }
        idx[2 * P]        = numLeaves;
        const int nIdx    = 2 * P + 1;
        CSB_TRY(pickBuf_.resize(size_t(3) * nIdx, s));
        int* idxDev = reinterpret_cast<int*>(pickBuf_.p);
        CSB_CHECK(cudaMemcpyAsync(idxDev, idx.data(), nIdx * sizeof(int), cudaMemcpyHostToDevice, s));
        CSB_TRY(pickU32(layout_.p, idxDev, nIdx, pickBuf_.p + nIdx, s));
        CSB_TRY(pickU32(runStarts_.p, idxDev, nIdx, pickBuf_.p + 2 * nIdx, s));
        std::vector<uint32_t> picked(size_t(2) * nIdx);
        int status[2];
        CSB_CHECK(cudaMemcpyAsync(picked.data(), pickBuf_.p + nIdx, picked.size() * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaMemcpyAsync(status, statusDev, sizeof(status), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        layoutAt_.assign(picked.begin(), picked.begin() + nIdx);
        runsAt_.assign(picked.begin() + nIdx, picked.end());
        *fail = status[1] ? -1 : (status[0] ? 1 : 0);
        return 0;
    }

    //! Halos::exchangeRequests (halos/halos.hpp:64-80, halos/halo_peers.hpp:20-48, domain/exchange_keys.hpp:45-99)
    int haloExchangeRequests(cudaStream_t s)
    {
        Comm& comm          = *comm_;
        const int P         = comm.size();
        const int me        = comm.rank();
        const int numLeaves = fTree_.numLeaves;
        std::vector<int> extFlags(P, 0), allFlags(size_t(P) * P);
        for (int r = 0; r < P; ++r)
            if (r != me) { extFlags[r] = layoutAt_[2 * r + 1] > layoutAt_[2 * r] ? 1 : 0; }
        CSB_TRY(comm.allgatherHost(extFlags.data(), P * sizeof(int), allFlags.data(), s));
        haloExtPeers_.clear();
        haloIntPeers_.clear();
        for (int r = 0; r < P; ++r)
        {
            if (extFlags[r]) { haloExtPeers_.push_back(r); }
            if (allFlags[size_t(r) * P + me]) { haloIntPeers_.push_back(r); }
        }
        // request keys: one (first key, end key) pair per run of consecutive halo leaves in a peer's range
        const uint32_t totalRuns = runsAt_[2 * P];
        CSB_TRY(reqKeys_.resize(std::max<size_t>(size_t(2) * totalRuns, 1), s));
        CSB_TRY(haloRequestKeys<K>(layout_.p, numLeaves, fAssignDev_.p, P, me, runStarts_.p, fLeaves_.p, reqKeys_.p, s));
        std::vector<const void*> sendPtr(P, nullptr);
        std::vector<size_t> sendBytes(P, 0), recvOff, recvBytes;
        for (int p : haloExtPeers_)
        {
            sendPtr[p]   = reqKeys_.p + size_t(2) * runsAt_[2 * p];
            sendBytes[p] = size_t(2) * (runsAt_[2 * p + 1] - runsAt_[2 * p]) * sizeof(K);
        }
        CSB_TRY(alltoallvDevice(sendPtr, sendBytes, reqRecv_, recvOff, recvBytes, s));

        // requested keys -> outgoing index ranges of my layout, as the [scan | start] tables gatherRanges reads
        outTableOff_.assign(P, 0);
        outNumRanges_.assign(P, 0);
        outTotals_.assign(P, 0);
        size_t tableSize = 0;
        for (int p : haloIntPeers_)
        {
            outNumRanges_[p] = int(recvBytes[p] / (2 * sizeof(K)));
            outTableOff_[p]  = tableSize;
            tableSize += size_t(2) * outNumRanges_[p] + 1;
        }
        CSB_TRY(rangeTables_.resize(std::max<size_t>(tableSize, 1), s));
        std::vector<uint32_t> totals(P, 0);
        CSB_TRY(pickBuf_.resize(std::max<size_t>(size_t(3) * (2 * P + 1), size_t(P)), s, true));
        for (int p : haloIntPeers_)
        {
            const int nr   = outNumRanges_[p];
            uint32_t* scan = rangeTables_.p + outTableOff_[p]; // nr + 1 entries, then nr starts
            CSB_TRY(haloRanges<K>(reinterpret_cast<const K*>(reqRecv_.p + recvOff[p]), nr, fLeaves_.p, numLeaves + 1,
                                  layout_.p, scan, scan + nr + 1, s));
            CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(nr) + 1), s));
            CSB_TRY(exclusiveScanU32(scan, scan, size_t(nr) + 1, scanTmp_.p, s));
            CSB_CHECK(cudaMemcpyAsync(pickBuf_.p + p, scan + nr, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
        if (!haloIntPeers_.empty())
        {
            CSB_CHECK(cudaMemcpyAsync(totals.data(), pickBuf_.p, P * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaStreamSynchronize(s));
        }
        for (int p : haloIntPeers_)
            outTotals_[p] = totals[p];
        incoming_.assign(P, {0, 0});
        for (int p : haloExtPeers_)
            incoming_[p] = {layoutAt_[2 * p], layoutAt_[2 * p + 1]};
        return 0;
    }

    /*! haloExchangeGpu (halos/exchange_halos_gpu.cuh:34-119) for x, y, z, h: outgoing index ranges are packed per peer
     *  into [x | y | z | h] blocks, incoming halos of a peer are contiguous and land directly in the arrays */
    int exchangeHalos(cudaStream_t s)
    {
        Comm& comm  = *comm_;
        const int P = comm.size();
        auto blockElems = [](size_t c) { return (c + 3) & ~size_t(3); };
        size_t sendTotal = 0;
        for (int p = 0; p < P; ++p)
            sendTotal += 4 * blockElems(outTotals_[p]);
        CSB_TRY(sendBuf_.resize(std::max<size_t>(sendTotal, 1), s));
        std::vector<CommMessage> sends, recvs;
        size_t off = 0;
        for (int p = 0; p < P; ++p)
        {
            if (outTotals_[p] == 0) { continue; }
            const int nr         = outNumRanges_[p];
            const uint32_t* scan = rangeTables_.p + outTableOff_[p];
            size_t be            = blockElems(outTotals_[p]);
            CSB_TRY(gatherRanges4<T>(scan, scan + nr + 1, nr, uint32_t(outTotals_[p]), x_.p, y_.p, z_.p, h_.p,
                                     sendBuf_.p + off, be, s));
            for (int k = 0; k < 4; ++k)
                sends.push_back({p, sendBuf_.p + off + k * be, outTotals_[p] * sizeof(T)});
            off += 4 * be;
        }
        T* arrays[4] = {x_.p, y_.p, z_.p, h_.p};
        for (int p = 0; p < P; ++p)
        {
            size_t c = size_t(incoming_[p].second - incoming_[p].first);
            if (c == 0) { continue; }
            for (int k = 0; k < 4; ++k)
                recvs.push_back({p, arrays[k] + incoming_[p].first, c * sizeof(T)});
        }
        return comm.exchange(sends, recvs, s);
    }

    /*! Domain::exchangeHalos (domain/domain.hpp:332-337) for client fields: every array holds nParticlesWithHalos
     *  elements of elemBytes[k] bytes (a multiple of 4); the halo elements are overwritten with the owners' values
     *  through the send / receive pattern recorded by the last sync (haloExchangeGpu,
     *  halos/exchange_halos_gpu.cuh:34-119) */
    int exchangeHaloFields(void* const* arrays, const int* elemBytes, int numArrays, cudaStream_t s) override
    {
        CSB_REQUIRE(!firstCall_, "exchangeHalos needs a synchronised domain");
        Comm& comm  = *comm_;
        const int P = comm.size();
        if (P == 1 || numArrays == 0) { return 0; }
        size_t bytesPerParticle = 0;
        for (int k = 0; k < numArrays; ++k)
        {
            CSB_REQUIRE(arrays[k] != nullptr && elemBytes[k] > 0 && elemBytes[k] % 4 == 0,
                        "exchangeHalos: element sizes must be positive multiples of 4 bytes");
            bytesPerParticle += size_t(elemBytes[k]);
        }
        auto pad16 = [](size_t b) { return (b + 15) & ~size_t(15); };
        size_t sendTotal = 0;
        for (int p = 0; p < P; ++p)
            for (int k = 0; k < numArrays; ++k)
                sendTotal += pad16(outTotals_[p] * size_t(elemBytes[k]));
        CSB_TRY(fieldSendBuf_.resize(std::max<size_t>(sendTotal, 16), s));
        std::vector<CommMessage> sends, recvs;
        size_t off = 0;
        for (int p = 0; p < P; ++p)
        {
            if (outTotals_[p] == 0) { continue; }
            const int nr         = outNumRanges_[p];
            const uint32_t* scan = rangeTables_.p + outTableOff_[p];
            for (int k = 0; k < numArrays; ++k)
            {
                size_t bytes = outTotals_[p] * size_t(elemBytes[k]);
                CSB_TRY(gatherRangesWords(scan, scan + nr + 1, nr, uint32_t(outTotals_[p]), elemBytes[k] / 4, arrays[k],
                                          fieldSendBuf_.p + off, s));
                sends.push_back({p, fieldSendBuf_.p + off, bytes});
                off += pad16(bytes);
            }
        }
        for (int p = 0; p < P; ++p)
        {
            size_t c = size_t(incoming_[p].second - incoming_[p].first);
            if (c == 0) { continue; }
            for (int k = 0; k < numArrays; ++k)
                recvs.push_back({p, static_cast<char*>(arrays[k]) + size_t(incoming_[p].first) * elemBytes[k],
                                 c * size_t(elemBytes[k])});
        }
        CSB_TRY(comm.exchange(sends, recvs, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        return 0;
    }

    /*! Domain::reapplySync (domain/domain.hpp:297-329): send further fields of the particles through the exchange that
     *  the last sync performed.  before[k] holds the field in the particle order and buffer layout the caller had when
     *  it called sync (replayInfo()[0] elements); after[k] (replayInfo()[1] elements, the current buffer size) receives
     *  the values of the assigned particles at [startIndex, endIndex) in their new order, halos stay untouched.
     *  The reference replays redoExchange (assignment.hpp:206-215) into the enlarged arrays and then gathers through
     *  the stored ordering; here the incoming blocks land in a staging buffer and one kernel reads either side. */
    int reapplySync(const void* const* before, void* const* after, const int* elemBytes, int numArrays,
                    cudaStream_t s) override
    {
        CSB_REQUIRE(!firstCall_ && log_.valid, "reapplySync needs a completed sync to replay");
        Comm& comm   = *comm_;
        const int P  = comm.size();
        const int me = comm.rank();
        for (int k = 0; k < numArrays; ++k)
            CSB_REQUIRE(before[k] != nullptr && after[k] != nullptr && elemBytes[k] > 0 && elemBytes[k] % 4 == 0,
                        "reapplySync: element sizes must be positive multiples of 4 bytes");
        auto pad16 = [](size_t b) { return (b + 15) & ~size_t(15); };
        for (int k = 0; k < numArrays; ++k)
        {
            const size_t eb = size_t(elemBytes[k]);
            const int words = elemBytes[k] / 4;
            size_t recvOff  = 0;
            if (P > 1)
            {
                size_t sendTotal = 0;
                for (int r = 0; r < P; ++r)
                    if (r != me) { sendTotal += pad16(size_t(log_.sendIdx[r + 1] - log_.sendIdx[r]) * eb); }
                recvOff = sendTotal;
                CSB_TRY(fieldSendBuf_.resize(std::max<size_t>(sendTotal + pad16(size_t(log_.numRecv) * eb), 16), s));
                std::vector<CommMessage> sends, recvs;
                size_t off = 0;
                for (int r = 0; r < P; ++r)
                {
                    const uint32_t c = r == me ? 0 : log_.sendIdx[r + 1] - log_.sendIdx[r];
                    if (c == 0) { continue; }
                    CSB_TRY(gatherWords(ordering_.p + log_.prevStart + log_.sendIdx[r], c, words, before[k],
                                        fieldSendBuf_.p + off, s));
                    sends.push_back({r, fieldSendBuf_.p + off, size_t(c) * eb});
                    off += pad16(size_t(c) * eb);
                }
                size_t received = 0;
                for (int r = 0; r < P; ++r)
                {
                    const size_t c = log_.recvCounts[r];
                    if (c == 0 || r == me) { continue; }
                    recvs.push_back({r, fieldSendBuf_.p + recvOff + received * eb, c * eb});
                    received += c;
                }
                CSB_REQUIRE(received == log_.numRecv, "reapplySync: the recorded exchange is inconsistent");
                CSB_TRY(comm.exchange(sends, recvs, s));
            }
            CSB_TRY(replayGatherWords(log_.order, log_.numAssigned, words, before[k], fieldSendBuf_.p + recvOff,
                                      log_.recvStart, log_.numRecv, static_cast<char*>(after[k]) + size_t(start_) * eb,
                                      s));
            // the staging buffer is reused by the next field
            CSB_CHECK(cudaStreamSynchronize(s));
        }
        return 0;
    }

    //! {elements of the arrays before the sync, elements after, startIndex, endIndex}
    int replayInfo(uint64_t* info) const override
    {
        CSB_REQUIRE(!firstCall_ && log_.valid, "reapplySync needs a completed sync to replay");
        info[0] = log_.prevSize;
        info[1] = bufSize_;
        info[2] = start_;
        info[3] = end_;
        return 0;
    }

    int attachComm(Comm* c) override
    {
        CSB_REQUIRE(c != nullptr, "null communicator");
        CSB_REQUIRE(firstCall_, "the communicator must be attached before the first sync");
        CSB_REQUIRE(c->size() == numRanks_ && c->rank() == rank_, "communicator rank/size differ from the domain's");
        CSB_REQUIRE(c->size() <= LET_ERROR, "at most 120 ranks per domain");
        comm_ = c;
        return 0;
    }

    int info(uint64_t* out, double* box) const override
    {
        out[0] = start_;
        out[1] = end_;
        out[2] = bufSize_;
        out[3] = uint64_t(fTree_.numLeaves);
        out[4] = uint64_t(fTree_.numNodes);
        out[5] = uint64_t(numGlobalLeaves_);
        out[6] = uint64_t(gTree_.numNodes);
        out[7] = uint64_t(KeyTraits<K>::maxLevel);
        std::copy(lim_, lim_ + 6, box);
        return 0;
    }

    void* ptr(int field) override
    {
        switch (field)
        {
            case CS_FIELD_X: return x_.p;
            case CS_FIELD_Y: return y_.p;
            case CS_FIELD_Z: return z_.p;
            case CS_FIELD_H: return h_.p;
            case CS_FIELD_KEYS: return keys_.p;
            case CS_FIELD_FOCUS_LEAVES: return fLeaves_.p;
            case CS_FIELD_FOCUS_LEAF_COUNTS: return fLeafCounts_.p;
            case CS_FIELD_FOCUS_NODE_COUNTS: return fCounts_.p;
            case CS_FIELD_LAYOUT: return layout_.p;
            case CS_FIELD_PREFIXES: return fTree_.prefixes.p;
            case CS_FIELD_CHILD_OFFSETS: return fTree_.childOffsets.p;
            case CS_FIELD_PARENTS: return fTree_.parents.p;
            case CS_FIELD_LEVEL_RANGE: return fTree_.levelRange.p;
            case CS_FIELD_INTERNAL_TO_LEAF: return fTree_.internalToLeaf.p;
            case CS_FIELD_LEAF_TO_INTERNAL: return fTree_.leafToInternal.p;
            case CS_FIELD_GEO_CENTERS: return geoCenters_.p;
            case CS_FIELD_GEO_SIZES: return geoSizes_.p;
            case CS_FIELD_HALO_FLAGS: return macs_.p;
            case CS_FIELD_GLOBAL_LEAVES: return gLeaves_.p;
            case CS_FIELD_GLOBAL_COUNTS: return gCounts_.p;
            case CS_FIELD_GLOBAL_PREFIXES: return gTree_.prefixes.p;
            case CS_FIELD_GLOBAL_CHILD_OFFSETS: return gTree_.childOffsets.p;
            default: return nullptr;
        }
    }

    int neighbors(uint32_t ngmax, uint32_t* nb, uint32_t* nc, cudaStream_t s) override
    {
        CSB_REQUIRE(!firstCall_, "findNeighbors needs a synchronised domain");
        return findNeighbors<T, T>(x_.p, y_.p, z_.p, h_.p, start_, end_, lim_, bnd_, fTree_.numLeaves,
                                fTree_.childOffsets.p,
                                fTree_.parents.p, fTree_.internalToLeaf.p, layout_.p, geoCenters_.p, geoSizes_.p,
                                ngmax, nb, nc, s);
    }

    int download(void* x, void* y, void* z, void* h, void* keys, cudaStream_t s) override
    {
        size_t n = bufSize_;
        if (x) { CSB_CHECK(cudaMemcpyAsync(x, x_.p, n * sizeof(T), cudaMemcpyDeviceToHost, s)); }
        if (y) { CSB_CHECK(cudaMemcpyAsync(y, y_.p, n * sizeof(T), cudaMemcpyDeviceToHost, s)); }
        if (z) { CSB_CHECK(cudaMemcpyAsync(z, z_.p, n * sizeof(T), cudaMemcpyDeviceToHost, s)); }
        if (h) { CSB_CHECK(cudaMemcpyAsync(h, h_.p, n * sizeof(T), cudaMemcpyDeviceToHost, s)); }
        if (keys) { CSB_CHECK(cudaMemcpyAsync(keys, keys_.p, n * sizeof(K), cudaMemcpyDeviceToHost, s)); }
        return 0;
    }

private:
    //! CSB_TRACE=1: wall time of each phase of sync (stream synchronised at the phase boundaries) on stderr
    void phase(const char* name, cudaStream_t s)
    {
        if (!trace_) { return; }
        cudaStreamSynchronize(s);
        auto now = std::chrono::steady_clock::now();
        if (name)
        {
            fprintf(stderr, "[csb rank %d] %-28s %9.3f ms\n", rank_, name,
                    std::chrono::duration<double, std::milli>(now - lastPhase_).count());
        }
        lastPhase_ = now;
    }
    bool trace_{std::getenv("CSB_TRACE") != nullptr};
    std::chrono::steady_clock::time_point lastPhase_;

    /* ------------------------------------------------------------ box: makeGlobalBox (sfc/box_mpi.hpp:51-105) +
     *                                                              limitBoxShrinking (sfc/box.hpp:398-415) */
    int updateBox(size_t numPart, cudaStream_t s)
    {
        bool keep[3];
        bool anyOpen = false;
        for (int d = 0; d < 3; ++d)
        {
            keep[d] = bnd_[d] == 1 || bnd_[d] == 2;
            anyOpen = anyOpen || !keep[d];
        }
        T prev[6];
        for (int i = 0; i < 6; ++i)
            prev[i] = T(lim_[i]);
        T ext[6];
        std::copy(prev, prev + 6, ext);
        if (numPart && anyOpen)
        {
            constexpr int blocks = 592;
            CSB_TRY(partials_.resize(size_t(3) * 2 * blocks, s));
            const T* arrays[3] = {x_.p + start_, y_.p + start_, z_.p + start_};
            for (int d = 0; d < 3; ++d)
                if (!keep[d]) { CSB_TRY(minMaxPartials<T>(arrays[d], numPart, partials_.p + d * 2 * blocks, blocks, s)); }
            std::vector<T> hostPartials(size_t(3) * 2 * blocks);
            CSB_CHECK(cudaMemcpyAsync(hostPartials.data(), partials_.p, hostPartials.size() * sizeof(T),
                                      cudaMemcpyDeviceToHost, s));
            CSB_CHECK(cudaStreamSynchronize(s));
            for (int d = 0; d < 3; ++d)
            {
                if (keep[d]) { continue; }
                T mn = hostPartials[d * 2 * blocks], mx = hostPartials[d * 2 * blocks + 1];
                for (int b = 1; b < blocks; ++b)
                {
                    mn = std::min(mn, hostPartials[d * 2 * blocks + 2 * b]);
                    mx = std::max(mx, hostPartials[d * 2 * blocks + 2 * b + 1]);
                }
                ext[2 * d]     = mn;
                ext[2 * d + 1] = mx;
            }
        }
        if (comm_->size() > 1 && anyOpen)
        {
            // MPI_Allreduce(MIN) over {min, -max} of sfc/box_mpi.hpp:78-86; ranks without particles contribute the
            // previous limits, as in the reference
            std::vector<T> all(size_t(6) * comm_->size());
            CSB_TRY(comm_->allgatherHost(ext, sizeof(ext), all.data(), s));
            for (int r = 0; r < comm_->size(); ++r)
                for (int d = 0; d < 3; ++d)
                {
                    ext[2 * d]     = std::min(ext[2 * d], all[6 * r + 2 * d]);
                    ext[2 * d + 1] = std::max(ext[2 * d + 1], all[6 * r + 2 * d + 1]);
                }
        }
        const T maxSide = std::max({ext[1] - ext[0], ext[3] - ext[2], ext[5] - ext[4]});
        for (int d = 0; d < 3; ++d)
            if (bnd_[d] == 3) { ext[2 * d + 1] = std::max(ext[2 * d + 1], ext[2 * d] + maxSide); }

        if (!firstCall_)
        {
            const T shrink = T(0.05);
            for (int d = 0; d < 3; ++d)
            {
                T len          = prev[2 * d + 1] - prev[2 * d];
                ext[2 * d]     = std::min(ext[2 * d], prev[2 * d] + shrink * len);
                ext[2 * d + 1] = std::max(ext[2 * d + 1], prev[2 * d + 1] - shrink * len);
            }
        }
        for (int i = 0; i < 6; ++i)
            lim_[i] = double(ext[i]);
        return 0;
    }

    /* ------------------------------------------------------------ exchangeParticlesGpu
     *  (domain/domaindecomp_mpi_gpu.cuh:70-166).  Outgoing ranges are gathered through the ordering (from the particle
     *  records, recBuf_) into one packed
     *  buffer per destination ([x | y | z | h], each block 16-byte aligned); incoming blocks land directly in the
     *  particle arrays at [recvStart, recvStart + numRecv), sources in ascending rank order (the reference takes
     *  them in arrival order, domaindecomp_mpi.hpp:116-140; the stable key sort that follows makes the final order
     *  independent of it for distinct keys).  Keys are not sent, the receiver recomputes them (assignment.hpp:197). */
    int exchangeParticles(const std::vector<uint32_t>& sendIdx, LocalIndex recvStart, LocalIndex numRecv,
                          std::vector<uint32_t>& recvCounts, cudaStream_t s)
    {
        Comm& comm   = *comm_;
        const int P  = comm.size();
        const int me = comm.rank();

        std::vector<uint32_t> sendCounts(P, 0), allCounts(size_t(P) * P);
        for (int r = 0; r < P; ++r)
            sendCounts[r] = r == me ? 0 : sendIdx[r + 1] - sendIdx[r];
        CSB_TRY(comm.allgatherHost(sendCounts.data(), P * sizeof(uint32_t), allCounts.data(), s));
        recvCounts.assign(P, 0);
        for (int r = 0; r < P; ++r)
            if (r != me) { recvCounts[r] = allCounts[size_t(r) * P + me]; }

        /* Peer-memory path: the pack kernel of every destination gathers through the ordering and stores straight into
         * the destination rank's particle arrays over NVLink (CUDA IPC mappings, cached) - pack and transfer are one
         * kernel and nothing is staged.  Source r's block lands behind the blocks of the lower ranks, exactly where
         * the send/recv path below puts it. */
        if (peerPush_)
        {
            void* mine[4] = {x_.p, y_.p, z_.p, h_.p};
            std::vector<void*> peers;
            std::vector<uint64_t> recvStarts;
            x_.sharedWithPeers = y_.sharedWithPeers = z_.sharedWithPeers = h_.sharedWithPeers = true;
            phase("  xp counts", s);
            int st = comm.sharePointers(mine, 4, uint64_t(recvStart), peers, recvStarts, s);
            phase("  xp sharePointers", s);
            if (st == 2) { peerPush_ = false; }
            else if (st != 0) { return st; }
            else
            {
                size_t received = 0;
                for (int r = 0; r < P; ++r)
                    if (r != me) { received += allCounts[size_t(r) * P + me]; }
                CSB_REQUIRE(received == numRecv,
                            "exchangeParticles: incoming particle count does not match the assignment");
                // one pack kernel per destination, all in flight together on side streams (a single kernel does not
                // keep enough bytes in flight to fill an NVLink direction); destinations in the order me+1, me+2, ...
                // so that the senders start with different targets
                if (pushStreams_.empty())
                {
                    pushStreams_.resize(4);
                    for (auto& ps : pushStreams_)
                        CSB_CHECK(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
                    CSB_CHECK(cudaEventCreateWithFlags(&pushReady_, cudaEventDisableTiming));
                    pushDone_.resize(pushStreams_.size());
                    for (auto& e : pushDone_)
                        CSB_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                }
                CSB_CHECK(cudaEventRecord(pushReady_, s)); // ordering and particle arrays are final on s
                for (auto& ps : pushStreams_)
                    CSB_CHECK(cudaStreamWaitEvent(ps, pushReady_, 0));
                int launched = 0;
                for (int k = 1; k < P; ++k)
                {
                    const int r = (me + k) % P;
                    size_t c    = sendCounts[r];
                    if (c == 0) { continue; }
                    size_t dstOffset = size_t(recvStarts[r]);
                    for (int src = 0; src < me; ++src)
                        if (src != r) { dstOffset += allCounts[size_t(src) * P + r]; }
                    void* dst4[4];
                    for (int k = 0; k < 4; ++k)
                        dst4[k] = static_cast<T*>(peers[size_t(r) * 4 + k]) + dstOffset;
                    cudaStream_t ps = pushStreams_[size_t(launched++) % pushStreams_.size()];
                    CSB_TRY(gatherFromRecords4(ordering_.p + start_ + sendIdx[r], c, recBuf_.p, dst4, int(sizeof(T)),
                                               ps));
                    comm.bytesSent += 4 * c * sizeof(T);
                }
                for (size_t q = 0; q < pushStreams_.size(); ++q)
                {
                    CSB_CHECK(cudaEventRecord(pushDone_[q], pushStreams_[q]));
                    CSB_CHECK(cudaStreamWaitEvent(s, pushDone_[q], 0));
                }
                CSB_CHECK(cudaStreamSynchronize(s)); // my stores have landed ...
                phase("  xp push", s);
                int bst = comm.barrier(s); // ... and so have everybody else's
                phase("  xp barrier", s);
                return bst;
            }
        }

        auto blockElems = [](size_t c) { return (c + 3) & ~size_t(3); }; // 16-byte multiples for 4- and 8-byte reals
        size_t totalSend = 0;
        for (int r = 0; r < P; ++r)
            totalSend += 4 * blockElems(sendCounts[r]);
        CSB_TRY(sendBuf_.resize(std::max<size_t>(totalSend, 1), s));

        std::vector<CommMessage> sends, recvs;
        size_t off = 0;
        for (int r = 0; r < P; ++r)
        {
            size_t c = sendCounts[r];
            if (c == 0) { continue; }
            size_t be          = blockElems(c);
            void* dst[4]       = {sendBuf_.p + off, sendBuf_.p + off + be, sendBuf_.p + off + 2 * be,
                                  sendBuf_.p + off + 3 * be};
            CSB_TRY(gatherFromRecords4(ordering_.p + start_ + sendIdx[r], c, recBuf_.p, dst, int(sizeof(T)), s));
            for (int k = 0; k < 4; ++k)
                sends.push_back({r, dst[k], c * sizeof(T)});
            off += 4 * be;
        }
        size_t received = 0;
        T* arrays[4]    = {x_.p, y_.p, z_.p, h_.p};
        for (int r = 0; r < P; ++r)
        {
            size_t c = allCounts[size_t(r) * P + me];
            if (c == 0 || r == me) { continue; }
            for (int k = 0; k < 4; ++k)
                recvs.push_back({r, arrays[k] + recvStart + received, c * sizeof(T)});
            received += c;
        }
        CSB_REQUIRE(received == numRecv, "exchangeParticles: incoming particle count does not match the assignment");
        CSB_TRY(comm.exchange(sends, recvs, s));
        return 0;
    }

    //! iotaStart >= 0: values are produced as the sorting permutation of iotaStart, iotaStart + 1, ... (sequence + sort)
    int sortPairs(K* keys, uint32_t* values, size_t n, cudaStream_t s, long long iotaStart = -1)
    {
        CSB_TRY(keyBuf_.resize(std::max<size_t>(n, 1), s));
        CSB_TRY(valueBuf_.resize(std::max<size_t>(n, 1), s));
        size_t tb = sortTempBytesT<K>(n);
        CSB_TRY(sortTmp_.resize(tb, s));
        if (iotaStart >= 0 && n > 0)
        {
            return sortIotaDispatch(keys, values, uint32_t(iotaStart), n, keyBuf_.p, valueBuf_.p, sortTmp_.p, tb, s);
        }
        return sortDispatch(keys, values, n, keyBuf_.p, valueBuf_.p, sortTmp_.p, tb, s);
    }

    int linkTree(DevBuf<K>& leaves, int numLeaves, OctreeBufs<K>& tree, cudaStream_t s)
    {
        CSB_TRY(tree.resize(numLeaves, s));
        size_t tb = linkTempBytesT<K>(numLeaves);
        CSB_TRY(linkTmp_.resize(tb, s));
        CSB_TRY(buildOctree<K>(leaves.p, numLeaves, tree.prefixes.p, tree.childOffsets.p, tree.parents.p,
                               tree.levelRange.p, tree.internalToLeaf.p, tree.leafToInternal.p, linkTmp_.p, tb, s));
        CSB_CHECK(cudaMemcpyAsync(tree.levelRangeHost.data(), tree.levelRange.p,
                                  tree.levelRangeHost.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        return 0;
    }

    /* ------------------------------------------------------------ updateOctreeGlobal (tree/update_mpi.hpp:64-97) */
    int updateGlobalTree(const K* keys, size_t n, unsigned* maxCountOut, cudaStream_t s)
    {
        int newNumLeaves = 0, converged = 0;
        CSB_TRY(nodeOps_.resize(size_t(numGlobalLeaves_) + 1, s));
        CSB_TRY(opsTmp_.resize(nodeOpsTempBytes(numGlobalLeaves_), s));
        CSB_TRY(computeNodeOps<K>(gLeaves_.p, numGlobalLeaves_, gCounts_.p, bucket_, nodeOps_.p, opsTmp_.p,
                                  &newNumLeaves, &converged, s));
        CSB_TRY(gLeavesAlt_.resize(size_t(newNumLeaves) + 1, s));
        CSB_TRY(rebalanceTree<K>(gLeaves_.p, numGlobalLeaves_, newNumLeaves, nodeOps_.p, gLeavesAlt_.p, s));
        gLeaves_.swap(gLeavesAlt_);
        numGlobalLeaves_ = newNumLeaves;
        CSB_TRY(gCounts_.resize(newNumLeaves, s));
        CSB_TRY(computeNodeCounts<K>(gLeaves_.p, gCounts_.p, newNumLeaves, keys, n,
                                     std::numeric_limits<unsigned>::max(), s));
        if (comm_->size() > 1)
        {
            // update_mpi.hpp:86-97: counts = max(local, sum over ranks); the sum saturates nowhere below 2^32
            CSB_TRY(gCountsLocal_.resize(newNumLeaves, s));
            CSB_CHECK(cudaMemcpyAsync(gCountsLocal_.p, gCounts_.p, size_t(newNumLeaves) * sizeof(uint32_t),
                                      cudaMemcpyDeviceToDevice, s));
            CSB_TRY(comm_->allreduceSumU32(gCounts_.p, newNumLeaves, s));
            CSB_TRY(maxInto(gCounts_.p, gCountsLocal_.p, newNumLeaves, s));
        }
        if (converged)
        {
            *maxCountOut = 0;
            return 0;
        }
        CSB_TRY(maxU32(gCounts_.p, newNumLeaves, scalars_.p, s));
        CSB_CHECK(cudaMemcpyAsync(maxCountOut, scalars_.p, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        return 0;
    }

    /* ------------------------------------------------------------ FocusedOctree::updateTree
     *  = CombinedUpdate::updateFocus (focus/octree_focus.hpp:65-127) + updateGeoCenters */
    int updateFocusTree(int* convergedOut, cudaStream_t s)
    {
        const int numNodes  = fTree_.numNodes;
        const int numLeaves = fTree_.numLeaves;
        const int me       = comm_->rank();
        const K focusStart = assignment_.boundaries[me], focusEnd = assignment_.boundaries[me + 1];

        if (comm_->size() == 1)
        {
            CSB_TRY(macs_.resize(numNodes, s));
            CSB_CHECK(cudaMemsetAsync(macs_.p, 0, size_t(numNodes), s)); // irrelevant when every node is in focus
        }
        else { CSB_REQUIRE(macs_.n == size_t(numNodes), "update of criteria required before updating the tree"); }
        CSB_TRY(nodeOpsAll_.resize(numNodes, s));
        int* statusDev  = reinterpret_cast<int*>(scalars_.p + 16);
        int* changesDev = statusDev + 1;
        int* notOneDev  = statusDev + 2;
        CSB_CHECK(cudaMemsetAsync(statusDev, 0, 3 * sizeof(int), s));

        CSB_TRY(essentialOps<K>(fTree_.prefixes.p, fTree_.childOffsets.p, fTree_.parents.p, fCounts_.p, macs_.p,
                                focusStart, focusEnd, bucketFocus_, nodeOpsAll_.p, numNodes, s));
        // mandatory keys: the global leaves of the own assignment, boundaries included (octree_focus_mpi.hpp:121-122)
        const int enforceFirst = assignment_.treeOffsets[me];
        const int enforceCount = assignment_.treeOffsets[me + 1] - enforceFirst + 1;
        CSB_TRY(enforceKeys<K>(gLeaves_.p + enforceFirst, enforceCount, fTree_.prefixes.p, fTree_.childOffsets.p,
                               fTree_.parents.p, nodeOpsAll_.p, statusDev, s));
        CSB_TRY(protectAncestors<K>(fTree_.prefixes.p, fTree_.parents.p, nodeOpsAll_.p, numNodes, changesDev, s));

        CSB_TRY(nodeOps_.resize(size_t(numLeaves) + 1, s));
        CSB_TRY(gatherLeafOps(fTree_.leafToInternalLeaves(), numLeaves, nodeOpsAll_.p, nodeOps_.p, notOneDev, s));
        CSB_TRY(scanTmp_.resize(scanTempBytes(size_t(numLeaves) + 1), s));
        CSB_TRY(exclusiveScanU32(reinterpret_cast<uint32_t*>(nodeOps_.p), reinterpret_cast<uint32_t*>(nodeOps_.p),
                                 size_t(numLeaves) + 1, scanTmp_.p, s));

        int flags[3];
        int newNumLeaves = 0;
        CSB_CHECK(cudaMemcpyAsync(flags, statusDev, sizeof(flags), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaMemcpyAsync(&newNumLeaves, nodeOps_.p + numLeaves, sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        const int status = flags[0];
        bool converged   = flags[1] == 0;
        if (status == ENFORCE_CANCEL_MERGE) { converged = flags[2] == 0; }
        else if (status == ENFORCE_REBALANCE) { converged = false; }

        CSB_TRY(fLeavesAlt_.resize(size_t(newNumLeaves) + 1, s));
        CSB_TRY(rebalanceTree<K>(fLeaves_.p, numLeaves, newNumLeaves, nodeOps_.p, fLeavesAlt_.p, s));
        fLeaves_.swap(fLeavesAlt_);
        int numFocusLeaves = newNumLeaves;

        if (status == ENFORCE_FAILED)
        {
            converged = false;
            CSB_TRY(injectKeys(gLeaves_.p + enforceFirst, enforceCount, &numFocusLeaves, s));
        }

        CSB_TRY(linkTree(fLeaves_, numFocusLeaves, fTree_, s));
        *convergedOut = converged ? 1 : 0;
        phase("    rebalance+link", s);
        if (comm_->size() > 1) { CSB_TRY(letSyncWithPeers(s)); }
        phase("    letSyncWithPeers", s);
        // FocusedOctree::updateTree stores the box for all property updates until the next call and recomputes the
        // geometric centres (octree_focus_mpi.hpp:166-175)
        std::copy(lim_, lim_ + 6, focusLim_);
        return updateGeoCenters(s);
    }

    int updateGeoCenters(cudaStream_t s)
    {
        CSB_TRY(geoCenters_.resize(size_t(3) * fTree_.numNodes, s));
        CSB_TRY(geoSizes_.resize(size_t(3) * fTree_.numNodes, s));
        return computeGeoCenters<K, T>(0, fTree_.prefixes.p, fTree_.numNodes, geoCenters_.p, geoSizes_.p, focusLim_,
                                       bnd_, s);
    }

    //! focus/inject.hpp:50-84: leaves <- span(sort(leaves U keys))
    int injectKeys(const K* keys, int numKeys, int* numLeavesInOut, cudaStream_t s)
    {
        size_t total = size_t(*numLeavesInOut) + 1 + numKeys;
        CSB_TRY(fLeavesAlt_.resize(total, s));
        CSB_CHECK(cudaMemcpyAsync(fLeavesAlt_.p, fLeaves_.p, (size_t(*numLeavesInOut) + 1) * sizeof(K),
                                  cudaMemcpyDeviceToDevice, s));
        CSB_CHECK(cudaMemcpyAsync(fLeavesAlt_.p + *numLeavesInOut + 1, keys, size_t(numKeys) * sizeof(K),
                                  cudaMemcpyDeviceToDevice, s));
        CSB_TRY(keyBuf_.resize(total, s));
        size_t tb = sortTempBytesT<K>(total);
        CSB_TRY(sortTmp_.resize(tb, s));
        CSB_TRY(sortDispatch(fLeavesAlt_.p, nullptr, total, keyBuf_.p, nullptr, sortTmp_.p, tb, s));

        int numGaps = int(total) - 1;
        CSB_TRY(gapCounts_.resize(total, s));
        CSB_TRY(countGaps<K>(fLeavesAlt_.p, numGaps, gapCounts_.p, s));
        CSB_TRY(scanTmp_.resize(scanTempBytes(total), s));
        CSB_TRY(exclusiveScanU32(gapCounts_.p, gapCounts_.p, total, scanTmp_.p, s));
        uint32_t numNodesGap = 0;
        CSB_CHECK(cudaMemcpyAsync(&numNodesGap, gapCounts_.p + numGaps, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_TRY(fLeaves_.resize(size_t(numNodesGap) + 1, s));
        CSB_TRY(fillGaps<K>(fLeavesAlt_.p, numGaps, gapCounts_.p, fLeaves_.p, s));
        *numLeavesInOut = int(numNodesGap);
        return 0;
    }

    /* ------------------------------------------------------------ FocusedOctree::updateCounts
     *  (octree_focus_mpi.hpp:193-252): leaf counts, scatter to nodes, upsweep (no peers on one rank) */
    int updateFocusCounts(const K* keyView, size_t numKeys, cudaStream_t s)
    {
        CSB_TRY(fLeafCounts_.resize(fTree_.numLeaves, s));
        CSB_TRY(computeNodeCounts<K>(fLeaves_.p, fLeafCounts_.p, fTree_.numLeaves, keyView, numKeys,
                                     std::numeric_limits<unsigned>::max(), s));
        CSB_TRY(fCounts_.resize(fTree_.numNodes, s));
        CSB_TRY(scatterCounts(fTree_.leafToInternalLeaves(), fTree_.numLeaves, fLeafCounts_.p, fCounts_.p, s));
        CSB_TRY(upsweepSum(KeyTraits<K>::maxLevel, fTree_.levelRangeHost.data(), fTree_.childOffsets.p, fCounts_.p, s));
        return 0;
    }

    // multi-rank LET state (host mirrors are what the reference keeps on the host as well)
    double focusLim_[6]{0, 1, 0, 1, 0, 1};
    static constexpr int LET_INTS = 128, LET_ERROR = 120, LET_STATUS = 121; // header of letInts_: [peer flags | ...]
    std::vector<int> fLower_;                                                // findNodeAbove of the rank boundaries
    std::vector<std::pair<int, int>> fAssign_, peerRanges_;
    std::vector<int> extPeers_, intPeers_, haloExtPeers_, haloIntPeers_, globDispl_, treeletOffsets_, outNumRanges_;
    std::vector<int> fAssignFlat_, idxFromGlob_; // host sources of asynchronous uploads
    std::vector<uint32_t> layoutAt_, runsAt_; // layout / halo-run number at {first, last leaf of rank r}..., total
    std::vector<size_t> outTableOff_, outTotals_;
    std::vector<std::pair<uint32_t, uint32_t>> incoming_;
    DevBuf<int> treeletIdx_, idxBuf_, letInts_, fAssignDev_;
    DevBuf<uint64_t> gCountScan_;
    DevBuf<uint32_t> peerBuf_, rangeTables_, tlValid_, runStarts_, pickBuf_;
    DevBuf<K> tlKeys_, rejKeys_, reqKeys_;
    DevBuf<T> centers4_, searchCenters_, searchSizes_;
    DevBuf<char> tlRecv_, rejRecv_, reqRecv_, fieldSendBuf_, recBuf_;

    int rank_, numRanks_;
    // exchangeParticles through peer memory; cleared when the transport cannot map it (CSB_NO_PEER_PUSH forces the
    // send/recv path)
    bool peerPush_{std::getenv("CSB_NO_PEER_PUSH") == nullptr};
    SelfComm selfComm_;
    Comm* comm_{&selfComm_};
    SfcAssignment<K> assignment_;
    std::vector<K> gLeavesHost_;
    std::vector<uint32_t> gCountsHost_;
    unsigned bucket_, bucketFocus_;
    float theta_;
    double lim_[6];
    double lim0_[6];
    int bnd_[3];
    bool firstCall_{true};
    LocalIndex start_{0}, end_{0}, bufSize_{0};
    cudaStream_t copyStream_{nullptr}; // upload of h from host memory, concurrent with the first stages of sync
    std::vector<cudaStream_t> pushStreams_; // exchangeParticles through peer memory: pack kernels in flight together
    cudaEvent_t pushReady_{nullptr};
    std::vector<cudaEvent_t> pushDone_;
    cudaEvent_t copyReady_{nullptr}, copyDone_{nullptr};
    bool hUploadPending_{false};

    DevBuf<T> x_, y_, z_, h_, sx_, sy_, sz_, sh_, partials_, geoCenters_, geoSizes_, sendBuf_;
    DevBuf<K> keys_, keyBuf_, boundaryKeys_, assignedKeys_;
    DevBuf<uint32_t> ordering_, valueBuf_, scalars_, gapCounts_, layout_, assignedOrder_;
    //! the particle exchange of the last sync, kept for reapplySync; `order` points into ordering_ / assignedOrder_,
    //! which stay untouched until the next sync
    struct ReplayLog
    {
        bool valid{false};
        const uint32_t* order{nullptr};
        LocalIndex numAssigned{0}, prevStart{0}, prevSize{0}, recvStart{0}, numRecv{0};
        std::vector<uint32_t> sendIdx, recvCounts;
    } log_;
    DevBuf<unsigned char> sortTmp_, linkTmp_, opsTmp_, scanTmp_;
    DevBuf<int> nodeOps_, nodeOpsAll_;

    // global tree
    DevBuf<K> gLeaves_, gLeavesAlt_;
    DevBuf<uint32_t> gCounts_, gCountsLocal_;
    OctreeBufs<K> gTree_;
    int numGlobalLeaves_{0};

    // focus tree
    DevBuf<K> fLeaves_, fLeavesAlt_;
    DevBuf<uint32_t> fLeafCounts_, fCounts_;
    DevBuf<uint8_t> macs_;
    OctreeBufs<K> fTree_;
};

template<class K, class T>
cs_domain_t* createDomain(int rank, int numRanks, unsigned bucket, unsigned bucketFocus, float theta,
                          const double* lim, const int* bnd)
{
    if (numRanks < 1 || rank < 0 || rank >= numRanks)
    {
        setLastError("cs_domain_create: invalid rank / numRanks");
        return nullptr;
    }
    if (bucket < bucketFocus)
    {
        // domain.hpp:81-85
        setLastError("The bucket size of the global tree must not be smaller than the bucket size of the focused tree");
        return nullptr;
    }
    auto* d = new DomainImpl<K, T>(rank, numRanks, bucket, bucketFocus, theta, lim, bnd);
    if (d->init(nullptr) != 0)
    {
        delete d;
        return nullptr;
    }
    return reinterpret_cast<cs_domain_t*>(static_cast<DomainBase*>(d));
}

inline DomainBase* impl(cs_domain_t* d) { return reinterpret_cast<DomainBase*>(d); }

} // namespace

} // namespace csb

extern "C"
{

cs_domain_t* cs_domain_create_u32f(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd)
{
    return csb::createDomain<uint32_t, float>(rank, numRanks, bucketSize, bucketSizeFocus, theta, lim, bnd);
}
cs_domain_t* cs_domain_create_u64f(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd)
{
    return csb::createDomain<uint64_t, float>(rank, numRanks, bucketSize, bucketSizeFocus, theta, lim, bnd);
}
cs_domain_t* cs_domain_create_u64d(int rank, int numRanks, unsigned bucketSize, unsigned bucketSizeFocus, float theta,
                                   const double* lim, const int* bnd)
{
    return csb::createDomain<uint64_t, double>(rank, numRanks, bucketSize, bucketSizeFocus, theta, lim, bnd);
}

void cs_domain_destroy(cs_domain_t* d) { delete csb::impl(d); }

int cs_domain_sync(cs_domain_t* d, const void* x, const void* y, const void* z, const void* h, const void* keys,
                   size_t n, int hostInput, void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->sync(x, y, z, h, keys, n, hostInput != 0, cudaStream_t(stream));
}

int cs_domain_info(const cs_domain_t* d, uint64_t* out8, double* box6)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(const_cast<cs_domain_t*>(d))->info(out8, box6);
}

void* cs_domain_ptr(cs_domain_t* d, int field) { return d ? csb::impl(d)->ptr(field) : nullptr; }

int cs_domain_find_neighbors(cs_domain_t* d, uint32_t ngmax, uint32_t* neighbors, uint32_t* neighborsCount,
                             void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->neighbors(ngmax, neighbors, neighborsCount, cudaStream_t(stream));
}

int cs_domain_attach_comm(cs_domain_t* d, cs_comm_t* comm)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->attachComm(reinterpret_cast<csb::Comm*>(comm));
}

int cs_domain_exchange_halos(cs_domain_t* d, void* const* arrays, const int* elemBytes, int numArrays, void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->exchangeHaloFields(arrays, elemBytes, numArrays, cudaStream_t(stream));
}

int cs_domain_reapply_sync(cs_domain_t* d, const void* const* before, void* const* after, const int* elemBytes,
                           int numArrays, void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->reapplySync(before, after, elemBytes, numArrays, cudaStream_t(stream));
}

int cs_domain_set_halo_factor(cs_domain_t* d, float factor)
{
    CSB_REQUIRE(d != nullptr && factor > 0, "cs_domain_set_halo_factor: null domain or non-positive factor");
    csb::impl(d)->haloFactor_ = factor;
    return 0;
}

int cs_domain_replay_info(const cs_domain_t* d, uint64_t* info4)
{
    CSB_REQUIRE(d != nullptr && info4 != nullptr, "null argument");
    return csb::impl(const_cast<cs_domain_t*>(d))->replayInfo(info4);
}

int cs_domain_reset(cs_domain_t* d, void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->reset(cudaStream_t(stream));
}

int cs_domain_download(cs_domain_t* d, void* x, void* y, void* z, void* h, void* keys, void* stream)
{
    CSB_REQUIRE(d != nullptr, "null domain");
    return csb::impl(d)->download(x, y, z, h, keys, cudaStream_t(stream));
}

} // extern "C"
