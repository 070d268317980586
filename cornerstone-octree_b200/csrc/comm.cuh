/* Rank-to-rank plumbing of the multi-rank domain: the role MPI plays in the reference (domain/domaindecomp_mpi.hpp,
 * tree/update_mpi.hpp, halos/exchange_halos.hpp, focus/exchange_focus.hpp), reduced to the four operations the
 * hot path needs.  Two transports implement it:
 *   - NcclComm : one process per GPU; ncclAllReduce / grouped ncclSend+ncclRecv over NVLink.  libnccl is dlopen()ed so
 *                the library loads (and single-rank domains run) on machines without NCCL.
 *   - LocalComm: ranks are threads of one process (any rank -> device mapping, also all on one device); data moves
 *                with cudaMemcpyAsync between the ranks' buffers, metadata through shared host memory.  This is what
 *                lets the multi-rank parity tests run on a single-GPU box.
 * Every rank must call the collective operations in the same order (as with MPI).
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace csb
{

struct CommMessage
{
    int peer;
    void* ptr;      // device pointer
    size_t bytes;
};

class Comm
{
public:
    virtual ~Comm()         = default;
    virtual int rank() const = 0;
    virtual int size() const = 0;

    //! gather `bytes` bytes from every rank into out[size * bytes] (host memory); small metadata only
    virtual int allgatherHost(const void* in, size_t bytes, void* out, cudaStream_t s) = 0;

    //! element-wise sum over ranks of n uint32 values in device memory, in place, ordered on the stream
    virtual int allreduceSumU32(uint32_t* data, size_t n, cudaStream_t s) = 0;

    /*! personalised exchange of device buffers: message k from rank a to rank b is matched with the k-th entry of
     *  b's recv list that names peer a (sizes must agree).  Returns after the receives have been enqueued on the
     *  stream; send buffers may be reused after the next synchronisation of the stream. */
    virtual int exchange(const std::vector<CommMessage>& sends, const std::vector<CommMessage>& recvs,
                         cudaStream_t s) = 0;

    virtual int barrier(cudaStream_t s) = 0;

    /*! make `count` (<= MAX_SHARED) device allocations of this rank addressable by every other rank:
     *  peers[r * count + k] is a pointer through which the caller can store into allocation k of rank r (its own
     *  pointers for r == rank()); extras[r] carries one 64-bit word of rank r along.  `mine` must be allocation base
     *  pointers (cudaMalloc).  Collective; it is ordered after all earlier work on every rank's stream, so a rank may
     *  write into a peer's allocation as soon as the call returns.  Returns 2 (on every rank) when the transport cannot
     *  map peer memory - the caller then uses exchange(). */
    static constexpr int MAX_SHARED = 4;
    virtual int sharePointers(void* const* mine, int count, uint64_t extraMine, std::vector<void*>& peers,
                              std::vector<uint64_t>& extras, cudaStream_t s) = 0;

    //! statistics: bytes sent by this rank through exchange() and allreduce payload since creation
    uint64_t bytesSent{0};
};

//! the trivial communicator of a single-rank domain
class SelfComm final : public Comm
{
public:
    int rank() const override { return 0; }
    int size() const override { return 1; }
    int allgatherHost(const void* in, size_t bytes, void* out, cudaStream_t) override
    {
        std::memcpy(out, in, bytes);
        return 0;
    }
    int allreduceSumU32(uint32_t*, size_t, cudaStream_t) override { return 0; }
    int exchange(const std::vector<CommMessage>& sends, const std::vector<CommMessage>& recvs, cudaStream_t) override
    {
        CSB_REQUIRE(sends.empty() && recvs.empty(), "single-rank communicator cannot exchange messages");
        return 0;
    }
    int barrier(cudaStream_t) override { return 0; }
    int sharePointers(void* const* mine, int count, uint64_t extraMine, std::vector<void*>& peers,
                      std::vector<uint64_t>& extras, cudaStream_t) override
    {
        peers.assign(mine, mine + count);
        extras.assign(1, extraMine);
        return 0;
    }
};

} // namespace csb
