/* Communicators of the multi-rank domain (see comm.cuh): thread-backed LocalComm and NCCL-backed NcclComm, plus their
 * C ABI.  Replaces the MPI calls of the reference's exchange layer: MPI_Allreduce of the global node counts
 * (tree/update_mpi.hpp:86-97), the Isend/Recv pairs of exchangeParticles (domain/domaindecomp_mpi.hpp:69-152) and
 * haloexchange (halos/exchange_halos.hpp:26-90), and the small Alltoall/Allreduce metadata collectives.
 */
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "comm.cuh"
#include "cstone_b200.h"

namespace csb
{

namespace
{

__global__ void addU32Kernel(uint32_t* __restrict__ acc, const uint32_t* __restrict__ in, size_t n)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { acc[i] += in[i]; }
}

/* ------------------------------------------------------------------------------------------------ local (threads) */

struct LocalWorld
{
    explicit LocalWorld(int n)
        : size(n)
        , slot(n, nullptr)
        , sendLists(n, nullptr)
    {
    }

    int size;
    std::mutex mtx;
    std::condition_variable cv;
    int count{0};
    long long gen{0};
    std::vector<const void*> slot;
    std::vector<const std::vector<CommMessage>*> sendLists;

    //! false when a rank has given up (cs_local_world_abort): the waiting ranks return an error instead of hanging
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(mtx);
        if (aborted) { return false; }
        long long g = gen;
        if (++count == size)
        {
            count = 0;
            ++gen;
            cv.notify_all();
        }
        else { cv.wait(lk, [&] { return gen != g || aborted; }); }
        return !aborted;
    }

    void abort()
    {
        std::unique_lock<std::mutex> lk(mtx);
        aborted = true;
        cv.notify_all();
    }
    bool aborted{false};
};

#define CSB_LOCAL_BARRIER()                                                                                            \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!w_->barrier()) { return setLastError("local communicator: another rank failed"), 3; }                    \
    } while (0)

class LocalComm final : public Comm
{
public:
    LocalComm(LocalWorld* w, int rank)
        : w_(w)
        , rank_(rank)
    {
    }
    ~LocalComm() override
    {
        cudaFree(acc_);
        cudaFree(tmp_);
    }

    int rank() const override { return rank_; }
    int size() const override { return w_->size; }

    int allgatherHost(const void* in, size_t bytes, void* out, cudaStream_t) override
    {
        w_->slot[rank_] = in;
        CSB_LOCAL_BARRIER();
        for (int r = 0; r < w_->size; ++r)
            std::memcpy(static_cast<char*>(out) + size_t(r) * bytes, w_->slot[r], bytes);
        CSB_LOCAL_BARRIER();
        return 0;
    }

    int allreduceSumU32(uint32_t* data, size_t n, cudaStream_t s) override
    {
        if (n > cap_)
        {
            cudaFree(acc_);
            cudaFree(tmp_);
            acc_ = tmp_ = nullptr;
            cap_        = n + n / 8;
            CSB_CHECK(cudaMalloc(&acc_, cap_ * sizeof(uint32_t)));
            CSB_CHECK(cudaMalloc(&tmp_, cap_ * sizeof(uint32_t)));
        }
        CSB_CHECK(cudaStreamSynchronize(s)); // my contribution is complete
        w_->slot[rank_] = data;
        CSB_LOCAL_BARRIER();
        for (int r = 0; r < w_->size; ++r)
        {
            uint32_t* dst = r == 0 ? acc_ : tmp_;
            CSB_CHECK(cudaMemcpyAsync(dst, w_->slot[r], n * sizeof(uint32_t), cudaMemcpyDefault, s));
            if (r > 0 && n)
            {
                addU32Kernel<<<iceil(n, 256), 256, 0, s>>>(acc_, tmp_, n);
                CSB_LAUNCH_CHECK();
            }
        }
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_LOCAL_BARRIER(); // every rank has read every contribution
        CSB_CHECK(cudaMemcpyAsync(data, acc_, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        bytesSent += n * sizeof(uint32_t);
        return 0;
    }

    int exchange(const std::vector<CommMessage>& sends, const std::vector<CommMessage>& recvs, cudaStream_t s) override
    {
        CSB_CHECK(cudaStreamSynchronize(s)); // send buffers are complete
        w_->sendLists[rank_] = &sends;
        CSB_LOCAL_BARRIER();
        int status = 0;
        std::vector<int> taken(w_->size, 0);
        for (const CommMessage& r : recvs)
        {
            const auto& peerSends = *w_->sendLists[r.peer];
            int k = taken[r.peer]++, seen = 0;
            const CommMessage* match = nullptr;
            for (const CommMessage& m : peerSends)
            {
                if (m.peer == rank_ && seen++ == k)
                {
                    match = &m;
                    break;
                }
            }
            if (!match || match->bytes != r.bytes)
            {
                setLastError("LocalComm::exchange: unmatched message or size mismatch");
                status = 2;
                continue;
            }
            if (r.bytes && cudaMemcpyAsync(r.ptr, match->ptr, r.bytes, cudaMemcpyDefault, s) != cudaSuccess)
            {
                setLastError("LocalComm::exchange: copy failed");
                status = 1;
            }
        }
        cudaStreamSynchronize(s);
        CSB_LOCAL_BARRIER(); // all copies out of my send buffers are done
        for (const CommMessage& m : sends)
            bytesSent += m.bytes;
        return status;
    }

    int sharePointers(void* const* mine, int count, uint64_t extraMine, std::vector<void*>& peers,
                      std::vector<uint64_t>& extras, cudaStream_t s) override
    {
        // same process: the pointers themselves are valid on every rank's thread
        struct Rec
        {
            void* p[MAX_SHARED];
            uint64_t extra;
        } rec{};
        if (count > MAX_SHARED) { return setLastError("sharePointers: too many allocations"), 1; }
        for (int k = 0; k < count; ++k)
            rec.p[k] = mine[k];
        rec.extra = extraMine;
        CSB_CHECK(cudaStreamSynchronize(s)); // earlier work on my stream (resizes) is complete before peers write
        std::vector<Rec> all(w_->size);
        if (int e = allgatherHost(&rec, sizeof(Rec), all.data(), s)) { return e; }
        peers.resize(size_t(w_->size) * count);
        extras.resize(w_->size);
        // ranks may live on different devices: kernel stores into a peer's arrays need peer access.  Every rank checks
        // what it is going to address; if any pair cannot be mapped, all ranks report 2 (caller falls back to exchange())
        int ok = 1, myDev = 0;
        CSB_CHECK(cudaGetDevice(&myDev));
        for (int r = 0; r < w_->size; ++r)
        {
            for (int k = 0; k < count; ++k)
            {
                void* p                      = all[r].p[k];
                peers[size_t(r) * count + k] = p;
                cudaPointerAttributes attr{};
                if (p == nullptr || cudaPointerGetAttributes(&attr, p) != cudaSuccess)
                {
                    cudaGetLastError();
                    continue; // nothing to address
                }
                if (attr.type != cudaMemoryTypeDevice || attr.device == myDev) { continue; }
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, myDev, attr.device) != cudaSuccess || !can)
                {
                    cudaGetLastError();
                    ok = 0;
                    continue;
                }
                cudaError_t err = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; }
                cudaGetLastError();
            }
            extras[r] = all[r].extra;
        }
        std::vector<int> oks(w_->size);
        if (int e = allgatherHost(&ok, sizeof(int), oks.data(), s)) { return e; }
        for (int v : oks)
            ok = ok && v;
        return ok ? 0 : 2;
    }

    int barrier(cudaStream_t s) override
    {
        CSB_CHECK(cudaStreamSynchronize(s));
        CSB_LOCAL_BARRIER();
        return 0;
    }

private:
    LocalWorld* w_;
    int rank_;
    uint32_t* acc_{nullptr};
    uint32_t* tmp_{nullptr};
    size_t cap_{0};
};

/* ------------------------------------------------------------------------------------------------ NCCL */

struct NcclApi
{
    void* handle{nullptr};
    ncclResult_t (*GetUniqueId)(ncclUniqueId*){nullptr};
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int){nullptr};
    ncclResult_t (*CommDestroy)(ncclComm_t){nullptr};
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t){nullptr};
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t){nullptr};
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t){nullptr};
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t){nullptr};
    ncclResult_t (*GroupStart)(){nullptr};
    ncclResult_t (*GroupEnd)(){nullptr};
    const char* (*GetErrorString)(ncclResult_t){nullptr};
};

NcclApi* ncclApi()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once,
                   []
                   {
                       // an already loaded libnccl.so.2 (e.g. the one PyTorch ships) is reused by soname
                       void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
                       if (!h) { h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); }
                       if (!h) { return; }
                       api.handle = h;
#define CSB_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
                       CSB_NCCL_SYM(GetUniqueId);
                       CSB_NCCL_SYM(CommInitRank);
                       CSB_NCCL_SYM(CommDestroy);
                       CSB_NCCL_SYM(AllReduce);
                       CSB_NCCL_SYM(AllGather);
                       CSB_NCCL_SYM(Send);
                       CSB_NCCL_SYM(Recv);
                       CSB_NCCL_SYM(GroupStart);
                       CSB_NCCL_SYM(GroupEnd);
                       CSB_NCCL_SYM(GetErrorString);
#undef CSB_NCCL_SYM
                   });
    bool ok = api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather &&
              api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return ok ? &api : nullptr;
}

#define CSB_NCCL_CHECK(call)                                                                                           \
    do                                                                                                                 \
    {                                                                                                                  \
        ncclResult_t r__ = (call);                                                                                     \
        if (r__ != ncclSuccess)                                                                                        \
        {                                                                                                              \
            ::csb::setLastError(std::string(#call) + " failed: " +                                                     \
                                (::csb::ncclApi() && ::csb::ncclApi()->GetErrorString ? ::csb::ncclApi()->GetErrorString(r__) : "?") +     \
                                " at " + __FILE__ + ":" + std::to_string(__LINE__));                                   \
            return 1;                                                                                                  \
        }                                                                                                              \
    } while (0)

class NcclComm final : public Comm
{
public:
    NcclComm(NcclApi* api, ncclComm_t comm, int rank, int size)
        : api_(api)
        , comm_(comm)
        , rank_(rank)
        , size_(size)
    {
    }
    ~NcclComm() override
    {
        for (auto& kv : ipcCache_)
            cudaIpcCloseMemHandle(kv.second.ptr);
        if (comm_) { api_->CommDestroy(comm_); }
        cudaFree(stage_);
    }

    int rank() const override { return rank_; }
    int size() const override { return size_; }

    int allgatherHost(const void* in, size_t bytes, void* out, cudaStream_t s) override
    {
        size_t total = bytes * size_t(size_);
        if (total > stageCap_)
        {
            CSB_CHECK(cudaStreamSynchronize(s));
            cudaFree(stage_);
            stage_    = nullptr;
            stageCap_ = std::max<size_t>(2 * total, 4096);
            CSB_CHECK(cudaMalloc(&stage_, stageCap_));
        }
        char* mine = stage_ + size_t(rank_) * bytes;
        CSB_CHECK(cudaMemcpyAsync(mine, in, bytes, cudaMemcpyHostToDevice, s));
        CSB_NCCL_CHECK(api_->AllGather(mine, stage_, bytes, ncclChar, comm_, s));
        CSB_CHECK(cudaMemcpyAsync(out, stage_, total, cudaMemcpyDeviceToHost, s));
        CSB_CHECK(cudaStreamSynchronize(s));
        return 0;
    }

    int allreduceSumU32(uint32_t* data, size_t n, cudaStream_t s) override
    {
        if (n == 0) { return 0; }
        CSB_NCCL_CHECK(api_->AllReduce(data, data, n, ncclUint32, ncclSum, comm_, s));
        bytesSent += n * sizeof(uint32_t);
        return 0;
    }

    int exchange(const std::vector<CommMessage>& sends, const std::vector<CommMessage>& recvs, cudaStream_t s) override
    {
        if (sends.empty() && recvs.empty()) { return 0; }
        CSB_NCCL_CHECK(api_->GroupStart());
        for (const CommMessage& m : sends)
        {
            if (m.bytes) { CSB_NCCL_CHECK(api_->Send(m.ptr, m.bytes, ncclChar, m.peer, comm_, s)); }
            bytesSent += m.bytes;
        }
        for (const CommMessage& m : recvs)
        {
            if (m.bytes) { CSB_NCCL_CHECK(api_->Recv(m.ptr, m.bytes, ncclChar, m.peer, comm_, s)); }
        }
        CSB_NCCL_CHECK(api_->GroupEnd());
        return 0;
    }

    /*! CUDA IPC: every rank publishes the memory handles of its allocations, peers map them once (cached by handle)
     *  and then address the memory with ordinary loads/stores over NVLink */
    int sharePointers(void* const* mine, int count, uint64_t extraMine, std::vector<void*>& peers,
                      std::vector<uint64_t>& extras, cudaStream_t s) override
    {
        struct Rec
        {
            cudaIpcMemHandle_t h[MAX_SHARED];
            uint64_t extra;
            int ok;
        } rec{};
        if (count > MAX_SHARED) { return setLastError("sharePointers: too many allocations"), 1; }
        rec.extra = extraMine;
        rec.ok    = ipcDisabled_ ? 0 : 1;
        for (int k = 0; k < count && rec.ok; ++k)
            if (cudaIpcGetMemHandle(&rec.h[k], mine[k]) != cudaSuccess)
            {
                cudaGetLastError();
                rec.ok = 0;
            }
        std::vector<Rec> all(size_);
        if (int e = allgatherHost(&rec, sizeof(Rec), all.data(), s)) { return e; }
        int ok = 1;
        for (int r = 0; r < size_; ++r)
            ok = ok && all[r].ok;
        peers.assign(size_t(size_) * count, nullptr);
        extras.resize(size_);
        for (int r = 0; r < size_ && ok; ++r)
        {
            extras[r] = all[r].extra;
            for (int k = 0; k < count && ok; ++k)
            {
                if (r == rank_)
                {
                    peers[size_t(r) * count + k] = mine[k];
                    continue;
                }
                // one live mapping per (rank, slot): when rank r publishes a new handle for slot k (it reallocated),
                // the mapping it replaces is closed; nothing that the current call hands out is ever evicted
                std::string handle(reinterpret_cast<const char*>(&all[r].h[k]), sizeof(cudaIpcMemHandle_t));
                auto it = ipcCache_.find({r, k});
                if (it != ipcCache_.end() && it->second.handle != handle)
                {
                    cudaIpcCloseMemHandle(it->second.ptr);
                    ipcCache_.erase(it);
                    it = ipcCache_.end();
                }
                if (it == ipcCache_.end())
                {
                    void* p = nullptr;
                    if (cudaIpcOpenMemHandle(&p, all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
                    {
                        cudaGetLastError();
                        ok = 0;
                        break;
                    }
                    it = ipcCache_.emplace(std::make_pair(r, k), IpcMapping{handle, p}).first;
                }
                peers[size_t(r) * count + k] = it->second.ptr;
            }
        }
        // every rank must take the same path
        std::vector<int> oks(size_);
        if (int e = allgatherHost(&ok, sizeof(int), oks.data(), s)) { return e; }
        for (int v : oks)
            ok = ok && v;
        if (!ok)
        {
            ipcDisabled_ = true;
            return 2;
        }
        return 0;
    }

    int barrier(cudaStream_t s) override
    {
        int token = 0;
        std::vector<int> all(size_);
        return allgatherHost(&token, sizeof(int), all.data(), s);
    }

private:
    NcclApi* api_;
    ncclComm_t comm_;
    int rank_, size_;
    bool ipcDisabled_{std::getenv("CSB_NO_PEER_PUSH") != nullptr};
    struct IpcMapping
    {
        std::string handle;
        void* ptr;
    };
    std::map<std::pair<int, int>, IpcMapping> ipcCache_; // (rank, slot) -> the peer allocation currently mapped
    char* stage_{nullptr};
    size_t stageCap_{0};
};

} // namespace

} // namespace csb

extern "C"
{

void* cs_local_world_create(int size)
{
    if (size < 1)
    {
        csb::setLastError("cs_local_world_create: size must be positive");
        return nullptr;
    }
    return new csb::LocalWorld(size);
}

void cs_local_world_destroy(void* world) { delete static_cast<csb::LocalWorld*>(world); }

void cs_local_world_abort(void* world)
{
    if (world) { static_cast<csb::LocalWorld*>(world)->abort(); }
}

cs_comm_t* cs_comm_create_local(void* world, int rank)
{
    auto* w = static_cast<csb::LocalWorld*>(world);
    if (!w || rank < 0 || rank >= w->size)
    {
        csb::setLastError("cs_comm_create_local: invalid world or rank");
        return nullptr;
    }
    return reinterpret_cast<cs_comm_t*>(static_cast<csb::Comm*>(new csb::LocalComm(w, rank)));
}

int cs_nccl_unique_id(void* out128)
{
    csb::NcclApi* api = csb::ncclApi();
    CSB_REQUIRE(api != nullptr, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    CSB_NCCL_CHECK(api->GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
    return 0;
}

cs_comm_t* cs_comm_create_nccl(int rank, int size, const void* id128)
{
    csb::NcclApi* api = csb::ncclApi();
    if (!api)
    {
        csb::setLastError("libnccl.so.2 could not be loaded");
        return nullptr;
    }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r  = api->CommInitRank(&comm, size, id, rank);
    if (r != ncclSuccess)
    {
        csb::setLastError(std::string("ncclCommInitRank failed: ") + (api->GetErrorString ? api->GetErrorString(r) : ""));
        return nullptr;
    }
    return reinterpret_cast<cs_comm_t*>(static_cast<csb::Comm*>(new csb::NcclComm(api, comm, rank, size)));
}

void cs_comm_destroy(cs_comm_t* c) { delete reinterpret_cast<csb::Comm*>(c); }
int cs_comm_rank(const cs_comm_t* c) { return reinterpret_cast<const csb::Comm*>(c)->rank(); }
int cs_comm_size(const cs_comm_t* c) { return reinterpret_cast<const csb::Comm*>(c)->size(); }
uint64_t cs_comm_bytes_sent(const cs_comm_t* c) { return reinterpret_cast<const csb::Comm*>(c)->bytesSent; }

} // extern "C"
