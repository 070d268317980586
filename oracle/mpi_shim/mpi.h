/* TEST INFRASTRUCTURE ONLY — not part of the product.
 *
 * In-process, thread-backed stand-in for <mpi.h> so that the UNMODIFIED reference headers under
 * /root/reference/include (which need MPI for Domain/GlobalAssignment/FocusedOctree/Halos) can be compiled into
 * oracle/_ref/libcstone_ref.so and run with P "ranks" = P std::threads inside one process (no MPI in this image).
 *
 * Implements exactly the 14 entry points the reference calls (census in SURVEY.md §8c):
 * Comm_rank/size, Barrier, Allreduce (SUM, MIN, user ops), Alltoall, Allgatherv, Isend, Irecv, Recv, Probe,
 * Get_count, Waitall, Op_create/free.  Sends are eager (buffered); receives match (source|ANY, tag) FIFO per source.
 */
#pragma once

#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <stdexcept>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef void(MPI_User_function)(void*, void*, int*, MPI_Datatype*);

struct MPI_Op
{
    int kind;                // 0 sum, 1 min, 2 max, 3 user
    MPI_User_function* user; // for kind 3
};

struct MPI_Status
{
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
    long long bytes_;
};

struct MPI_Request
{
    int kind_{0}; // 0 = complete, 1 = pending receive
    void* buf_{nullptr};
    long long bytes_{0};
    int src_{0};
    int tag_{0};
};

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void*)(-1))
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

enum : int
{
    MPI_DATATYPE_NULL = 0,
    MPI_DOUBLE,
    MPI_FLOAT,
    MPI_CHAR,
    MPI_SIGNED_CHAR,
    MPI_UNSIGNED_CHAR,
    MPI_SHORT,
    MPI_UNSIGNED_SHORT,
    MPI_INT,
    MPI_UNSIGNED,
    MPI_LONG,
    MPI_UNSIGNED_LONG,
    MPI_LONG_LONG,
    MPI_UNSIGNED_LONG_LONG,
    MPI_BYTE
};

static const MPI_Op MPI_SUM{0, nullptr};
static const MPI_Op MPI_MIN{1, nullptr};
static const MPI_Op MPI_MAX{2, nullptr};

namespace mpishim
{

inline int typeSize(MPI_Datatype t)
{
    switch (t)
    {
        case MPI_DOUBLE: return 8;
        case MPI_FLOAT: return 4;
        case MPI_CHAR:
        case MPI_SIGNED_CHAR:
        case MPI_UNSIGNED_CHAR:
        case MPI_BYTE: return 1;
        case MPI_SHORT:
        case MPI_UNSIGNED_SHORT: return 2;
        case MPI_INT:
        case MPI_UNSIGNED: return 4;
        case MPI_LONG:
        case MPI_UNSIGNED_LONG:
        case MPI_LONG_LONG:
        case MPI_UNSIGNED_LONG_LONG: return 8;
        default: throw std::runtime_error("mpishim: bad datatype");
    }
}

struct Message
{
    int src, tag;
    std::vector<char> data;
};

struct World
{
    int size{1};
    std::mutex mtx;
    std::condition_variable cv;
    std::vector<std::deque<Message>> mailbox; // per destination
    // collective staging
    std::vector<const void*> slots;
    std::vector<const void*> slots2;
    int barrierCount{0};
    long long barrierGen{0};

    void reset(int P)
    {
        size = P;
        mailbox.assign(P, {});
        slots.assign(P, nullptr);
        slots2.assign(P, nullptr);
        barrierCount = 0;
        barrierGen   = 0;
    }
};

inline World& world()
{
    static World w;
    static std::once_flag once;
    std::call_once(once, [] { w.reset(1); });
    return w;
}

inline int& rankRef()
{
    static thread_local int r = 0;
    return r;
}

inline bool trace()
{
    static bool t = getenv("CS_MPI_TRACE") != nullptr;
    return t;
}
#define MPISHIM_TRACE(...)                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        if (::mpishim::trace())                                                                                        \
        {                                                                                                              \
            fprintf(stderr, "[r%d] ", ::mpishim::rankRef());                                                           \
            fprintf(stderr, __VA_ARGS__);                                                                              \
            fputc('\n', stderr);                                                                                       \
        }                                                                                                              \
    } while (0)

inline void barrier()
{
    World& w = world();
    std::unique_lock<std::mutex> lk(w.mtx);
    long long gen = w.barrierGen;
    if (++w.barrierCount == w.size)
    {
        w.barrierCount = 0;
        w.barrierGen++;
        w.cv.notify_all();
    }
    else { w.cv.wait(lk, [&] { return w.barrierGen != gen; }); }
}

template<class T>
inline void reduceTyped(const T* in, T* inout, int n, int kind)
{
    for (int i = 0; i < n; ++i)
    {
        if (kind == 0) inout[i] = inout[i] + in[i];
        else if (kind == 1) inout[i] = in[i] < inout[i] ? in[i] : inout[i];
        else inout[i] = in[i] > inout[i] ? in[i] : inout[i];
    }
}

inline void reduceInto(const void* in, void* inout, int n, MPI_Datatype t, const MPI_Op& op)
{
    if (op.kind == 3)
    {
        op.user(const_cast<void*>(in), inout, &n, &t);
        return;
    }
    switch (t)
    {
        case MPI_DOUBLE: reduceTyped((const double*)in, (double*)inout, n, op.kind); break;
        case MPI_FLOAT: reduceTyped((const float*)in, (float*)inout, n, op.kind); break;
        case MPI_INT: reduceTyped((const int*)in, (int*)inout, n, op.kind); break;
        case MPI_UNSIGNED: reduceTyped((const unsigned*)in, (unsigned*)inout, n, op.kind); break;
        case MPI_LONG:
        case MPI_LONG_LONG: reduceTyped((const long long*)in, (long long*)inout, n, op.kind); break;
        case MPI_UNSIGNED_LONG:
        case MPI_UNSIGNED_LONG_LONG:
            reduceTyped((const unsigned long long*)in, (unsigned long long*)inout, n, op.kind);
            break;
        case MPI_CHAR:
        case MPI_SIGNED_CHAR: reduceTyped((const signed char*)in, (signed char*)inout, n, op.kind); break;
        case MPI_UNSIGNED_CHAR: reduceTyped((const unsigned char*)in, (unsigned char*)inout, n, op.kind); break;
        case MPI_SHORT: reduceTyped((const short*)in, (short*)inout, n, op.kind); break;
        case MPI_UNSIGNED_SHORT: reduceTyped((const unsigned short*)in, (unsigned short*)inout, n, op.kind); break;
        default: throw std::runtime_error("mpishim: reduce type");
    }
}

//! find first message in my mailbox matching (src, tag); caller holds lock
inline int findMatch(std::deque<Message>& q, int src, int tag)
{
    for (size_t i = 0; i < q.size(); ++i)
    {
        if ((src == MPI_ANY_SOURCE || q[i].src == src) && (tag == MPI_ANY_TAG || q[i].tag == tag)) return int(i);
    }
    return -1;
}

inline void blockingRecv(void* buf, long long maxBytes, int src, int tag, MPI_Status* status)
{
    World& w = world();
    int me   = rankRef();
    std::unique_lock<std::mutex> lk(w.mtx);
    int idx = -1;
    w.cv.wait(lk, [&] { return (idx = findMatch(w.mailbox[me], src, tag)) >= 0; });
    Message& m = w.mailbox[me][idx];
    if ((long long)m.data.size() > maxBytes) throw std::runtime_error("mpishim: message truncated");
    std::memcpy(buf, m.data.data(), m.data.size());
    if (status)
    {
        status->MPI_SOURCE = m.src;
        status->MPI_TAG    = m.tag;
        status->bytes_     = (long long)m.data.size();
    }
    w.mailbox[me].erase(w.mailbox[me].begin() + idx);
}

} // namespace mpishim

inline int MPI_Comm_rank(MPI_Comm, int* r)
{
    *r = mpishim::rankRef();
    return 0;
}
inline int MPI_Comm_size(MPI_Comm, int* s)
{
    *s = mpishim::world().size;
    return 0;
}
inline int MPI_Barrier(MPI_Comm)
{
    MPISHIM_TRACE("Barrier");
    mpishim::barrier();
    return 0;
}

inline int MPI_Op_create(MPI_User_function* f, int, MPI_Op* op)
{
    op->kind = 3;
    op->user = f;
    return 0;
}
inline int MPI_Op_free(MPI_Op*) { return 0; }

inline int
MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype t, const MPI_Op& op, MPI_Comm)
{
    MPISHIM_TRACE("Allreduce count=%d", count);
    using namespace mpishim;
    World& w     = world();
    int me       = rankRef();
    size_t bytes = size_t(count) * typeSize(t);
    std::vector<char> mine(bytes);
    std::memcpy(mine.data(), sendbuf == MPI_IN_PLACE ? recvbuf : sendbuf, bytes);
    w.slots[me] = mine.data();
    barrier();
    std::vector<char> acc(bytes);
    std::memcpy(acc.data(), w.slots[0], bytes);
    for (int r = 1; r < w.size; ++r)
        reduceInto(w.slots[r], acc.data(), count, t, op);
    std::memcpy(recvbuf, acc.data(), bytes);
    barrier();
    return 0;
}

inline int MPI_Alltoall(const void* sendbuf, int sendcount, MPI_Datatype st, void* recvbuf, int, MPI_Datatype, MPI_Comm)
{
    MPISHIM_TRACE("Alltoall");
    using namespace mpishim;
    World& w     = world();
    int me       = rankRef();
    size_t bytes = size_t(sendcount) * typeSize(st);
    std::vector<char> mine(bytes * w.size);
    std::memcpy(mine.data(), sendbuf, bytes * w.size);
    w.slots[me] = mine.data();
    barrier();
    for (int r = 0; r < w.size; ++r)
        std::memcpy((char*)recvbuf + r * bytes, (const char*)w.slots[r] + me * bytes, bytes);
    barrier();
    return 0;
}

inline int MPI_Allgatherv(const void* sendbuf,
                          int sendcount,
                          MPI_Datatype st,
                          void* recvbuf,
                          const int* recvcounts,
                          const int* displs,
                          MPI_Datatype rt,
                          MPI_Comm)
{
    MPISHIM_TRACE("Allgatherv");
    using namespace mpishim;
    World& w  = world();
    int me    = rankRef();
    size_t ts = typeSize(rt);
    std::vector<char> mine;
    if (sendbuf == MPI_IN_PLACE)
    {
        mine.assign((char*)recvbuf + displs[me] * ts, (char*)recvbuf + (displs[me] + recvcounts[me]) * ts);
    }
    else { mine.assign((const char*)sendbuf, (const char*)sendbuf + size_t(sendcount) * typeSize(st)); }
    w.slots[me] = mine.data();
    barrier();
    for (int r = 0; r < w.size; ++r)
        std::memcpy((char*)recvbuf + displs[r] * ts, w.slots[r], recvcounts[r] * ts);
    barrier();
    return 0;
}

inline int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request* req)
{
    MPISHIM_TRACE("Isend dest=%d tag=%d count=%d", dest, tag, count);
    using namespace mpishim;
    World& w = world();
    Message m;
    m.src = rankRef();
    m.tag = tag;
    m.data.assign((const char*)buf, (const char*)buf + size_t(count) * typeSize(t));
    {
        std::lock_guard<std::mutex> lk(w.mtx);
        w.mailbox[dest].push_back(std::move(m));
    }
    w.cv.notify_all();
    if (req) *req = MPI_Request{};
    return 0;
}

inline int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request* req)
{
    MPISHIM_TRACE("Irecv src=%d tag=%d count=%d", src, tag, count);
    // a message that is already here is matched right away (required by the Probe -> Irecv idiom of
    // focus/exchange_focus.hpp: a second Probe must not see the same message again); otherwise the receive stays
    // pending and is matched in Waitall
    using namespace mpishim;
    World& w = world();
    int me   = rankRef();
    req->kind_  = 1;
    req->buf_   = buf;
    req->bytes_ = (long long)count * typeSize(t);
    req->src_   = src;
    req->tag_   = tag;
    std::unique_lock<std::mutex> lk(w.mtx);
    int idx = findMatch(w.mailbox[me], src, tag);
    if (idx >= 0)
    {
        Message& m = w.mailbox[me][idx];
        if ((long long)m.data.size() > req->bytes_) throw std::runtime_error("mpishim: message truncated");
        std::memcpy(buf, m.data.data(), m.data.size());
        w.mailbox[me].erase(w.mailbox[me].begin() + idx);
        req->kind_ = 0;
    }
    return 0;
}

inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status* status)
{
    MPISHIM_TRACE("Recv src=%d tag=%d count=%d", src, tag, count);
    mpishim::blockingRecv(buf, (long long)count * mpishim::typeSize(t), src, tag, status);
    return 0;
}

inline int MPI_Probe(int src, int tag, MPI_Comm, MPI_Status* status)
{
    MPISHIM_TRACE("Probe src=%d tag=%d", src, tag);
    using namespace mpishim;
    World& w = world();
    int me   = rankRef();
    std::unique_lock<std::mutex> lk(w.mtx);
    int idx = -1;
    w.cv.wait(lk, [&] { return (idx = findMatch(w.mailbox[me], src, tag)) >= 0; });
    if (status)
    {
        status->MPI_SOURCE = w.mailbox[me][idx].src;
        status->MPI_TAG    = w.mailbox[me][idx].tag;
        status->bytes_     = (long long)w.mailbox[me][idx].data.size();
    }
    return 0;
}

inline int MPI_Get_count(const MPI_Status* status, MPI_Datatype t, int* count)
{
    *count = int(status->bytes_ / mpishim::typeSize(t));
    return 0;
}

inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status*)
{
    MPISHIM_TRACE("Waitall n=%d", n);
    for (int i = 0; i < n; ++i)
    {
        if (reqs[i].kind_ == 1)
        {
            mpishim::blockingRecv(reqs[i].buf_, reqs[i].bytes_, reqs[i].src_, reqs[i].tag_, nullptr);
            reqs[i].kind_ = 0;
        }
    }
    return 0;
}
