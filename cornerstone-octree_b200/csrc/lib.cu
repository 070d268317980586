/* library-level state: last error string, launch counter, version */
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>

#include "common.cuh"
#include "cstone_b200.h"

namespace csb
{

static std::mutex g_errMutex;
static std::string g_lastError;
static std::atomic<uint64_t> g_launches{0};

void setLastError(const std::string& msg)
{
    std::lock_guard<std::mutex> lk(g_errMutex);
    g_lastError = msg;
}

static std::atomic<int> g_tuning[TUNE_COUNT] = {};
int tuning(int knob) { return (knob >= 0 && knob < TUNE_COUNT) ? g_tuning[knob].load(std::memory_order_relaxed) : 0; }

void countLaunch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

/* Library-owned scratch for entry points that need internal temporaries.  The reference takes these from the
 * stream-ordered allocator on every call (primitives_gpu.cu:289,302, octree_gpu.cu:190-199); with the default pool's
 * release threshold of zero that returns the memory to the driver at every synchronisation, which on a GPU holding
 * >100 GB of allocations costs up to hundreds of ms per call (measured: profiles/r1_notes.md).  Instead each
 * (device, stream, host thread, slot) owns one buffer that only ever grows.  State is keyed per device, stream AND
 * calling thread: two streams, two devices, or two host threads that share a stream (ranks as threads, the C++
 * forwarder's default stream) never see each other's temporaries, so the multi-kernel sequences that use them
 * (findNeighbors, computeNodeCounts, mergeSortedRuns, gatherArrays) cannot interleave on shared scratch (SURVEY.md 8b,
 * "Threading / stream semantics").  A buffer is only freed or grown by the thread that owns it. */
namespace
{
struct ScratchKey
{
    int device;
    cudaStream_t stream;
    std::thread::id thread;
    int slot;
    bool operator<(const ScratchKey& o) const
    {
        if (device != o.device) { return device < o.device; }
        if (stream != o.stream) { return stream < o.stream; }
        if (thread != o.thread) { return thread < o.thread; }
        return slot < o.slot;
    }
};
struct ScratchBuf
{
    void* p{nullptr};
    size_t cap{0};
};
std::mutex g_scratchMutex;
std::map<ScratchKey, ScratchBuf> g_scratch;
} // namespace

void* scratch(cudaStream_t s, int slot, size_t bytes)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
    {
        setLastError("scratch: no CUDA device");
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_scratchMutex);
    ScratchBuf& b = g_scratch[ScratchKey{dev, s, std::this_thread::get_id(), slot}];
    if (bytes > b.cap)
    {
        if (b.p)
        {
            cudaStreamSynchronize(s);
            cudaFree(b.p);
            b.p   = nullptr;
            b.cap = 0;
        }
        size_t newCap = std::max<size_t>(bytes + bytes / 8, 4096);
        if (cudaMalloc(&b.p, newCap) != cudaSuccess)
        {
            cudaGetLastError();
            b.p = nullptr;
            setLastError("scratch: cudaMalloc of " + std::to_string(newCap) + " bytes failed");
            return nullptr;
        }
        b.cap = newCap;
    }
    return b.p;
}

int releaseScratch()
{
    std::lock_guard<std::mutex> lk(g_scratchMutex);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : g_scratch)
    {
        cudaSetDevice(kv.first.device);
        cudaDeviceSynchronize();
        cudaFree(kv.second.p);
    }
    g_scratch.clear();
    cudaSetDevice(cur);
    return 0;
}

} // namespace csb

extern "C"
{

const char* cs_last_error(void)
{
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(csb::g_errMutex);
    copy = csb::g_lastError;
    return copy.c_str();
}

int cs_version(void) { return 100; }

/* experiment hook: L2 fetch granularity hint (bytes: 32, 64 or 128) for the current device */
int cs_set_l2_fetch_granularity(int bytes)
{
    CSB_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(bytes)));
    return 0;
}

/* experiment hook: select kernel variants (csb::TuningKnob); the defaults are what the library ships with */
int cs_tuning_set(int knob, int value)
{
    CSB_REQUIRE(knob >= 0 && knob < csb::TUNE_COUNT, "unknown tuning knob");
    csb::g_tuning[knob].store(value);
    return 0;
}

int cs_release_scratch(void) { return csb::releaseScratch(); }

uint64_t cs_kernel_launch_count(void) { return csb::g_launches.load(); }

} // extern "C"
