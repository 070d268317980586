/* library-level state: last error string, launch counter, version */
#include <atomic>
#include <mutex>

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

void countLaunch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

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

uint64_t cs_kernel_launch_count(void) { return csb::g_launches.load(); }

} // extern "C"
