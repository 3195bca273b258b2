// Library-wide state: last error (per thread), launch counter, device properties.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace dodo {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;

void timing_events(cudaEvent_t *a, cudaEvent_t *b) { *a = g_ev_start; *b = g_ev_stop; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;  // B200
    }
    return cached;
}

}  // namespace dodo

extern "C" const char *dodo_last_error(void) { return dodo::g_err; }
extern "C" int dodo_version(void) { return 100; }
extern "C" void dodo_set_timing_events(void *start, void *stop) {
    dodo::g_ev_start = (cudaEvent_t)start;
    dodo::g_ev_stop = (cudaEvent_t)stop;
}
extern "C" uint64_t dodo_launch_count(void) { return dodo::g_launches.load(); }
