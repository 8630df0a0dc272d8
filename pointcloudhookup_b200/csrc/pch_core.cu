// Library-wide state: thread-local error string, SM count cache, version.
#include "pch_common.cuh"

// ------------------------------------------------------------------------------------------------
// library-wide helpers
// ------------------------------------------------------------------------------------------------
static thread_local char g_pch_err[512] = "";

void pch_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_pch_err, sizeof(g_pch_err), fmt, ap);
    va_end(ap);
}

int pch_sm_count() {
    static thread_local int cached_dev = -1, cached = PCH_SM_COUNT_FALLBACK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PCH_SM_COUNT_FALLBACK;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

extern "C" const char* pch_last_error(void) { return g_pch_err; }
extern "C" int pch_version(void) { return 100; }

