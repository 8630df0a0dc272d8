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


// ------------------------------------------------------------------------------------------------
// launch counting + optional per-kernel event timing
// ------------------------------------------------------------------------------------------------
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static thread_local cudaEvent_t g_pending_e0 = nullptr;
static thread_local const char* g_pending_name = nullptr;
static std::vector<cudaEvent_t> g_pool;   // recycled events: creating one per launch costs more than the record

static cudaEvent_t prof_get_event() {
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (!g_pool.empty()) {
            cudaEvent_t e = g_pool.back();
            g_pool.pop_back();
            return e;
        }
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    return e;
}

void pch_prof_begin(cudaStream_t st, const char* name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t e0 = prof_get_event();
    if (!e0) return;
    cudaEventRecord(e0, st);
    g_pending_e0 = e0;
    g_pending_name = name;
}

void pch_prof_end(cudaStream_t st) {
    if (!g_pending_e0) return;
    cudaEvent_t e1 = prof_get_event();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (e1) {
        cudaEventRecord(e1, st);
        g_prof.push_back({g_pending_name, g_pending_e0, e1});
    } else {
        g_pool.push_back(g_pending_e0);
    }
    g_pending_e0 = nullptr;
}

extern "C" long long pch_launch_count(void) { return g_launches.load(); }

extern "C" void pch_profile_enable(int on) {
    if (on) {   // pre-create a pool so the timed region only pays for cudaEventRecord
        std::lock_guard<std::mutex> lk(g_prof_mu);
        while (g_pool.size() < 2048) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) break;
            g_pool.push_back(e);
        }
    }
    g_prof_on.store(on ? 1 : 0);
}

// Synchronises, then writes "name launches total_ms\n" lines into buf (and clears the records).
extern "C" int pch_profile_report(char* buf, size_t cap) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        pch_set_error("cudaDeviceSynchronize -> %s", cudaGetErrorString(e));
        return PCH_ERR_CUDA;
    }
    std::map<std::string, std::pair<long long, double>> agg;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (auto& r : g_prof) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
                auto& a = agg[r.name];
                a.first += 1;
                a.second += ms;
            }
            g_pool.push_back(r.e0);
            g_pool.push_back(r.e1);
        }
        g_prof.clear();
    }
    size_t off = 0;
    if (buf && cap) buf[0] = 0;
    for (auto& kv : agg) {
        int n = snprintf(buf + off, off < cap ? cap - off : 0, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first,
                         kv.second.second);
        if (n < 0 || off + (size_t)n >= cap) break;
        off += (size_t)n;
    }
    return PCH_OK;
}
