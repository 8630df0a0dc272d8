// Host-side staging for the PCIe hop.  No arithmetic happens here: the hot path reads only the X,Y,Z
// int32 triple at bytes 0..11 of every LAS point record (laspy's las.X/.Y/.Z, ui/import_PC.py:47-48,
// utils/tower_extraction.py:60-62), so instead of shipping whole 20..256-byte records over PCIe the host
// gathers those 12 bytes into a dense record stream (a LAS-like stream with record length 12, which every
// device kernel accepts unchanged) written with non-temporal stores into pinned staging memory.
//
// The gather is memory-bound (every cache line of the source is touched), so it runs on a persistent
// pool of host threads that pull fixed-size blocks of records from a shared counter: a thread that loses
// its core to the Python thread driving the GPU only delays its own block, not a 1/T share of the slice.
#include "pch_common.cuh"

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#include <xmmintrin.h>
#define PCH_HOST_SSE 1
#endif

#define PCH_PACK_BLOCK 16384   // records per work item (multiple of 4: every block starts 16-byte aligned in dst)

static void pack_range(const uint8_t* __restrict__ src, int64_t lo, int64_t hi, int32_t rec_len,
                       uint8_t* __restrict__ dst) {
    int64_t i = lo;
#if PCH_HOST_SSE
    // 4 records -> 48 bytes = three aligned 16-byte streaming stores.  A 16-byte load from byte 0 of a
    // record never leaves the record (rec_len >= 16; shorter records take the scalar loop).
    if (rec_len >= 16 && (((uintptr_t)(dst + i * 12)) & 15) == 0) {
        const uint8_t* p = src + i * (int64_t)rec_len;
        float* q = reinterpret_cast<float*>(dst + i * 12);
        const int64_t step = 4 * (int64_t)rec_len;
        for (; i + 4 <= hi; i += 4, p += step, q += 12) {
            _mm_prefetch(reinterpret_cast<const char*>(p + 16 * step), _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(p + 16 * step + 64), _MM_HINT_T0);
            if (step > 128) _mm_prefetch(reinterpret_cast<const char*>(p + 16 * step + 128), _MM_HINT_T0);
            const __m128 a = _mm_loadu_ps(reinterpret_cast<const float*>(p));
            const __m128 b = _mm_loadu_ps(reinterpret_cast<const float*>(p + rec_len));
            const __m128 c = _mm_loadu_ps(reinterpret_cast<const float*>(p + 2 * (int64_t)rec_len));
            const __m128 d = _mm_loadu_ps(reinterpret_cast<const float*>(p + 3 * (int64_t)rec_len));
            const __m128 t0 = _mm_shuffle_ps(a, b, _MM_SHUFFLE(0, 0, 2, 2));   // Z0 Z0 X1 X1
            const __m128 r0 = _mm_shuffle_ps(a, t0, _MM_SHUFFLE(2, 0, 1, 0));  // X0 Y0 Z0 X1
            const __m128 r1 = _mm_shuffle_ps(b, c, _MM_SHUFFLE(1, 0, 2, 1));   // Y1 Z1 X2 Y2
            const __m128 t2 = _mm_shuffle_ps(c, d, _MM_SHUFFLE(0, 0, 2, 2));   // Z2 Z2 X3 X3
            const __m128 r2 = _mm_shuffle_ps(t2, d, _MM_SHUFFLE(2, 1, 2, 0));  // Z2 X3 Y3 Z3
            _mm_stream_ps(q, r0);
            _mm_stream_ps(q + 4, r1);
            _mm_stream_ps(q + 8, r2);
        }
        _mm_sfence();
    }
#endif
    for (; i < hi; ++i) std::memcpy(dst + i * 12, src + i * (int64_t)rec_len, 12);
}

namespace {
struct PackJob {
    const uint8_t* src = nullptr;
    uint8_t* dst = nullptr;
    int64_t n = 0;
    int32_t rec_len = 0;
    std::atomic<int64_t> next{0};
};

// Persistent pool: workers sleep on a condition variable between jobs.  One job at a time (callers are
// serialised by `submit_mu`); the calling thread works on the job too.  The pool is intentionally leaked
// at process exit (detached threads parked on the condition variable).
struct PackPool {
    std::mutex mu, submit_mu;
    std::condition_variable cv, done_cv;
    PackJob* job = nullptr;
    uint64_t generation = 0;
    int wanted = 0;      // workers that should join the current job
    int joined = 0;      // workers that took the current generation
    int running = 0;     // workers still inside the current job
    int n_workers = 0;

    static void drain(PackJob* j) {
        for (;;) {
            const int64_t lo = j->next.fetch_add(PCH_PACK_BLOCK, std::memory_order_relaxed);
            if (lo >= j->n) break;
            const int64_t hi = lo + PCH_PACK_BLOCK < j->n ? lo + PCH_PACK_BLOCK : j->n;
            pack_range(j->src, lo, hi, j->rec_len, j->dst);
        }
    }

    void worker() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return generation != seen && joined < wanted; });
            seen = generation;
            ++joined;
            ++running;
            PackJob* j = job;
            lk.unlock();
            drain(j);
            lk.lock();
            if (--running == 0) done_cv.notify_all();
        }
    }

    void ensure(int workers) {   // called with mu held
        while (n_workers < workers) {
            std::thread(&PackPool::worker, this).detach();
            ++n_workers;
        }
    }

    void run(PackJob* j, int threads) {
        std::lock_guard<std::mutex> submit(submit_mu);
        const int helpers = threads - 1;
        {
            std::lock_guard<std::mutex> lk(mu);
            ensure(helpers);
            job = j;
            wanted = helpers;
            joined = 0;
            ++generation;
        }
        if (helpers > 0) cv.notify_all();
        drain(j);
        std::unique_lock<std::mutex> lk(mu);
        wanted = joined;         // late wakers must not pick up a finished job
        done_cv.wait(lk, [&] { return running == 0; });
        job = nullptr;
    }
};

PackPool& pack_pool() {
    static PackPool* p = new PackPool();   // never destroyed: worker threads outlive static destructors
    return *p;
}
}  // namespace

extern "C" int pch_host_pack_xyz(const void* records_host, int64_t n, int32_t rec_len, void* xyz12_host,
                                 int32_t n_threads) {
    PCH_CHECK_ARG(n >= 0 && rec_len >= 12 && rec_len <= 65535, "bad n/rec_len");
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(records_host && xyz12_host, "null pointer");
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 256) nt = 256;
    const int64_t blocks = (n + PCH_PACK_BLOCK - 1) / PCH_PACK_BLOCK;
    if (nt > blocks) nt = blocks;
    PackJob j;
    j.src = static_cast<const uint8_t*>(records_host);
    j.dst = static_cast<uint8_t*>(xyz12_host);
    j.n = n;
    j.rec_len = rec_len;
    if (nt <= 1) {
        PackPool::drain(&j);
        return PCH_OK;
    }
    pack_pool().run(&j, (int)nt);
    return PCH_OK;
}
