// Host-side staging for the PCIe hop.  No arithmetic happens here: the hot path reads only the X,Y,Z
// int32 triple at bytes 0..11 of every LAS point record (laspy's las.X/.Y/.Z, ui/import_PC.py:47-48,
// utils/tower_extraction.py:60-62), so instead of shipping whole 20..256-byte records over PCIe the host
// gathers those 12 bytes into a dense record stream (a LAS-like stream with record length 12, which every
// device kernel accepts unchanged) written with non-temporal stores into pinned staging memory.
#include "pch_common.cuh"

#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#include <xmmintrin.h>
#define PCH_HOST_SSE 1
#endif

static void pack_range(const uint8_t* __restrict__ src, int64_t lo, int64_t hi, int32_t rec_len,
                       uint8_t* __restrict__ dst) {
    int64_t i = lo;
#if PCH_HOST_SSE
    // 4 records -> 48 bytes = three aligned 16-byte streaming stores (dst is 16-byte aligned and the
    // range starts at a multiple of 4 records).  A 16-byte load from byte 0 of a record never leaves
    // the record (rec_len >= 20 on this path; shorter records take the scalar loop).
    if (rec_len >= 16 && (((uintptr_t)(dst + i * 12)) & 15) == 0) {
        const uint8_t* p = src + i * (int64_t)rec_len;
        float* q = reinterpret_cast<float*>(dst + i * 12);
        for (; i + 4 <= hi; i += 4, p += 4 * (int64_t)rec_len, q += 12) {
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

extern "C" int pch_host_pack_xyz(const void* records_host, int64_t n, int32_t rec_len, void* xyz12_host,
                                 int32_t n_threads) {
    PCH_CHECK_ARG(n >= 0 && rec_len >= 12 && rec_len <= 65535, "bad n/rec_len");
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(records_host && xyz12_host, "null pointer");
    const uint8_t* src = static_cast<const uint8_t*>(records_host);
    uint8_t* dst = static_cast<uint8_t*>(xyz12_host);
    int64_t nt = n_threads > 0 ? n_threads : (int64_t)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    const int64_t grain = 1 << 16;   // records per thread at least; multiples of 4 keep every range 16-byte aligned
    if (nt > (n + grain - 1) / grain) nt = (n + grain - 1) / grain;
    if (nt <= 1) {
        pack_range(src, 0, n, rec_len, dst);
        return PCH_OK;
    }
    const int64_t per = ((n + nt - 1) / nt + 3) / 4 * 4;
    std::vector<std::thread> pool;
    pool.reserve((size_t)nt);
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        pool.emplace_back(pack_range, src, lo, hi, rec_len, dst);
    }
    for (auto& th : pool) th.join();
    return PCH_OK;
}
