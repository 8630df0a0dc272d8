// pointcloudhookup_b200 — shared device/host helpers for the sm_100a kernels.
// Everything here is header-inline so every .cu translation unit is self-contained (no -rdc).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <limits.h>
#include <string.h>
#include "../../include/pch_b200.h"

#define PCH_SM_COUNT_FALLBACK 148

void pch_set_error(const char* fmt, ...);
int pch_sm_count();

#define PCH_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            pch_set_error(__VA_ARGS__);                            \
            return PCH_ERR_INVALID;                                \
        }                                                          \
    } while (0)

#define PCH_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            pch_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PCH_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define PCH_LAUNCH_CHECK() PCH_CUDA(cudaGetLastError())

// Every kernel launch goes through PCH_LAUNCH: it counts the launch (pch_launch_count) and, when
// profiling is switched on (pch_profile_enable), brackets it with CUDA events on the launching
// stream so bench.py can attribute device time per kernel without a profiler attached.
void pch_prof_begin(cudaStream_t st, const char* name);
void pch_prof_end(cudaStream_t st);
#define PCH_LAUNCH(st, name, ...)       \
    do {                                \
        pch_prof_begin((st), (name));   \
        __VA_ARGS__;                    \
        pch_prof_end((st));             \
    } while (0)

static inline int64_t pch_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t pch_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) — used to stream raw LAS record bytes and
// geoid grid rows into shared memory with 16-byte-aligned bulk transfers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pch_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void pch_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pch_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pch_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pch_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pch_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pch_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void pch_tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pch_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(pch_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pch_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PCH_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PCH_DONE;\n"
        "bra PCH_WAIT;\n"
        "PCH_DONE:\n"
        "}\n" ::"r"(pch_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// Decoupled look-back (single-pass chained scan).  A status word carries flag + value in ONE
// 64-bit store, so no fence is needed between value and flag.  Tiles take their id from an atomic
// ticket, hence a tile only ever waits on tiles that already started: forward progress holds on
// any grid size.  Spins are bounded; on overflow a device error flag is raised instead of hanging.
// ---------------------------------------------------------------------------------------------
#define PCH_FLAG_EMPTY 0ull
#define PCH_FLAG_AGG 1ull
#define PCH_FLAG_INCL 2ull
#define PCH_SPIN_LIMIT (1u << 27)
#define PCH_LB_WIN 8

__device__ __forceinline__ uint64_t pch_ld_volatile_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pch_st_volatile_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t pch_ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pch_st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 64-bit status: [63:62] flag, [61:0] value.  Called by ONE thread per tile.
// `first` is the first tile of the scan domain (tile == first publishes inclusive directly).
// Split in two so a kernel can publish its aggregate early, do its heavy work, and only then walk back.
__device__ __forceinline__ void pch_lookback_publish_u64(uint64_t* status, int64_t tile, int64_t first, uint64_t aggregate) {
    pch_st_volatile_u64(&status[tile], ((tile == first ? PCH_FLAG_INCL : PCH_FLAG_AGG) << 62) | aggregate);
}
__device__ __forceinline__ uint64_t pch_lookback_walk_u64(uint64_t* status, int64_t tile, int64_t first, uint64_t aggregate,
                                                          int* err_flag) {
    if (tile == first) return 0;
    uint64_t excl = 0;
    // windowed walk: PCH_LB_WIN predecessors are fetched with independent loads, then consumed in order
    int64_t t = tile - 1;
    bool done = false;
    while (!done) {
        uint64_t w[PCH_LB_WIN];
#pragma unroll
        for (int j = 0; j < PCH_LB_WIN; ++j) {
            const int64_t tt = t - j;
            w[j] = tt >= first ? pch_ld_volatile_u64(&status[tt]) : (PCH_FLAG_INCL << 62);
        }
#pragma unroll
        for (int j = 0; j < PCH_LB_WIN; ++j) {
            if (done) break;
            uint32_t spins = 0;
            while ((w[j] >> 62) == PCH_FLAG_EMPTY) {
                w[j] = pch_ld_volatile_u64(&status[t - j]);
                if (++spins > PCH_SPIN_LIMIT) {
                    if (err_flag) atomicExch(err_flag, 1);
                    return excl;
                }
            }
            excl += w[j] & ((1ull << 62) - 1);
            if ((w[j] >> 62) == PCH_FLAG_INCL) done = true;
        }
        t -= PCH_LB_WIN;
    }
    pch_st_volatile_u64(&status[tile], (PCH_FLAG_INCL << 62) | (excl + aggregate));
    return excl;
}
__device__ __forceinline__ uint64_t pch_lookback_u64(uint64_t* status, int64_t tile, int64_t first, uint64_t aggregate,
                                                     int* err_flag) {
    pch_lookback_publish_u64(status, tile, first, aggregate);
    return pch_lookback_walk_u64(status, tile, first, aggregate, err_flag);
}

// ---------------------------------------------------------------------------------------------
// unaligned little-endian XYZ fetch from a raw LAS record (X,Y,Z int32 at byte 0,4,8)
// ALIGN = guaranteed alignment of the record start in bytes (1, 2 or 4).
// ---------------------------------------------------------------------------------------------
template <int ALIGN>
__device__ __forceinline__ void pch_load_xyz(const uint8_t* p, int& X, int& Y, int& Z) {
    if (ALIGN >= 4) {
        const int* q = reinterpret_cast<const int*>(p);
        X = q[0]; Y = q[1]; Z = q[2];
    } else if (ALIGN == 2) {
        // 2-byte aligned record: four aligned 32-bit loads + funnel shifts instead of six 16-bit loads
        // (the 4th word lies inside the same record: LAS records are >= 20 bytes)
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a & 2) * 8u;
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3];
        X = (int)__funnelshift_r(w0, w1, sh);
        Y = (int)__funnelshift_r(w1, w2, sh);
        Z = (int)__funnelshift_r(w2, w3, sh);
    } else {
        X = (int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
        Y = (int)((uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24));
        Z = (int)((uint32_t)p[8] | ((uint32_t)p[9] << 8) | ((uint32_t)p[10] << 16) | ((uint32_t)p[11] << 24));
    }
}

// x = X*scale + offset exactly as numpy/laspy evaluate it: one rounded multiply, one rounded add
// (never contracted into an FMA).
// int32 -> float64, exact, without the conversion unit: the XU pipe that executes I2F.F64 issues a quarter
// warp per cycle and was the top stall of the key kernel (56 % of its samples); 2^52 + 2^31 + X is assembled
// from two integer words and one exact float64 subtract brings X back.
__device__ __forceinline__ double pch_i2d(int X) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)((unsigned)X ^ 0x80000000u)), 4503601774854144.0);
}
__device__ __forceinline__ double pch_scaled(int X, double scale, double offset) {
    return __dadd_rn(__dmul_rn(pch_i2d(X), scale), offset);
}

// Correctly rounded a / b from y = RN(1/b) (a true IEEE divide done once, on the host or per CTA): a product
// and two FMA refinement steps.  The last step is Markstein's: with q faithful, r = a - b*q exact (FMA) and
// y the correctly rounded reciprocal, RN(q + r*y) = RN(a/b) provided b's significand is not all ones and
// nothing under/overflows.  pch_recip_ok(b) vets b; operands outside a comfortable range take __ddiv_rn.
// (Checked against IEEE division on 1.5e9 CPU samples and on the device by pch_selftest_fastdiv.)
__device__ __forceinline__ double pch_div_by_nocheck(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, y, q);
}
// operands for which the FMA sequence cannot under/overflow (zero, inf and nan are NOT in range)
__device__ __forceinline__ bool pch_div_inrange(double a) {
    const double m = fabs(a);
    return m > 1e-200 && m < 1e200;
}
__device__ __forceinline__ bool pch_div_inrange3(double a, double b, double c) {
    const double x = fabs(a), y = fabs(b), z = fabs(c);
    return fmin(fmin(x, y), z) > 1e-200 && fmax(fmax(x, y), z) < 1e200;
}
__device__ __forceinline__ double pch_div_by(double a, double b, double y) {
    if (!pch_div_inrange(a)) return __ddiv_rn(a, b);   // zero, denormal-ish, inf, nan: the slow exact path
    return pch_div_by_nocheck(a, b, y);
}
static inline bool pch_recip_ok(double b) {
    if (!(b == b) || b == 0.0) return false;
    const double m = b < 0 ? -b : b;
    if (!(m > 1e-60 && m < 1e60)) return false;
    uint64_t u;
    memcpy(&u, &b, 8);
    return (u & 0xFFFFFFFFFFFFFull) != 0xFFFFFFFFFFFFFull;
}

__device__ __forceinline__ int pch_warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int pch_warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }

// order-preserving float32 <-> uint32 (for radix select / atomic min on floats)
__device__ __forceinline__ uint32_t pch_f32_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pch_ordered_to_f32(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
