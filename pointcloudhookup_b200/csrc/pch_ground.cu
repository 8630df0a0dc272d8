// Tower-extraction stages A and B (utils/tower_extraction.py:57-93) on the device:
//   centroid  : np.mean(raw_points_f32, axis=0) — numpy reduces axis 0 of a C-contiguous (M,3)
//               array as a plain SEQUENTIAL float32 running sum (verified against numpy 2.3.5;
//               it is not pairwise), then divides in float64 and casts to float32.
//   shift     : points = raw_points - centroid (float32 subtract)
//   select    : the two order statistics np.percentile(z, 25) interpolates between (exact radix
//               select on order-preserving keys; the lerp itself is numpy's, done by the caller)
//   compact   : filtered_points = points[z > thr], order preserving, single pass (look-back)
//   grid mode : north_star's grid min-z / height-above-ground extension (no reference code).
#include "pch_common.cuh"

// ------------------------------------------------------------------------------------------------
// exact emulation of numpy's sequential float32 column sums
// ------------------------------------------------------------------------------------------------
// v1: one thread per column walks the array (bit-exact by construction).  The loads are batched so
// the dependent chain is only the FADD.
__global__ void k_seq_sum_serial(const float* __restrict__ xyz, int64_t m, float* __restrict__ sums) {
    const int c = threadIdx.x;
    if (c >= 3) return;
    float s = 0.0f;
    int64_t i = 0;
    for (; i + 8 <= m; i += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(&xyz[(i + j) * 3 + c]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = __fadd_rn(s, v[j]);
    }
    for (; i < m; ++i) s = __fadd_rn(s, __ldg(&xyz[i * 3 + c]));
    sums[c] = s;
}

__global__ void k_centroid_from_sums(const float* __restrict__ sums, int64_t m, float* __restrict__ centroid) {
    int c = threadIdx.x;
    if (c < 3) centroid[c] = (float)((double)sums[c] / (double)m);  // true_divide in f64, cast to f32
}

// v2: exact parallel evaluation of the same sequential float32 sums.
//
// While the running sum s stays inside one binade, s = m * 2^k with a 24-bit integer mantissa m, and
// fl(s + a) adds an INTEGER to m: q + [r > h] + [r == h] * ((m + q) & 1), where a = q*2^k + r (r < 2^k,
// h = 2^(k-1)) — round-to-nearest-even only looks at the parity of m.  So for a fixed k a run of
// elements is a map  parity(m) -> increment of m, maps compose associatively, and:
//   k_sum_prep   per tile of SQ_TILE elements: float64 sum / max / sign flags          (parallel)
//   k_sum_window exclusive prefix of the tile sums -> predicted binade window per tile  (one CTA)
//   k_sum_tables per tile and per k in its window: the composed map (D[parity 0], D[parity 1])
//   k_sum_chain  one warp per column walks the tiles 32 at a time: a warp scan composes the maps,
//                a ballot finds the first tile where the sum would leave the binade (or whose window
//                misses k); that tile alone is summed with real float32 adds, then the walk resumes.
// The result is bit-identical to the serial loop for ANY input: maps are only used where they are
// provably exact (non-negative finite data, no binade crossing), everything else takes the real adds.
#define SQ_TILE 2048
#define SQ_THREADS 256
#define SQ_EPT (SQ_TILE / SQ_THREADS)
#define SQ_W 4
#define SQ_SAT 0x7fffffffu

struct SqTileInfo {
    double sum;      // float64 sum of the tile (prediction only)
    float maxv;
    int32_t ok;      // all elements finite and >= 0
};

// 8 consecutive points (24 floats, 96 bytes, 16-byte aligned) per thread; zero beyond m
__device__ __forceinline__ void sq_load8(const float* __restrict__ xyz, int64_t p0, int64_t m, float v[24]) {
    if (p0 + SQ_EPT <= m) {
        const float4* q = reinterpret_cast<const float4*>(xyz + p0 * 3);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float4 f = __ldg(q + i);
            v[4 * i + 0] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = (p0 * 3 + i < m * 3) ? xyz[p0 * 3 + i] : 0.0f;
    }
}

__global__ void __launch_bounds__(SQ_THREADS)
k_sum_prep(const float* __restrict__ xyz, int64_t m, int64_t n_tiles, SqTileInfo* __restrict__ info /*[3][n_tiles]*/) {
    __shared__ double s_sum[3][SQ_THREADS / 32];
    __shared__ float s_max[3][SQ_THREADS / 32];
    __shared__ int s_ok[3][SQ_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        float v[24];
        sq_load8(xyz, t * SQ_TILE + (int64_t)tid * SQ_EPT, m, v);
        double sum[3] = {0.0, 0.0, 0.0};
        float mx[3] = {0.f, 0.f, 0.f};
        int ok[3] = {1, 1, 1};
#pragma unroll
        for (int e = 0; e < SQ_EPT; ++e)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float x = v[e * 3 + c];
                sum[c] += (double)x;
                mx[c] = fmaxf(mx[c], x);
                if (!(x >= 0.0f) || !(x < INFINITY)) ok[c] = 0;
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
                ok[c] &= __shfl_xor_sync(0xffffffffu, ok[c], o);
            }
            if (lane == 0) { s_sum[c][warp] = sum[c]; s_max[c][warp] = mx[c]; s_ok[c][warp] = ok[c]; }
        }
        __syncthreads();
        if (tid < 3) {
            double sm = 0.0; float mm = 0.f; int kk = 1;
            for (int w = 0; w < SQ_THREADS / 32; ++w) { sm += s_sum[tid][w]; mm = fmaxf(mm, s_max[tid][w]); kk &= s_ok[tid][w]; }
            SqTileInfo ti; ti.sum = sm; ti.maxv = mm; ti.ok = kk;
            info[tid * n_tiles + t] = ti;
        }
        __syncthreads();
    }
}

// 3 CTAs (one per column) x 1024 threads: exclusive float64 prefix of the tile sums + column max ->
// predicted binade window per tile: klo[c][t] .. klo[c][t] + SQ_W - 1
__global__ void __launch_bounds__(1024) k_sum_window(const SqTileInfo* __restrict__ info, int64_t n_tiles, int32_t* __restrict__ klo) {
    __shared__ double s_w[32];
    __shared__ float s_m[32];
    __shared__ double s_carry;
    __shared__ int s_kcap;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SqTileInfo* ti = info + c * n_tiles;
    float amax = 0.f;
    for (int64_t t = tid; t < n_tiles; t += 1024) amax = fmaxf(amax, ti[t].maxv);
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) s_m[warp] = amax;
    if (tid == 0) s_carry = 0.0;
    __syncthreads();
    if (tid == 0) {
        float a = 0.f;
        for (int w = 0; w < 32; ++w) a = fmaxf(a, s_m[w]);
        s_kcap = (a > 0.f ? ilogbf(a) : -126) + 2;
    }
    __syncthreads();
    const int kcap = s_kcap;
    for (int64_t t0 = 0; t0 < n_tiles; t0 += 1024) {
        const int64_t t = t0 + tid;
        const double v = t < n_tiles ? ti[t].sum : 0.0;
        double x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        double base = s_carry;
        for (int w = 0; w < warp; ++w) base += s_w[w];
        const double excl = base + x - v;
        if (t < n_tiles) {
            int kc = excl > 0.0 ? (ilogb(excl) - 23) : -149;
            if (kc > kcap) kc = kcap;
            klo[c * n_tiles + t] = kc - SQ_W / 2;
        }
        __syncthreads();
        if (tid == 1023) s_carry = base + x;
        __syncthreads();
    }
}

struct SqMap { uint32_t d0, d1; };

__device__ __forceinline__ uint32_t sq_sat_add(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return (a >= SQ_SAT || b >= SQ_SAT || s >= SQ_SAT || s < a) ? SQ_SAT : s;
}
// apply F first, then G
__device__ __forceinline__ SqMap sq_compose(const SqMap& F, const SqMap& G) {
    SqMap H;
    H.d0 = sq_sat_add(F.d0, (F.d0 & 1u) ? G.d1 : G.d0);
    H.d1 = sq_sat_add(F.d1, ((1u + F.d1) & 1u) ? G.d1 : G.d0);
    return H;
}

// increment of the mantissa m (ulp 2^k) when adding the float with bit pattern `bits`:
// q + gt, plus 1 more on a tie (eq) when (m + q) is odd.
// a = ma * 2^(ea-150);  a / 2^k = ma * 2^-shift.  With t = (ma << 32) >> shift (64-bit), q is the high
// word and the low word is the fraction in units of 2^-32: > 0x80000000 rounds up, == is the tie.
// (shift <= 32 loses no bits; for shift > 32 the value is < 2^-8 of an ulp, so q = gt = eq = 0.)
__device__ __forceinline__ void sq_elem(uint32_t bits, int k, uint32_t& q, uint32_t& gt, uint32_t& eq) {
    int ea = (int)((bits >> 23) & 0xffu);
    uint32_t ma = bits & 0x7fffffu;
    if (ea == 0) ea = 1; else ma |= 0x800000u;
    const int shift = k - (ea - 150);
    if (shift >= 0) {
        const unsigned long long t = shift < 64 ? (((unsigned long long)ma << 32) >> shift) : 0ull;
        q = (uint32_t)(t >> 32);
        const uint32_t f = (uint32_t)t;
        gt = f > 0x80000000u;
        eq = f == 0x80000000u;
    } else {
        gt = 0; eq = 0;
        q = ma == 0 ? 0u : ((-shift >= 8) ? SQ_SAT : (ma << (-shift)));
    }
}

// All SQ_W window slots of one element at once: a / 2^(k0+w), w = 0..SQ_W-1.  Logical right shifts compose
// ((x >> s) >> w == x >> (s+w)), so the slots share one variable 64-bit shift and differ by constant shifts;
// the values are exactly those of sq_elem(bits, k0 + w, ...).
__device__ __forceinline__ void sq_elem_window(uint32_t bits, int k0, unsigned long long acc[SQ_W], uint32_t ties[SQ_W]) {
    int ea = (int)((bits >> 23) & 0xffu);
    uint32_t ma = bits & 0x7fffffu;
    if (ea == 0) ea = 1; else ma |= 0x800000u;
    const int shift0 = k0 - (ea - 150);
    if (shift0 >= 0) {
        const unsigned long long T = shift0 < 64 ? (((unsigned long long)ma << 32) >> shift0) : 0ull;
#pragma unroll
        for (int w = 0; w < SQ_W; ++w) {
            const unsigned long long t = T >> w;
            const uint32_t f = (uint32_t)t;
            acc[w] += (t >> 32) + (f > 0x80000000u ? 1u : 0u);
            ties[w] += f == 0x80000000u ? 1u : 0u;
        }
    } else {
#pragma unroll
        for (int w = 0; w < SQ_W; ++w) {
            uint32_t q, gt, eq;
            sq_elem(bits, k0 + w, q, gt, eq);
            acc[w] += (unsigned long long)q + gt;
            ties[w] += eq;
        }
    }
}

// Per tile, column and window slot: the composed map.  Fast path: when no element of the tile is an
// exact tie for this k, the map does not depend on parity and is a plain (order-free) sum, so the
// threads read the tile strided/coalesced and one block reduction finishes it; a (column, slot) that
// does contain exact ties (r == 2^(k-1)) is re-evaluated as an ordered composition.
__global__ void __launch_bounds__(SQ_THREADS, 4)
k_sum_tables(const float* __restrict__ xyz, int64_t m, int64_t n_tiles, const SqTileInfo* __restrict__ info,
             const int32_t* __restrict__ klo, SqMap* __restrict__ table /*[3][n_tiles][SQ_W]*/) {
    __shared__ unsigned long long s_sum[SQ_THREADS / 32][3 * SQ_W];
    __shared__ uint32_t s_tie[SQ_THREADS / 32][3 * SQ_W];
    __shared__ uint32_t s_ties[3 * SQ_W];
    __shared__ SqMap s_cw[SQ_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t base = t * SQ_TILE;
        // the fast path is order-free, so each thread takes 8 consecutive points with six 16-byte loads
        float v[24];
        sq_load8(xyz, base + (int64_t)tid * SQ_EPT, m, v);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int k0 = klo[c * n_tiles + t];
            unsigned long long acc[SQ_W];
            uint32_t ties[SQ_W];
#pragma unroll
            for (int w = 0; w < SQ_W; ++w) { acc[w] = 0; ties[w] = 0; }
#pragma unroll
            for (int e = 0; e < SQ_EPT; ++e) sq_elem_window(__float_as_uint(v[e * 3 + c]), k0, acc, ties);
#pragma unroll
            for (int w = 0; w < SQ_W; ++w) {
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    acc[w] += __shfl_xor_sync(0xffffffffu, acc[w], o);
                    ties[w] += __shfl_xor_sync(0xffffffffu, ties[w], o);
                }
                if (lane == 0) { s_sum[warp][c * SQ_W + w] = acc[w]; s_tie[warp][c * SQ_W + w] = ties[w]; }
            }
        }
        __syncthreads();
        if (tid < 3 * SQ_W) {
            unsigned long long a = 0; uint32_t ti = 0;
            for (int ww = 0; ww < SQ_THREADS / 32; ++ww) { a += s_sum[ww][tid]; ti += s_tie[ww][tid]; }
            s_ties[tid] = ti;
            if (ti == 0) {
                SqMap mp; mp.d0 = mp.d1 = a >= SQ_SAT ? SQ_SAT : (uint32_t)a;
                table[((size_t)(tid / SQ_W) * n_tiles + t) * SQ_W + (tid % SQ_W)] = mp;
            }
        }
        __syncthreads();
        for (int cw = 0; cw < 3 * SQ_W; ++cw) {
            if (s_ties[cw] == 0) continue;  // block-uniform
            const int c = cw / SQ_W, w = cw % SQ_W;
            const int k = klo[c * n_tiles + t] + w;
            SqMap mp; mp.d0 = 0; mp.d1 = 0;
            for (int e = 0; e < SQ_EPT; ++e) {   // ordered: thread owns SQ_EPT consecutive elements
                const int64_t i = base + (int64_t)tid * SQ_EPT + e;
                const uint32_t bits = i < m ? __float_as_uint(__ldg(&xyz[i * 3 + c])) : 0u;
                uint32_t q, gt, eq;
                sq_elem(bits, k, q, gt, eq);
                const uint32_t base_inc = sq_sat_add(q, gt);
                const uint32_t i0 = sq_sat_add(base_inc, eq & ((mp.d0 + q) & 1u));
                const uint32_t i1 = sq_sat_add(base_inc, eq & ((1u + mp.d1 + q) & 1u));
                mp.d0 = sq_sat_add(mp.d0, i0);
                mp.d1 = sq_sat_add(mp.d1, i1);
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                SqMap nb;
                nb.d0 = __shfl_down_sync(0xffffffffu, mp.d0, o);
                nb.d1 = __shfl_down_sync(0xffffffffu, mp.d1, o);
                if (lane + o < 32 && (lane % (2 * o)) == 0) mp = sq_compose(mp, nb);
            }
            if (lane == 0) s_cw[warp] = mp;
            __syncthreads();
            if (tid == 0) {
                SqMap r = s_cw[0];
                for (int ww = 1; ww < SQ_THREADS / 32; ++ww) r = sq_compose(r, s_cw[ww]);
                table[((size_t)c * n_tiles + t) * SQ_W + w] = r;
            }
            __syncthreads();
        }
    }
}

// Level 2: 32 consecutive tiles (65 536 elements) whose windows start at the same binade are composed
// into one "super" map per window slot, in parallel (one warp per super-tile).  The chain then walks
// super-tiles and only descends to tile level where a super-tile cannot be applied.
#define SQ_SUPER 32
__global__ void __launch_bounds__(256)
k_sum_super(int64_t n_tiles, int64_t n_super, const SqTileInfo* __restrict__ info, const int32_t* __restrict__ klo,
            const SqMap* __restrict__ table, SqMap* __restrict__ stable /*[3][n_super][SQ_W]*/,
            int32_t* __restrict__ sklo /*[3][n_super]; INT_MIN = unusable*/) {
    const int lane = threadIdx.x & 31;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < 3 * n_super; w += nw) {
        const int c = (int)(w / n_super);
        const int64_t su = w - (int64_t)c * n_super;
        const int64_t t = su * SQ_SUPER + lane;
        const bool have = t < n_tiles;
        const int k_me = have ? klo[c * n_tiles + t] : 0;
        const int ok_me = have ? info[c * n_tiles + t].ok : 0;
        const int k0 = __shfl_sync(0xffffffffu, k_me, 0);
        // usable only when all 32 tiles exist, are non-negative and share one window
        const bool uniform = __all_sync(0xffffffffu, have && ok_me && k_me == k0);
        if (!uniform) {
            if (lane == 0) sklo[c * n_super + su] = INT_MIN;
            continue;
        }
#pragma unroll
        for (int ww = 0; ww < SQ_W; ++ww) {
            SqMap F = table[((size_t)c * n_tiles + t) * SQ_W + ww];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                SqMap pv;
                pv.d0 = __shfl_up_sync(0xffffffffu, F.d0, o);
                pv.d1 = __shfl_up_sync(0xffffffffu, F.d1, o);
                if (lane >= o) F = sq_compose(pv, F);
            }
            if (lane == 31) stable[((size_t)c * n_super + su) * SQ_W + ww] = F;
        }
        if (lane == 0) sklo[c * n_super + su] = k0;
    }
}

// Level 3: 32 consecutive super-tiles (= one SQ_BATCH of 1024 tiles, 2 M elements) sharing one window are
// composed once more (one warp per batch), so the chain crosses a whole batch with a single map wherever the
// running sum stays inside one binade — which is most of a large cloud, and all of it once the sum stagnates.
__global__ void __launch_bounds__(256)
k_sum_super2(int64_t n_super, int64_t n_batch, const SqMap* __restrict__ stable, const int32_t* __restrict__ sklo,
             SqMap* __restrict__ btable /*[3][n_batch][SQ_W]*/, int32_t* __restrict__ bklo /*[3][n_batch]*/) {
    const int lane = threadIdx.x & 31;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < 3 * n_batch; w += nw) {
        const int c = (int)(w / n_batch);
        const int64_t b = w - (int64_t)c * n_batch;
        const int64_t su = b * SQ_SUPER + lane;
        const bool have = su < n_super;
        const int k_me = have ? sklo[c * n_super + su] : INT_MIN;
        const int k0 = __shfl_sync(0xffffffffu, k_me, 0);
        const bool uniform = __all_sync(0xffffffffu, have && k_me != INT_MIN && k_me == k0);
        if (!uniform) {
            if (lane == 0) bklo[c * n_batch + b] = INT_MIN;
            continue;
        }
#pragma unroll
        for (int ww = 0; ww < SQ_W; ++ww) {
            SqMap F = stable[((size_t)c * n_super + su) * SQ_W + ww];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                SqMap pv;
                pv.d0 = __shfl_up_sync(0xffffffffu, F.d0, o);
                pv.d1 = __shfl_up_sync(0xffffffffu, F.d1, o);
                if (lane >= o) F = sq_compose(pv, F);
            }
            if (lane == 31) btable[((size_t)c * n_batch + b) * SQ_W + ww] = F;
        }
        if (lane == 0) bklo[c * n_batch + b] = k0;
    }
}

// One CTA per column.  All 256 threads stage the maps of SQ_BATCH tiles into shared memory (coalesced,
// many loads in flight); warp 0 then walks the batch 32 tiles at a time with warp scans — no global
// latency on the serial path.
#define SQ_BATCH 1024
#define SQ_CHAIN_THREADS 256

__global__ void __launch_bounds__(SQ_CHAIN_THREADS)
k_sum_chain(const float* __restrict__ xyz, int64_t m, int64_t n_tiles, const SqTileInfo* __restrict__ info,
            const int32_t* __restrict__ klo, const SqMap* __restrict__ table, int64_t n_super,
            const SqMap* __restrict__ stable, const int32_t* __restrict__ sklo, const SqMap* __restrict__ btable,
            const int32_t* __restrict__ bklo, float* __restrict__ sums,
            int* __restrict__ stats /*[3][2]: tiles via maps, tiles via real adds*/) {
    __shared__ SqMap s_tab[SQ_BATCH * SQ_W];
    __shared__ int32_t s_klo[SQ_BATCH];
    __shared__ uint8_t s_ok[SQ_BATCH];
    __shared__ SqMap s_stab[(SQ_BATCH / SQ_SUPER) * SQ_W];
    __shared__ int32_t s_sklo[SQ_BATCH / SQ_SUPER];
    __shared__ __align__(16) float s_col[SQ_TILE];
    __shared__ float s_state[2];   // double-buffered by batch parity: the next batch's decision never races this one's readers
    __shared__ int s_skip[2];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_batch = (n_tiles + SQ_BATCH - 1) / SQ_BATCH;
    float s = 0.0f;
    int n_map = 0, n_real = 0;
    for (int64_t b0 = 0; b0 < n_tiles; b0 += SQ_BATCH) {
        const int nb = (int)min((int64_t)SQ_BATCH, n_tiles - b0);
        const int slot = (int)((b0 / SQ_BATCH) & 1);
        // ---- level 3: the whole batch with one map (thread 0 decides; nothing is staged when it applies)
        if (tid == 0) {
            int skip = 0;
            const int32_t bk = bklo[c * n_batch + b0 / SQ_BATCH];
            const uint32_t sb = __float_as_uint(s);
            const int es = (int)((sb >> 23) & 0xffu);
            if (bk != INT_MIN && es != 0 && es != 0xff && !(sb >> 31)) {
                const int idx = (es - 150) - bk;
                if (idx >= 0 && idx < SQ_W) {
                    const SqMap F = btable[((size_t)c * n_batch + b0 / SQ_BATCH) * SQ_W + idx];
                    const uint32_t ms = (sb & 0x7fffffu) | 0x800000u;
                    const uint32_t D = (ms & 1u) ? F.d1 : F.d0;
                    if (D < SQ_SAT && (uint64_t)ms + D < (1ull << 24)) {
                        s = __uint_as_float(((uint32_t)es << 23) | ((ms + D) & 0x7fffffu));
                        n_map += nb;
                        skip = 1;
                    }
                }
            }
            s_state[slot] = s;
            s_skip[slot] = skip;
        }
        __syncthreads();
        s = s_state[slot];
        if (s_skip[slot]) continue;    // block-uniform
        const int nsb = (int)min((int64_t)(SQ_BATCH / SQ_SUPER), n_super - b0 / SQ_SUPER);   // super-tiles in this batch
        for (int i = tid; i < nsb * SQ_W; i += SQ_CHAIN_THREADS)
            s_stab[i] = stable[((size_t)c * n_super + b0 / SQ_SUPER) * SQ_W + i];
        for (int i = tid; i < nsb; i += SQ_CHAIN_THREADS) s_sklo[i] = sklo[c * n_super + b0 / SQ_SUPER + i];
        {
            const uint4* src = reinterpret_cast<const uint4*>(table + ((size_t)c * n_tiles + b0) * SQ_W);
            uint4* dst = reinterpret_cast<uint4*>(s_tab);
            for (int i = tid; i < nb * SQ_W / 2; i += SQ_CHAIN_THREADS) dst[i] = __ldg(src + i);
            for (int i = tid; i < nb; i += SQ_CHAIN_THREADS) {
                s_klo[i] = klo[c * n_tiles + b0 + i];
                s_ok[i] = (uint8_t)(info[c * n_tiles + b0 + i].ok != 0);
            }
        }
        __syncthreads();
        {
            // Every warp runs the same walk redundantly (identical state in all threads), so the whole CTA can
            // take part in the rare tile that needs real adds.
            int t = 0;   // tile index inside the batch
            while (t < nb) {
                const uint32_t sb = __float_as_uint(s);
                const int es = (int)((sb >> 23) & 0xffu);
                int advanced = 0;
                // ---- level 2: whole super-tiles (only from a super-tile boundary)
                if ((t % SQ_SUPER) == 0 && es != 0 && es != 0xff && !(sb >> 31)) {
                    const int k = es - 150;
                    const uint32_t ms = (sb & 0x7fffffu) | 0x800000u;
                    const int su = t / SQ_SUPER + lane;
                    SqMap F; F.d0 = SQ_SAT; F.d1 = SQ_SAT;
                    if (su < nsb && s_sklo[su] != INT_MIN) {
                        const int idx = k - s_sklo[su];
                        if (idx >= 0 && idx < SQ_W) F = s_stab[su * SQ_W + idx];
                    }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        SqMap pv;
                        pv.d0 = __shfl_up_sync(0xffffffffu, F.d0, o);
                        pv.d1 = __shfl_up_sync(0xffffffffu, F.d1, o);
                        if (lane >= o) F = sq_compose(pv, F);
                    }
                    const uint32_t D = (ms & 1u) ? F.d1 : F.d0;
                    const bool ok = D < SQ_SAT && (uint64_t)ms + D < (1ull << 24);
                    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                    const int L = okmask == 0xffffffffu ? 32 : (__ffs(~okmask) - 1);
                    if (L > 0) {
                        const uint32_t Dl = __shfl_sync(0xffffffffu, D, L - 1);
                        s = __uint_as_float(((uint32_t)es << 23) | ((ms + Dl) & 0x7fffffu));
                        t += L * SQ_SUPER;
                        n_map += L * SQ_SUPER;
                        continue;
                    }
                }
                // ---- level 1: the tiles up to the next super-tile boundary (so level 2 can resume there)
                const int cap = min(nb, (t / SQ_SUPER + 1) * SQ_SUPER);
                if (es != 0 && es != 0xff && !(sb >> 31)) {
                    const int k = es - 150;
                    const uint32_t ms = (sb & 0x7fffffu) | 0x800000u;
                    SqMap F; F.d0 = SQ_SAT; F.d1 = SQ_SAT;
                    const int tt = t + lane;
                    if (tt < cap && s_ok[tt]) {
                        const int idx = k - s_klo[tt];
                        if (idx >= 0 && idx < SQ_W) F = s_tab[tt * SQ_W + idx];
                    }
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        SqMap pv;
                        pv.d0 = __shfl_up_sync(0xffffffffu, F.d0, o);
                        pv.d1 = __shfl_up_sync(0xffffffffu, F.d1, o);
                        if (lane >= o) F = sq_compose(pv, F);
                    }
                    const uint32_t D = (ms & 1u) ? F.d1 : F.d0;
                    const bool ok = D < SQ_SAT && (uint64_t)ms + D < (1ull << 24);
                    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                    const int L = okmask == 0xffffffffu ? 32 : (__ffs(~okmask) - 1);
                    if (L > 0) {
                        const uint32_t Dl = __shfl_sync(0xffffffffu, D, L - 1);
                        s = __uint_as_float(((uint32_t)es << 23) | ((ms + Dl) & 0x7fffffu));
                        advanced = L;
                        n_map += L;
                    }
                }
                t += advanced;
                if (t >= cap) continue;   // reached the boundary without a failing tile
                {
                    // Tile b0+t leaves the binade / is outside its window / has negative data: the literal float32
                    // running sum.  The CTA stages the column in shared memory (coalesced, all loads in flight), then
                    // every thread performs the same 2048 dependent adds from broadcast shared-memory reads (~4 cycles
                    // each) — cheaper than any map construction for the handful of tiles per column that get here.
                    const int64_t tg = b0 + t;
                    const int64_t lo = tg * SQ_TILE, hi = min(lo + (int64_t)SQ_TILE, m);
                    const int n_el = (int)(hi - lo);
                    __syncthreads();     // previous use of s_col is over
#pragma unroll
                    for (int j = 0; j < SQ_TILE / SQ_CHAIN_THREADS; ++j) {
                        const int e = j * SQ_CHAIN_THREADS + tid;
                        s_col[e] = e < n_el ? __ldg(&xyz[(lo + e) * 3 + c]) : 0.0f;
                    }
                    __syncthreads();
                    if (n_el == SQ_TILE) {
                        const float4* q = reinterpret_cast<const float4*>(s_col);
#pragma unroll 4
                        for (int j = 0; j < SQ_TILE / 4; ++j) {
                            const float4 v = q[j];
                            s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
                        }
                    } else {
                        for (int j = 0; j < n_el; ++j) s = __fadd_rn(s, s_col[j]);
                    }
                    ++t;
                    ++n_real;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        sums[c] = s;
        if (stats) { stats[c * 2] = n_map; stats[c * 2 + 1] = n_real; }
    }
}

extern "C" size_t pch_f32_centroid_workspace_bytes(int64_t m) {
    int64_t nt = pch_ceil_div(m > 0 ? m : 1, SQ_TILE);
    int64_t ns = pch_ceil_div(nt, 32);
    int64_t nbt = pch_ceil_div(nt, SQ_BATCH);
    return 256 + pch_align_up((size_t)3 * nt * sizeof(SqTileInfo), 256) + pch_align_up((size_t)3 * nt * 4, 256) +
           pch_align_up((size_t)3 * nt * SQ_W * sizeof(SqMap), 256) + pch_align_up((size_t)3 * ns * SQ_W * sizeof(SqMap), 256) +
           pch_align_up((size_t)3 * ns * 4, 256) + pch_align_up((size_t)3 * nbt * SQ_W * sizeof(SqMap), 256) +
           pch_align_up((size_t)3 * nbt * 4, 256);
}

extern "C" int pch_f32_centroid(const float* xyz, int64_t m, float* sums3, float* centroid3, void* workspace,
                                size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 1, "centroid of an empty cloud");
    PCH_CHECK_ARG(xyz && sums3 && centroid3, "null pointer");
    if (!workspace) {
        PCH_LAUNCH(st, "k_seq_sum_serial", k_seq_sum_serial<<<1, 32, 0, st>>>(xyz, m, sums3));
        PCH_LAUNCH_CHECK();
    } else {
        size_t need = pch_f32_centroid_workspace_bytes(m);
        if (workspace_bytes < need) {
            pch_set_error("centroid workspace too small: %zu < %zu", workspace_bytes, need);
            return PCH_ERR_WORKSPACE;
        }
        int64_t nt = pch_ceil_div(m, SQ_TILE);
        uint8_t* base = (uint8_t*)workspace;
        int* stats = (int*)base;
        size_t off = 256;
        SqTileInfo* info = (SqTileInfo*)(base + off); off += pch_align_up((size_t)3 * nt * sizeof(SqTileInfo), 256);
        int32_t* klo = (int32_t*)(base + off); off += pch_align_up((size_t)3 * nt * 4, 256);
        SqMap* table = (SqMap*)(base + off); off += pch_align_up((size_t)3 * nt * SQ_W * sizeof(SqMap), 256);
        const int64_t ns = pch_ceil_div(nt, SQ_SUPER);
        SqMap* stable = (SqMap*)(base + off); off += pch_align_up((size_t)3 * ns * SQ_W * sizeof(SqMap), 256);
        int32_t* sklo = (int32_t*)(base + off); off += pch_align_up((size_t)3 * ns * 4, 256);
        const int64_t nbt = pch_ceil_div(nt, SQ_BATCH);
        SqMap* btable = (SqMap*)(base + off); off += pch_align_up((size_t)3 * nbt * SQ_W * sizeof(SqMap), 256);
        int32_t* bklo = (int32_t*)(base + off);
        unsigned grid = (unsigned)(nt < (int64_t)pch_sm_count() * 8 ? nt : (int64_t)pch_sm_count() * 8);
        PCH_LAUNCH(st, "k_sum_prep", k_sum_prep<<<grid, SQ_THREADS, 0, st>>>(xyz, m, nt, info));
        PCH_LAUNCH(st, "k_sum_window", k_sum_window<<<3, 1024, 0, st>>>(info, nt, klo));
        PCH_LAUNCH(st, "k_sum_tables", k_sum_tables<<<grid, SQ_THREADS, 0, st>>>(xyz, m, nt, info, klo, table));
        PCH_LAUNCH(st, "k_sum_super", k_sum_super<<<(unsigned)(pch_ceil_div(3 * ns, 8) < 1184 ? pch_ceil_div(3 * ns, 8) : 1184), 256, 0, st>>>(
                                          nt, ns, info, klo, table, stable, sklo));
        PCH_LAUNCH(st, "k_sum_super2", k_sum_super2<<<(unsigned)(pch_ceil_div(3 * nbt, 8) < 1184 ? pch_ceil_div(3 * nbt, 8) : 1184), 256, 0, st>>>(
                                           ns, nbt, stable, sklo, btable, bklo));
        PCH_LAUNCH(st, "k_sum_chain", k_sum_chain<<<3, SQ_CHAIN_THREADS, 0, st>>>(xyz, m, nt, info, klo, table, ns, stable, sklo,
                                                                                 btable, bklo, sums3, stats));
        PCH_LAUNCH_CHECK();
    }
    PCH_LAUNCH(st, "k_centroid_from_sums", k_centroid_from_sums<<<1, 32, 0, st>>>(sums3, m, centroid3));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// shifted z column (+ optional full shifted copy)
// ------------------------------------------------------------------------------------------------
__global__ void k_shift(const float* __restrict__ xyz, int64_t m, const float* __restrict__ centroid,
                        float* __restrict__ zs, float* __restrict__ shifted) {
    const float cx = centroid[0], cy = centroid[1], cz = centroid[2];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
        float z2 = __fsub_rn(z, cz);
        if (zs) zs[i] = z2;
        if (shifted) {
            shifted[i * 3 + 0] = __fsub_rn(x, cx);
            shifted[i * 3 + 1] = __fsub_rn(y, cy);
            shifted[i * 3 + 2] = z2;
        }
    }
}

__global__ void k_column(const float* __restrict__ xyz, int64_t m, int col, float* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) out[i] = xyz[i * 3 + col];
}

static unsigned grid_for(int64_t n, int threads, int per_sm) {
    int64_t b = pch_ceil_div(n, threads);
    int64_t cap = (int64_t)pch_sm_count() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" int pch_f32_shift(const float* xyz, int64_t m, const float* centroid3, float* zs, float* shifted,
                             pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && centroid3 && (zs || shifted), "null pointer");
    PCH_LAUNCH(st, "k_shift", k_shift<<<grid_for(m, 256, 8), 256, 0, st>>>(xyz, m, centroid3, zs, shifted));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_f32_column(const float* xyz, int64_t m, int32_t column, float* out, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0 && column >= 0 && column < 3, "bad m/column");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && out, "null pointer");
    PCH_LAUNCH(st, "k_column", k_column<<<grid_for(m, 256, 8), 256, 0, st>>>(xyz, m, column, out));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// exact order statistics (radix select, 4 x 8 bits, two ranks at once)
// ------------------------------------------------------------------------------------------------
struct SelState {
    uint32_t prefix[2];
    unsigned long long rem[2];  // rank still to skip inside the prefix bucket
    uint32_t hist[4][2][256];
};

__global__ void k_sel_init(SelState* s, long long r0, long long r1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* h = &s->hist[0][0][0];
    for (int k = i; k < 4 * 2 * 256; k += gridDim.x * blockDim.x) h[k] = 0;
    if (i == 0) {
        s->prefix[0] = s->prefix[1] = 0;
        s->rem[0] = (unsigned long long)r0;
        s->rem[1] = (unsigned long long)r1;
    }
}

__global__ void __launch_bounds__(256) k_sel_hist(const float* __restrict__ v, int64_t n, SelState* __restrict__ s, int pass) {
    __shared__ uint32_t sh[2][256];
    const int tid = threadIdx.x;
    sh[0][tid] = 0;
    sh[1][tid] = 0;
    __syncthreads();
    const uint32_t pa = s->prefix[0], pb = s->prefix[1];
    const int shift = 24 - 8 * pass;
    const uint32_t hi_mask = pass == 0 ? 0u : (0xffffffffu << (32 - 8 * pass));
    const bool two = pa != pb;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint32_t u = pch_f32_to_ordered(v[i]);
        uint32_t top = u & hi_mask;
        uint32_t bin = (u >> shift) & 255u;
        if (top == pa) atomicAdd(&sh[0][bin], 1u);
        if (two && top == pb) atomicAdd(&sh[1][bin], 1u);
    }
    __syncthreads();
    if (sh[0][tid]) atomicAdd(&s->hist[pass][0][tid], sh[0][tid]);
    if (two && sh[1][tid]) atomicAdd(&s->hist[pass][1][tid], sh[1][tid]);
}

__global__ void k_sel_decide(SelState* s, int pass, float* out2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const bool two = s->prefix[0] != s->prefix[1];
    const int shift = 24 - 8 * pass;
    for (int r = 0; r < 2; ++r) {
        const uint32_t* h = s->hist[pass][(two && r == 1) ? 1 : 0];
        unsigned long long rem = s->rem[r], cum = 0;
        int b = 0;
        for (; b < 255; ++b) {
            if (cum + h[b] > rem) break;
            cum += h[b];
        }
        s->rem[r] = rem - cum;
        s->prefix[r] |= ((uint32_t)b << shift);
    }
    if (pass == 3) {
        out2[0] = pch_ordered_to_f32(s->prefix[0]);
        out2[1] = pch_ordered_to_f32(s->prefix[1]);
    }
}

extern "C" size_t pch_select_workspace_bytes(void) { return pch_align_up(sizeof(SelState), 256); }

extern "C" int pch_select_f32(const float* v, int64_t n, int64_t rank0, int64_t rank1, float* out2, void* workspace,
                              size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 1, "select on an empty array");
    PCH_CHECK_ARG(rank0 >= 0 && rank0 < n && rank1 >= 0 && rank1 < n, "rank out of range");
    PCH_CHECK_ARG(v && out2 && workspace, "null pointer");
    if (workspace_bytes < sizeof(SelState)) {
        pch_set_error("select workspace too small");
        return PCH_ERR_WORKSPACE;
    }
    SelState* s = (SelState*)workspace;
    PCH_LAUNCH(st, "k_sel_init", k_sel_init<<<8, 256, 0, st>>>(s, rank0, rank1));
    PCH_LAUNCH_CHECK();
    unsigned grid = grid_for(n, 256 * 8, 8);
    for (int pass = 0; pass < 4; ++pass) {
        PCH_LAUNCH(st, "k_sel_hist", k_sel_hist<<<grid, 256, 0, st>>>(v, n, s, pass));
        PCH_LAUNCH_CHECK();
        PCH_LAUNCH(st, "k_sel_decide", k_sel_decide<<<1, 32, 0, st>>>(s, pass, out2));
        PCH_LAUNCH_CHECK();
    }
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// grid min-z ground model, shared pieces (north_star subsystem 2; no reference implementation exists):
//   p = raw - centroid (float32), cell (i, j) = floor((p.xy - min_xy) / cell) in float32 arithmetic,
//   ground_z(cell) = min p.z of the cell, keep = (p.z - ground_z) > hag.   oracle/ground.py::grid_min_keep_mask
// ------------------------------------------------------------------------------------------------
struct GridTest {
    const uint32_t* cell_min;   // [nx*ny] order-preserving encoding of the float32 minima; NULL = no grid test
    float minx, miny, cell, hag;
    float rcell;                // RN(1/cell) for the reciprocal division
    int32_t fast;               // cell passed grid_recip_ok: (d / cell) may be evaluated as products + FMA corrections
    int32_t nx, ny;
    long long n_cells;
};
// Correctly rounded float32 d / c from y = RN(1/c): the float32 twin of pch_div_by (Markstein: with a faithful
// quotient, an exact FMA residual and the correctly rounded reciprocal, the last correction returns RN(d/c) unless
// c's significand is all ones or something under/overflows).  Operands outside a comfortable range take the true
// divide.  Checked against __fdiv_rn on the device by pch_selftest_fastdiv_f32.
__device__ __forceinline__ float grid_div(float d, const GridTest& gt) {
    const float m = fabsf(d);
    if (!(gt.fast && m > 1e-30f && m < 1e30f)) return __fdiv_rn(d, gt.cell);
    float q = __fmul_rn(d, gt.rcell);
    float r = __fmaf_rn(-gt.cell, q, d);
    q = __fmaf_rn(r, gt.rcell, q);
    r = __fmaf_rn(-gt.cell, q, d);
    return __fmaf_rn(r, gt.rcell, q);
}
// floor(d / cell) for the cell index, without the range guard of grid_div: the guard only protects quotients that
// cannot be a valid cell anyway.  d == 0 gives exactly 0; 0 < d < 1e-30 gives a quotient below 1 whatever its last
// bit, i.e. floor 0 like the true divide; |d| >= 1e30, inf and nan give a quotient that fails the range test of
// grid_cell_of exactly like the true divide's would; a negative d cannot occur for a point inside the bounding box
// the grid origin was taken from (and floors to a negative, rejected index if it does).
__device__ __forceinline__ float grid_floor_div(float d, const GridTest& gt) {
    if (!gt.fast) return floorf(__fdiv_rn(d, gt.cell));
    float q = __fmul_rn(d, gt.rcell);
    float r = __fmaf_rn(-gt.cell, q, d);
    q = __fmaf_rn(r, gt.rcell, q);
    r = __fmaf_rn(-gt.cell, q, d);
    return floorf(__fmaf_rn(r, gt.rcell, q));
}
static inline bool grid_recip_ok(float c) {
    if (!(c == c) || c == 0.f) return false;
    const float m = c < 0 ? -c : c;
    if (!(m > 1e-15f && m < 1e15f)) return false;
    uint32_t u;
    memcpy(&u, &c, 4);
    return (u & 0x7FFFFFu) != 0x7FFFFFu;
}
__device__ __forceinline__ int grid_cell_of(float sx, float sy, const GridTest& gt) {
    const float qx = grid_floor_div(__fsub_rn(sx, gt.minx), gt), qy = grid_floor_div(__fsub_rn(sy, gt.miny), gt);
    // one range test on the floats (NaN fails it), then 32-bit arithmetic: nx * ny fits 31 bits (checked on the host)
    if (!(qx >= 0.f && qy >= 0.f && qy < (float)gt.ny && qx < (float)gt.nx)) return -1;
    const uint32_t cid = (uint32_t)__float2int_rz(qx) * (uint32_t)gt.ny + (uint32_t)__float2int_rz(qy);
    return cid < (uint32_t)gt.n_cells ? (int)cid : -1;      // (float)nx may round up for grids wider than 2^24 cells
}
__device__ __forceinline__ bool grid_keep(float sx, float sy, float sz, const GridTest& gt) {
    const int cid = grid_cell_of(sx, sy, gt);
    const float g = cid >= 0 ? pch_ordered_to_f32(__ldg(gt.cell_min + cid)) : sz;
    return __fsub_rn(sz, g) > gt.hag;
}

// ------------------------------------------------------------------------------------------------
// order-preserving compaction of points with z_shifted > thr
// ------------------------------------------------------------------------------------------------
#define CP_THREADS 256
#define CP_ROWS 8
#define CP_TILE (CP_THREADS * CP_ROWS)

extern "C" size_t pch_compact_workspace_bytes(int64_t m) { return 256 + (size_t)(pch_ceil_div(m > 0 ? m : 1, CP_TILE)) * 8; }

// keep[i] = zs[i] > thr  (mode 0)   or   keep_mask[i] != 0 (mode 1, zs == nullptr)   or
// float32(z[i] - centroid.z) > thr from the cloud itself (mode 2, zs == keep_mask == nullptr)
__global__ void __launch_bounds__(CP_THREADS)
k_compact(const float* __restrict__ xyz, const float* __restrict__ zs, const uint8_t* __restrict__ keep_mask, GridTest gt,
          int64_t m, const float* __restrict__ centroid, float thr, float* __restrict__ out_xyz,
          int32_t* __restrict__ out_src, uint8_t* __restrict__ out_mask, long long* __restrict__ count_out,
          uint64_t* __restrict__ status, uint32_t* __restrict__ counter, int* __restrict__ err) {
    __shared__ uint32_t s_wcount[CP_THREADS / 32];
    __shared__ uint64_t s_off;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t n_tiles = (m + CP_TILE - 1) / CP_TILE;
    if (tile >= n_tiles) return;
    const int64_t start = tile * CP_TILE;
    const int wbase = warp * (32 * CP_ROWS);
    uint32_t rank[CP_ROWS];
    uint32_t keep_bits = 0, wtotal = 0;
    const float cz_keep = (!zs && !keep_mask && centroid) ? centroid[2] : 0.f;
    const float cx_keep = (gt.cell_min && centroid) ? centroid[0] : 0.f, cy_keep = (gt.cell_min && centroid) ? centroid[1] : 0.f;
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        int64_t i = start + wbase + j * 32 + lane;
        bool keep = false;
        if (i < m) {
            if (gt.cell_min) keep = grid_keep(__fsub_rn(xyz[i * 3 + 0], cx_keep), __fsub_rn(xyz[i * 3 + 1], cy_keep), __fsub_rn(xyz[i * 3 + 2], cz_keep), gt);
            else keep = zs ? (zs[i] > thr) : (keep_mask ? (keep_mask[i] != 0) : (__fsub_rn(xyz[i * 3 + 2], cz_keep) > thr));
        }
        if (out_mask && i < m) out_mask[i] = keep ? 1 : 0;
        uint32_t b = __ballot_sync(0xffffffffu, keep);
        rank[j] = wtotal + __popc(b & ((1u << lane) - 1u));
        wtotal += __popc(b);
        if (keep) keep_bits |= 1u << j;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();
    uint32_t wprefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        uint32_t c = s_wcount[w];
        if (w < warp) wprefix += c;
        total += c;
    }
    if (tid == 0) pch_lookback_publish_u64(status, tile, 0, total);
    // fetch the kept rows BEFORE the look-back walk: the chain wait then overlaps the loads, and only the
    // stores are left once the offset is known
    float vx[CP_ROWS], vy[CP_ROWS], vz[CP_ROWS];
    if (out_xyz) {
        const float cx = centroid ? centroid[0] : 0.f, cy = centroid ? centroid[1] : 0.f, cz = centroid ? centroid[2] : 0.f;
#pragma unroll
        for (int j = 0; j < CP_ROWS; ++j) {
            vx[j] = vy[j] = vz[j] = 0.f;
            if (keep_bits & (1u << j)) {
                const int64_t i = start + wbase + j * 32 + lane;
                vx[j] = __fsub_rn(xyz[i * 3 + 0], cx);
                vy[j] = __fsub_rn(xyz[i * 3 + 1], cy);
                vz[j] = __fsub_rn(xyz[i * 3 + 2], cz);
            }
        }
    }
    if (tid == 0) {
        s_off = pch_lookback_walk_u64(status, tile, 0, total, err);
        if (tile == n_tiles - 1) *count_out = (long long)(s_off + total);
    }
    __syncthreads();
    const uint64_t off = s_off + wprefix;
    if (!out_xyz && !out_src) return;
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        if (!(keep_bits & (1u << j))) continue;
        int64_t i = start + wbase + j * 32 + lane;
        uint64_t o = off + rank[j];
        if (out_xyz) {
            out_xyz[o * 3 + 0] = vx[j];
            out_xyz[o * 3 + 1] = vy[j];
            out_xyz[o * 3 + 2] = vz[j];
        }
        if (out_src) out_src[o] = (int32_t)i;
    }
}

// Staged variant for the hot case (a float32 cloud in, compacted cloud out): the tile's 2048 rows are fetched
// with coalesced 16-byte loads into shared memory, the kept rows are packed there (after every warp has its
// rows in registers), and the packed span leaves with fully coalesced stores — instead of 4-byte accesses at
// a 12-byte stride on both sides.  Same flags, same order, same outputs as k_compact.
template <bool GRID>
__global__ void __launch_bounds__(CP_THREADS)
k_compact_xyz(const float* __restrict__ xyz, const float* __restrict__ zs, const uint8_t* __restrict__ keep_mask, GridTest gt,
              int64_t m, const float* __restrict__ centroid, float thr, float* __restrict__ out_xyz,
              int32_t* __restrict__ out_src, uint8_t* __restrict__ out_mask, long long* __restrict__ count_out,
              uint64_t* __restrict__ status, uint32_t* __restrict__ counter, int* __restrict__ err) {
    __shared__ __align__(16) float s_row[CP_TILE * 3];
    __shared__ uint32_t s_wcount[CP_THREADS / 32];
    __shared__ uint64_t s_off;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t n_tiles = (m + CP_TILE - 1) / CP_TILE;
    if (tile >= n_tiles) return;
    const int64_t start = tile * CP_TILE;
    const int cnt = (int)min((int64_t)CP_TILE, m - start);
    if (cnt == CP_TILE) {
        const float4* src = reinterpret_cast<const float4*>(xyz + start * 3);   // 24 576-byte tiles of a 16-byte aligned base
        float4* dst = reinterpret_cast<float4*>(s_row);
#pragma unroll
        for (int k = 0; k < CP_TILE * 3 / 4 / CP_THREADS; ++k) dst[tid + k * CP_THREADS] = __ldg(src + tid + k * CP_THREADS);
    } else {
        for (int w = tid; w < cnt * 3; w += CP_THREADS) s_row[w] = xyz[start * 3 + w];
    }
    __syncthreads();
    const float cx = centroid ? centroid[0] : 0.f, cy = centroid ? centroid[1] : 0.f, cz = centroid ? centroid[2] : 0.f;
    const int wbase = warp * (32 * CP_ROWS);
    uint32_t rank[CP_ROWS];
    uint32_t keep_bits = 0, wtotal = 0;
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        const int li = wbase + j * 32 + lane;
        const int64_t i = start + li;
        bool keep = false;
        if (li < cnt) {
            if (GRID) keep = grid_keep(__fsub_rn(s_row[li * 3 + 0], cx), __fsub_rn(s_row[li * 3 + 1], cy), __fsub_rn(s_row[li * 3 + 2], cz), gt);
            else keep = zs ? (zs[i] > thr) : (keep_mask ? (keep_mask[i] != 0) : (__fsub_rn(s_row[li * 3 + 2], cz) > thr));
        }
        if (out_mask && li < cnt) out_mask[i] = keep ? 1 : 0;
        uint32_t b = __ballot_sync(0xffffffffu, keep);
        rank[j] = wtotal + __popc(b & ((1u << lane) - 1u));
        wtotal += __popc(b);
        if (keep) keep_bits |= 1u << j;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    float vx[CP_ROWS], vy[CP_ROWS], vz[CP_ROWS];
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        const int li = wbase + j * 32 + lane;
        vx[j] = vy[j] = vz[j] = 0.f;
        if (keep_bits & (1u << j)) {
            vx[j] = __fsub_rn(s_row[li * 3 + 0], cx);
            vy[j] = __fsub_rn(s_row[li * 3 + 1], cy);
            vz[j] = __fsub_rn(s_row[li * 3 + 2], cz);
        }
    }
    __syncthreads();      // counts visible; every warp holds its kept rows in registers: s_row may be overwritten
    uint32_t wprefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        uint32_t c = s_wcount[w];
        if (w < warp) wprefix += c;
        total += c;
    }
    if (tid == 0) pch_lookback_publish_u64(status, tile, 0, total);
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j)
        if (keep_bits & (1u << j)) {
            const uint32_t p = wprefix + rank[j];
            s_row[p * 3 + 0] = vx[j]; s_row[p * 3 + 1] = vy[j]; s_row[p * 3 + 2] = vz[j];
        }
    if (tid == 0) {
        s_off = pch_lookback_walk_u64(status, tile, 0, total, err);
        if (tile == n_tiles - 1) *count_out = (long long)(s_off + total);
    }
    __syncthreads();
    const uint64_t off = s_off;
    if (out_xyz) {
        float* dst = out_xyz + off * 3;
        for (int w = tid; w < (int)total * 3; w += CP_THREADS) dst[w] = s_row[w];
    }
    if (out_src) {
#pragma unroll
        for (int j = 0; j < CP_ROWS; ++j)
            if (keep_bits & (1u << j)) out_src[off + wprefix + rank[j]] = (int32_t)(start + wbase + j * 32 + lane);
    }
}

// Index list of the set flags of a byte mask (DBSCAN cluster heads: a few thousand set bytes in tens of
// millions).  16 flags per 16-byte load, 8192 flags per tile, so the look-back chain is 4x shorter than with
// k_compact's 2048-element tiles and no row data is touched.
#define CF_ROWS 2
#define CF_TILE (CP_THREADS * 16 * CF_ROWS)
__global__ void __launch_bounds__(CP_THREADS)
k_compact_flags(const uint8_t* __restrict__ mask, int64_t m, int32_t* __restrict__ out_src, long long* __restrict__ count_out,
                uint64_t* __restrict__ status, uint32_t* __restrict__ counter, int* __restrict__ err) {
    __shared__ uint32_t s_wsum[CF_ROWS][CP_THREADS / 32];
    __shared__ uint64_t s_off;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t n_tiles = (m + CF_TILE - 1) / CF_TILE;
    if (tile >= n_tiles) return;
    const int64_t start = tile * CF_TILE;
    uint32_t w[CF_ROWS][4];
    uint32_t c[CF_ROWS], incl[CF_ROWS];
#pragma unroll
    for (int r = 0; r < CF_ROWS; ++r) {
        const int64_t i0 = start + (int64_t)r * (CP_THREADS * 16) + tid * 16;
        if (i0 + 16 <= m) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(mask + i0));
            w[r][0] = v.x; w[r][1] = v.y; w[r][2] = v.z; w[r][3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t x = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int64_t i = i0 + q * 4 + b;
                    if (i < m) x |= (uint32_t)mask[i] << (8 * b);
                }
                w[r][q] = x;
            }
        }
        uint32_t n = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // a byte is "set" when non-zero: fold every byte onto its bit 0
            uint32_t x = w[r][q];
            x |= x >> 4; x |= x >> 2; x |= x >> 1;
            x &= 0x01010101u;
            w[r][q] = x;
            n += __popc(x);
        }
        c[r] = n;
        uint32_t x = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        incl[r] = x;
        if (lane == 31) s_wsum[r][warp] = x;
    }
    __syncthreads();
    uint32_t total = 0, pre[CF_ROWS];
#pragma unroll
    for (int r = 0; r < CF_ROWS; ++r) {
        uint32_t wp = 0, rt = 0;
#pragma unroll
        for (int ww = 0; ww < CP_THREADS / 32; ++ww) {
            const uint32_t v = s_wsum[r][ww];
            if (ww < warp) wp += v;
            rt += v;
        }
        pre[r] = total + wp + incl[r] - c[r];
        total += rt;
    }
    if (tid == 0) {
        s_off = pch_lookback_u64(status, tile, 0, total, err);
        if (tile == n_tiles - 1) *count_out = (long long)(s_off + total);
    }
    __syncthreads();
    if (!out_src) return;
    const uint64_t off = s_off;
#pragma unroll
    for (int r = 0; r < CF_ROWS; ++r) {
        if (c[r] == 0) continue;
        uint64_t o = off + pre[r];
        const int64_t i0 = start + (int64_t)r * (CP_THREADS * 16) + tid * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (w[r][q] & (1u << (8 * b))) out_src[o++] = (int32_t)(i0 + q * 4 + b);
    }
}

static int compact_impl(const float* xyz, const float* zs, const uint8_t* keep_mask, GridTest gt, int64_t m,
                        const float* centroid3, float thr, float* out_xyz, int32_t* out_src,
                        uint8_t* out_mask, int64_t* count_dev, void* workspace, size_t workspace_bytes,
                        pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    PCH_CHECK_ARG(count_dev && workspace, "null pointer");
    PCH_CHECK_ARG(m < (1ll << 31), "more than 2^31-1 points per compaction");
    PCH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int64_t), st));
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz, "null pointer");
    size_t need = pch_compact_workspace_bytes(m);
    if (workspace_bytes < need) {
        pch_set_error("compact workspace too small: %zu < %zu", workspace_bytes, need);
        return PCH_ERR_WORKSPACE;
    }
    PCH_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    int* err = (int*)workspace;
    uint32_t* counter = (uint32_t*)((uint8_t*)workspace + 64);
    uint64_t* status = (uint64_t*)((uint8_t*)workspace + 256);
    int64_t tiles = pch_ceil_div(m, CP_TILE);
    if (keep_mask && !gt.cell_min && !zs && !out_xyz && !out_mask && (reinterpret_cast<uintptr_t>(keep_mask) & 15) == 0) {
        // flags -> index list only (needs CF_TILE/CP_TILE times fewer status words than were zeroed above)
        PCH_LAUNCH(st, "k_compact_flags", k_compact_flags<<<(unsigned)pch_ceil_div(m, CF_TILE), CP_THREADS, 0, st>>>(
                                              keep_mask, m, out_src, (long long*)count_dev, status, counter, err));
    } else if (out_xyz && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0) {
        if (gt.cell_min)
            PCH_LAUNCH(st, "k_compact_xyz", k_compact_xyz<true><<<(unsigned)tiles, CP_THREADS, 0, st>>>(xyz, zs, keep_mask, gt, m, centroid3, thr, out_xyz, out_src,
                                                                  out_mask, (long long*)count_dev, status, counter, err));
        else
            PCH_LAUNCH(st, "k_compact_xyz", k_compact_xyz<false><<<(unsigned)tiles, CP_THREADS, 0, st>>>(xyz, zs, keep_mask, gt, m, centroid3, thr, out_xyz, out_src,
                                                                  out_mask, (long long*)count_dev, status, counter, err));
    } else {
        PCH_LAUNCH(st, "k_compact", k_compact<<<(unsigned)tiles, CP_THREADS, 0, st>>>(xyz, zs, keep_mask, gt, m, centroid3, thr, out_xyz, out_src, out_mask,
                                                          (long long*)count_dev, status, counter, err));
    }
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_compact_points(const float* xyz, const float* zs, const uint8_t* keep_mask, int64_t m,
                                  const float* centroid3, float thr, float* out_xyz, int32_t* out_src,
                                  uint8_t* out_mask, int64_t* count_dev, void* workspace, size_t workspace_bytes,
                                  pch_stream_t stream) {
    GridTest gt;
    memset(&gt, 0, sizeof(gt));
    return compact_impl(xyz, zs, keep_mask, gt, m, centroid3, thr, out_xyz, out_src, out_mask, count_dev, workspace,
                        workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------
// grid min-z pass.  The cloud arrives in voxel order (x-major per chunk), so the 2048 points of a tile fall
// into a few dozen XY cells.  The tile's rows are staged in shared memory with 16-byte loads; each warp row
// groups its lanes by cell (match_any), takes the minimum of every group with ONE redux (warp-shuffle minimum)
// and its leader folds it into a per-CTA cell table in shared memory; per tile and cell, one global
// atomicMin leaves the CTA.  The centroid shift is applied on the fly: no shifted cloud is materialised.
// ------------------------------------------------------------------------------------------------
#define GM_THREADS 256
#define GM_ROWS 8
#define GM_TILE (GM_THREADS * GM_ROWS)
#define GM_SLOTS 1024
#define GM_PROBES 12

__global__ void __launch_bounds__(GM_THREADS)
k_grid_min(const float* __restrict__ xyz, int64_t m, const float* __restrict__ centroid, GridTest gt,
           uint32_t* __restrict__ cell_min) {
    __shared__ __align__(16) float s_row[GM_TILE * 3];
    __shared__ int s_key[GM_SLOTS];
    __shared__ uint32_t s_val[GM_SLOTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float cx = centroid ? centroid[0] : 0.f, cy = centroid ? centroid[1] : 0.f, cz = centroid ? centroid[2] : 0.f;
    const int64_t n_tiles = (m + GM_TILE - 1) / GM_TILE;
    const bool aligned = (reinterpret_cast<uintptr_t>(xyz) & 15) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t start = tile * GM_TILE;
        const int cnt = (int)min((int64_t)GM_TILE, m - start);
        if (cnt == GM_TILE && aligned) {
            const float4* src = reinterpret_cast<const float4*>(xyz + start * 3);   // 24 576-byte tiles of a 16-byte aligned base
            float4* dst = reinterpret_cast<float4*>(s_row);
#pragma unroll
            for (int k = 0; k < GM_TILE * 3 / 4 / GM_THREADS; ++k) dst[tid + k * GM_THREADS] = __ldg(src + tid + k * GM_THREADS);
        } else {
            for (int w = tid; w < cnt * 3; w += GM_THREADS) s_row[w] = xyz[start * 3 + w];
        }
        for (int k = tid; k < GM_SLOTS; k += GM_THREADS) { s_key[k] = -1; s_val[k] = 0xffffffffu; }
        __syncthreads();
        const int wbase = warp * (32 * GM_ROWS);
#pragma unroll
        for (int j = 0; j < GM_ROWS; ++j) {
            const int li = wbase + j * 32 + lane;
            int key = -1;
            uint32_t zo = 0xffffffffu;
            if (li < cnt) {
                const float sx = __fsub_rn(s_row[li * 3 + 0], cx), sy = __fsub_rn(s_row[li * 3 + 1], cy);
                key = grid_cell_of(sx, sy, gt);
                zo = pch_f32_to_ordered(__fsub_rn(s_row[li * 3 + 2], cz));
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            const uint32_t mn = __reduce_min_sync(peers, zo);                 // warp-shuffle minimum of the lanes in my cell
            if (key >= 0 && (peers & ((1u << lane) - 1u)) == 0) {             // the group's first lane owns the update
                uint32_t slot = ((uint32_t)key * 2654435761u) >> 22;          // GM_SLOTS = 2^10
                bool done = false;
#pragma unroll 1
                for (int pr = 0; pr < GM_PROBES && !done; ++pr) {
                    const int prev = atomicCAS(&s_key[slot], -1, key);
                    if (prev == -1 || prev == key) {
                        atomicMin(&s_val[slot], mn);
                        done = true;
                    }
                    slot = (slot + 1) & (GM_SLOTS - 1);
                }
                if (!done) atomicMin(&cell_min[key], mn);                     // table full (an unordered cloud): straight to global
            }
        }
        __syncthreads();
        for (int k = tid; k < GM_SLOTS; k += GM_THREADS) {
            const int key = s_key[k];
            if (key >= 0) atomicMin(&cell_min[key], s_val[k]);
        }
        __syncthreads();
    }
}

__global__ void k_grid_label(const float* __restrict__ xyz, int64_t m, const float* __restrict__ centroid, GridTest gt,
                             uint8_t* __restrict__ keep, float* __restrict__ ground_z) {
    const float cx = centroid ? centroid[0] : 0.f, cy = centroid ? centroid[1] : 0.f, cz = centroid ? centroid[2] : 0.f;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        const float x = __fsub_rn(xyz[i * 3 + 0], cx), y = __fsub_rn(xyz[i * 3 + 1], cy), z = __fsub_rn(xyz[i * 3 + 2], cz);
        const int cid = grid_cell_of(x, y, gt);
        const float g = cid >= 0 ? pch_ordered_to_f32(gt.cell_min[cid]) : z;
        if (ground_z) ground_z[i] = g;
        keep[i] = (__fsub_rn(z, g) > gt.hag) ? 1 : 0;
    }
}

static int grid_args(int64_t m, float cell, int32_t nx, int32_t ny, GridTest& gt, float minx, float miny, float hag,
                     const uint32_t* cell_min) {
    PCH_CHECK_ARG(m >= 0 && nx >= 1 && ny >= 1 && cell > 0.f, "bad grid");
    PCH_CHECK_ARG((int64_t)nx * ny <= 2147483647ll, "grid of %d x %d cells does not fit 31 bits", nx, ny);
    gt.cell_min = cell_min;
    gt.minx = minx; gt.miny = miny; gt.cell = cell; gt.hag = hag;
    gt.rcell = 1.0f / cell;
    gt.fast = grid_recip_ok(cell) ? 1 : 0;
    gt.nx = nx; gt.ny = ny;
    gt.n_cells = (long long)nx * ny;
    return PCH_OK;
}

extern "C" int pch_grid_min(const float* xyz, int64_t m, const float* centroid3, float minx, float miny, float cell,
                            int32_t nx, int32_t ny, uint32_t* cell_min, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GridTest gt;
    int rc = grid_args(m, cell, nx, ny, gt, minx, miny, 0.f, cell_min);
    if (rc) return rc;
    PCH_CHECK_ARG(cell_min && (m == 0 || xyz), "null pointer");
    PCH_CUDA(cudaMemsetAsync(cell_min, 0xff, (size_t)gt.n_cells * 4, st));
    if (m == 0) return PCH_OK;
    int64_t tiles = pch_ceil_div(m, GM_TILE);
    int64_t grid = (int64_t)pch_sm_count() * 6;
    if (grid > tiles) grid = tiles;
    PCH_LAUNCH(st, "k_grid_min", k_grid_min<<<(unsigned)grid, GM_THREADS, 0, st>>>(xyz, m, centroid3, gt, cell_min));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_compact_points_grid(const float* xyz, int64_t m, const float* centroid3, float minx, float miny,
                                       float cell, int32_t nx, int32_t ny, float hag, const uint32_t* cell_min,
                                       float* out_xyz, int32_t* out_src, uint8_t* out_mask, int64_t* count_dev,
                                       void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    GridTest gt;
    int rc = grid_args(m, cell, nx, ny, gt, minx, miny, hag, cell_min);
    if (rc) return rc;
    PCH_CHECK_ARG(cell_min, "null cell table");
    return compact_impl(xyz, nullptr, nullptr, gt, m, centroid3, 0.f, out_xyz, out_src, out_mask, count_dev, workspace,
                        workspace_bytes, stream);
}

extern "C" int pch_grid_min_ground(const float* xyz, int64_t m, float minx, float miny, float cell, int32_t nx,
                                   int32_t ny, float hag, uint32_t* cell_min, uint8_t* keep, float* ground_z,
                                   pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && cell_min && keep, "null pointer");
    int rc = pch_grid_min(xyz, m, nullptr, minx, miny, cell, nx, ny, cell_min, stream);
    if (rc) return rc;
    GridTest gt;
    if ((rc = grid_args(m, cell, nx, ny, gt, minx, miny, hag, cell_min))) return rc;
    PCH_LAUNCH(st, "k_grid_label", k_grid_label<<<grid_for(m, 256, 8), 256, 0, st>>>(xyz, m, nullptr, gt, keep, ground_z));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// self-test of the float32 reciprocal division used by the grid kernels: mismatches against __fdiv_rn
__global__ void k_selftest_fastdiv_f32(const float* __restrict__ a, int64_t n, GridTest gt, unsigned long long* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (; i < n; i += stride) mine += __float_as_uint(grid_div(a[i], gt)) != __float_as_uint(__fdiv_rn(a[i], gt.cell));
    if (mine) atomicAdd(bad, mine);
}
extern "C" int pch_selftest_fastdiv_f32(const float* a_dev, int64_t n, float b, int64_t* mismatches_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && mismatches_dev && (n == 0 || a_dev), "bad arguments");
    PCH_CHECK_ARG(grid_recip_ok(b), "divisor %g is not eligible for the reciprocal path", (double)b);
    PCH_CUDA(cudaMemsetAsync(mismatches_dev, 0, sizeof(int64_t), st));
    if (n == 0) return PCH_OK;
    GridTest gt;
    memset(&gt, 0, sizeof(gt));
    gt.cell = b; gt.rcell = 1.0f / b; gt.fast = 1;
    PCH_LAUNCH(st, "k_selftest_fastdiv_f32", k_selftest_fastdiv_f32<<<grid_for(n, 256, 8), 256, 0, st>>>(a_dev, n, gt, (unsigned long long*)mismatches_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// componentwise float32 min/max of an (m,3) array -> out6 (minx,miny,minz,maxx,maxy,maxz).  The flat array is read
// with 16-byte loads (three float4 = four rows per thread and step), so the pass runs at the streaming rate.
__global__ void k_minmax_f32(const float* __restrict__ xyz, int64_t m, uint32_t* __restrict__ out6) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int64_t n_quads = m / 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(xyz) & 15) == 0;
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (aligned) {
        const float4* src = reinterpret_cast<const float4*>(xyz);
        for (; q < n_quads; q += stride) {
            const float4 a = __ldg(src + q * 3), b = __ldg(src + q * 3 + 1), c = __ldg(src + q * 3 + 2);
            const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 12; ++k) { mn[k % 3] = fminf(mn[k % 3], v[k]); mx[k % 3] = fmaxf(mx[k % 3], v[k]); }
        }
        q = n_quads * 4 + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x);
    } else {
        q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    }
    for (int64_t i = q; i < m; i += stride) {      // the tail (or everything, for an unaligned base)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = xyz[i * 3 + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uint32_t a = pch_f32_to_ordered(mn[c]), b = pch_f32_to_ordered(mx[c]);
        a = __reduce_min_sync(0xffffffffu, a);
        b = __reduce_max_sync(0xffffffffu, b);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&out6[c], a);
            atomicMax(&out6[3 + c], b);
        }
    }
}
__global__ void k_minmax_f32_init(uint32_t* out6) {
    if (threadIdx.x < 3) out6[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) out6[threadIdx.x] = 0u;
}
__global__ void k_minmax_f32_fin(uint32_t* out6) {
    if (threadIdx.x < 6) reinterpret_cast<float*>(out6)[threadIdx.x] = pch_ordered_to_f32(out6[threadIdx.x]);
}

extern "C" int pch_f32_minmax(const float* xyz, int64_t m, float* out6, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 1 && xyz && out6, "bad arguments");
    PCH_LAUNCH(st, "k_minmax_f32_init", k_minmax_f32_init<<<1, 32, 0, st>>>((uint32_t*)out6));
    PCH_LAUNCH(st, "k_minmax_f32", k_minmax_f32<<<grid_for(m, 256, 8), 256, 0, st>>>(xyz, m, (uint32_t*)out6));
    PCH_LAUNCH(st, "k_minmax_f32_fin", k_minmax_f32_fin<<<1, 32, 0, st>>>((uint32_t*)out6));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
