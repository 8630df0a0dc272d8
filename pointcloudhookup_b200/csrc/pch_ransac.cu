// Tiled RANSAC ground removal (reference: test/main_ground.py:77-115 calling :8-32).
//
// The reference cuts the cloud into tile_size x tile_size squares with np.arange edges, and for every square
// with at least 10 points fits z = a*x + b*y + c with scikit-learn's RANSACRegressor (LinearRegression on 3
// random points per trial, |residual| <= distance_threshold, most inliers wins, ties by R^2, the trial budget
// shrinks with the best inlier ratio) — ground = inliers, non-ground = outliers, tile by tile.
//
// Here: every point gets the word  tile << 32 | index  (k_ransac_tile_words; points the reference's edge
// list leaves out go to a last bucket), the words are sorted with the library's radix sort and the rows
// gathered (k_gather_rows_f64), so each tile is one contiguous slice in the reference's `points[tile_mask]`
// order.  k_ransac_tiles runs ONE CTA per tile through scikit-learn's trial loop: thread 0 draws the three
// points (from a seeded counter-based generator, or from a caller-supplied list so a test can replay the very
// triples scikit-learn drew) and solves the plane in closed form, all threads count inliers and the sums R^2
// needs, thread 0 keeps the best.  k_ransac_split writes ground / non-ground rows in the reference's order.
//
// All plane arithmetic is individually rounded (no FMA contraction), in the order oracle/ransac.py uses.
#include "pch_common.cuh"
#include <math.h>

#define RZ_THREADS 256
#define RZ_WARPS (RZ_THREADS / 32)

// ------------------------------------------------------------------------------------------------
// xy bounding box of float64 rows
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long rz_d2ord(double d) {
    long long u = __double_as_longlong(d);
    return u < 0 ? (u ^ 0x7fffffffffffffffll) : u;     // order-preserving (signed compare)
}
__device__ __forceinline__ double rz_ord2d(long long u) {
    return __longlong_as_double(u < 0 ? (u ^ 0x7fffffffffffffffll) : u);
}

__global__ void k_xy_minmax_init(long long* __restrict__ acc) {
    if (threadIdx.x < 2) acc[threadIdx.x] = LLONG_MAX;
    else if (threadIdx.x < 4) acc[threadIdx.x] = LLONG_MIN;
}

__global__ void __launch_bounds__(RZ_THREADS) k_xy_minmax_f64(const double* __restrict__ P, int64_t n, long long* __restrict__ acc) {
    double mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double x = P[i * 3 + 0], y = P[i * 3 + 1];
        mnx = fmin(mnx, x); mxx = fmax(mxx, x);
        mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&acc[0], rz_d2ord(mnx));
        atomicMin(&acc[1], rz_d2ord(mny));
        atomicMax(&acc[2], rz_d2ord(mxx));
        atomicMax(&acc[3], rz_d2ord(mxy));
    }
}

__global__ void k_xy_minmax_done(const long long* __restrict__ acc, double* __restrict__ out) {
    if (threadIdx.x < 4) out[threadIdx.x] = rz_ord2d(acc[threadIdx.x]);
}

extern "C" int pch_xy_minmax_f64(const double* points_dev, int64_t n, double* out4_dev, void* scratch32_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n > 0 && points_dev && out4_dev && scratch32_dev, "pch_xy_minmax_f64: null pointer / empty input");
    long long* acc = (long long*)scratch32_dev;
    PCH_LAUNCH(st, "k_xy_minmax_init", k_xy_minmax_init<<<1, 32, 0, st>>>(acc));
    int64_t blocks = pch_ceil_div(n, RZ_THREADS * 4);
    const int64_t cap = (int64_t)pch_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    PCH_LAUNCH(st, "k_xy_minmax_f64", k_xy_minmax_f64<<<(unsigned)blocks, RZ_THREADS, 0, st>>>(points_dev, n, acc));
    PCH_LAUNCH(st, "k_xy_minmax_done", k_xy_minmax_done<<<1, 32, 0, st>>>(acc, out4_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// tile of a point: edges e[i] = start + i*delta exactly as np.arange fills them (one rounded multiply, one
// rounded add); the point belongs to i with e[i] <= x < e[i+1], i in [0, n_edges-1)
// ------------------------------------------------------------------------------------------------
struct RzAxis {
    double start, e1, delta;     // edges[0], edges[1] = start + step, delta = edges[1] - edges[0]
    int n_edges;
};
__device__ __forceinline__ double rz_edge(const RzAxis& a, int i) {
    if (i == 0) return a.start;
    if (i == 1) return a.e1;
    return __dadd_rn(a.start, __dmul_rn((double)i, a.delta));
}
__device__ __forceinline__ int rz_tile_of(double x, const RzAxis& a) {
    // x >= start always (start is the minimum); the guess is at most one off
    const double q = floor((x - a.start) / a.delta);
    int i = q > 0.0 ? (q > (double)(a.n_edges - 1) ? a.n_edges - 1 : (int)q) : 0;
    while (i > 0 && x < rz_edge(a, i)) --i;
    while (i < a.n_edges - 1 && x >= rz_edge(a, i + 1)) ++i;
    // i == n_edges-1: at or beyond the last edge, no tile (the reference's loops stop at len(edges)-1)
    if (x < rz_edge(a, i)) return -1;
    return i < a.n_edges - 1 ? i : -1;
}

__global__ void __launch_bounds__(RZ_THREADS)
k_ransac_tile_words(const double* __restrict__ P, int64_t n, RzAxis ax, RzAxis ay, uint64_t* __restrict__ words) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int nty = ay.n_edges - 1;
    const uint32_t none = (uint32_t)((ax.n_edges - 1) * nty);
    for (; i < n; i += stride) {
        const int ti = rz_tile_of(P[i * 3 + 0], ax);
        const int tj = rz_tile_of(P[i * 3 + 1], ay);
        const uint32_t t = (ti < 0 || tj < 0) ? none : (uint32_t)(ti * nty + tj);
        words[i] = ((uint64_t)t << 32) | (uint64_t)(uint32_t)i;
    }
}

extern "C" int pch_ransac_tile_words(const double* points_dev, int64_t n, const double* x_edges3, int32_t n_x_edges,
                                     const double* y_edges3, int32_t n_y_edges, uint64_t* words_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n > 0 && n < (1ll << 31) - 4096 && points_dev && words_dev && x_edges3 && y_edges3,
                  "pch_ransac_tile_words: bad arguments (n must stay below 2^31 - 4096: 32-bit row numbers)");
    RzAxis ax = {x_edges3[0], x_edges3[1], x_edges3[2], n_x_edges}, ay = {y_edges3[0], y_edges3[1], y_edges3[2], n_y_edges};
    PCH_CHECK_ARG(n_x_edges >= 2 && n_y_edges >= 2 && ax.delta > 0 && ay.delta > 0, "pch_ransac_tile_words: no tiles");
    PCH_CHECK_ARG((int64_t)(n_x_edges - 1) * (n_y_edges - 1) < (1ll << 31) - 1, "pch_ransac_tile_words: too many tiles");
    int64_t blocks = pch_ceil_div(n, RZ_THREADS * 2);
    const int64_t cap = (int64_t)pch_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    PCH_LAUNCH(st, "k_ransac_tile_words", k_ransac_tile_words<<<(unsigned)blocks, RZ_THREADS, 0, st>>>(
        points_dev, n, ax, ay, words_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

__global__ void __launch_bounds__(RZ_THREADS)
k_gather_rows_f64(const double* __restrict__ P, const uint64_t* __restrict__ words, int64_t m, double* __restrict__ out) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < m; j += stride) {
        const uint64_t i = words[j] & 0xffffffffull;
        out[j * 3 + 0] = P[i * 3 + 0];
        out[j * 3 + 1] = P[i * 3 + 1];
        out[j * 3 + 2] = P[i * 3 + 2];
    }
}

extern "C" int pch_gather_rows_f64(const double* points_dev, const uint64_t* words_dev, int64_t m, double* out_dev,
                                   pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(points_dev && words_dev && out_dev, "pch_gather_rows_f64: null pointer");
    int64_t blocks = pch_ceil_div(m, RZ_THREADS);
    const int64_t cap = (int64_t)pch_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    PCH_LAUNCH(st, "k_gather_rows_f64", k_gather_rows_f64<<<(unsigned)blocks, RZ_THREADS, 0, st>>>(points_dev, words_dev, m, out_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// the trial loop of sklearn.linear_model.RANSACRegressor.fit, one CTA per tile
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t rz_mix(uint64_t z) {      // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct RzPlane {
    double x0, y0, z0, a, b;
};

// residual of one point against a plane anchored at the first sample: (z - z0) - (a*(x - x0) + b*(y - y0))
__device__ __forceinline__ double rz_residual(const RzPlane& p, double x, double y, double z) {
    const double t1 = __dmul_rn(p.a, __dsub_rn(x, p.x0));
    const double t2 = __dmul_rn(p.b, __dsub_rn(y, p.y0));
    return fabs(__dsub_rn(__dsub_rn(z, p.z0), __dadd_rn(t1, t2)));
}

// _dynamic_max_trials(n_inliers, n_samples, min_samples=3, probability) of sklearn/linear_model/_ransac.py
__device__ double rz_dynamic_max_trials(int n_inliers, int n_samples, double probability) {
    const double eps = 2.220446049250313e-16;
    const double ratio = (double)n_inliers / (double)n_samples;
    const double nom = fmax(eps, 1.0 - probability);
    const double denom = fmax(eps, 1.0 - pow(ratio, 3.0));
    if (nom == 1.0) return 0.0;
    if (denom == 1.0) return INFINITY;
    return fabs(ceil(log(nom) / log(denom)));
}

__global__ void __launch_bounds__(RZ_THREADS)
k_ransac_tiles(const double* __restrict__ P, const int64_t* __restrict__ bounds, int32_t n_tiles, double thr, int32_t max_trials,
               double stop_probability, uint64_t seed, const int32_t* __restrict__ triples, int32_t min_tile_points,
               uint8_t* __restrict__ flags, pch_ransac_tile* __restrict__ res) {
    const int t = blockIdx.x;
    if (t >= n_tiles) return;
    const int64_t lo = bounds[t], hi = bounds[t + 1];
    const int n = (int)(hi - lo);
    const double* Q = P + lo * 3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    __shared__ RzPlane s_plane, s_best;
    __shared__ int s_state;                       // 0 stop, 1 evaluate, 2 skip this trial (degenerate sample)
    __shared__ int s_cnt[RZ_WARPS];
    __shared__ double s_sum[RZ_WARPS][3];

    if (n < min_tile_points) {                    // `if len(tile_points) < 10: continue` — in neither output
        for (int i = tid; i < n; i += RZ_THREADS) flags[lo + i] = 2;
        if (tid == 0) {
            pch_ransac_tile r;
            memset(&r, 0, sizeof(r));
            r.n_points = n;
            r.status = 1;
            res[t] = r;
        }
        return;
    }

    // thread 0's private view of the loop state
    int n_best = 1, trials = 0, have_best = 0;
    double score_best = -INFINITY, max_tr = (double)max_trials;

    while (true) {
        if (tid == 0) {
            int state = 0;
            if ((double)trials < max_tr) {
                ++trials;
                int i1, i2, i3;
                if (triples) {
                    const int32_t* tp = triples + ((size_t)t * max_trials + (trials - 1)) * 3;
                    i1 = tp[0]; i2 = tp[1]; i3 = tp[2];
                } else {
                    uint64_t s = rz_mix(rz_mix(rz_mix(seed) ^ (uint64_t)t) + (uint64_t)trials);
                    s = rz_mix(s); i1 = (int)(s % (uint64_t)n);
                    do { s = rz_mix(s); i2 = (int)(s % (uint64_t)n); } while (i2 == i1);
                    do { s = rz_mix(s); i3 = (int)(s % (uint64_t)n); } while (i3 == i1 || i3 == i2);
                }
                state = 2;
                if (i1 >= 0 && i1 < n && i2 >= 0 && i2 < n && i3 >= 0 && i3 < n) {
                    const double* q1 = Q + (size_t)i1 * 3;
                    const double* q2 = Q + (size_t)i2 * 3;
                    const double* q3 = Q + (size_t)i3 * 3;
                    const double x1 = q1[0], y1 = q1[1], z1 = q1[2];
                    const double ux = __dsub_rn(q2[0], x1), uy = __dsub_rn(q2[1], y1), uz = __dsub_rn(q2[2], z1);
                    const double vx = __dsub_rn(q3[0], x1), vy = __dsub_rn(q3[1], y1), vz = __dsub_rn(q3[2], z1);
                    const double det = __dsub_rn(__dmul_rn(ux, vy), __dmul_rn(uy, vx));
                    if (det != 0.0 && isfinite(det)) {
                        RzPlane p;
                        p.x0 = x1; p.y0 = y1; p.z0 = z1;
                        p.a = __ddiv_rn(__dsub_rn(__dmul_rn(uz, vy), __dmul_rn(uy, vz)), det);
                        p.b = __ddiv_rn(__dsub_rn(__dmul_rn(ux, vz), __dmul_rn(uz, vx)), det);
                        s_plane = p;
                        state = 1;
                    }
                }
            }
            s_state = state;
        }
        __syncthreads();
        const int state = s_state;
        if (state == 0) break;
        if (state == 1) {
            const RzPlane p = s_plane;
            int cnt = 0;
            double sz = 0.0, szz = 0.0, srr = 0.0;
            for (int i = tid; i < n; i += RZ_THREADS) {
                const double* q = Q + (size_t)i * 3;
                const double x = q[0], y = q[1], z = q[2];
                const double r = rz_residual(p, x, y, z);
                if (r <= thr) {
                    const double zz = z - p.z0;
                    ++cnt;
                    sz += zz;
                    szz += zz * zz;
                    srr += r * r;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cnt += __shfl_down_sync(0xffffffffu, cnt, o);
                sz += __shfl_down_sync(0xffffffffu, sz, o);
                szz += __shfl_down_sync(0xffffffffu, szz, o);
                srr += __shfl_down_sync(0xffffffffu, srr, o);
            }
            if (lane == 0) {
                s_cnt[warp] = cnt;
                s_sum[warp][0] = sz; s_sum[warp][1] = szz; s_sum[warp][2] = srr;
            }
        }
        __syncthreads();
        if (tid == 0 && state == 1) {
            int cnt = 0;
            double sz = 0.0, szz = 0.0, srr = 0.0;
            for (int w = 0; w < RZ_WARPS; ++w) {
                cnt += s_cnt[w];
                sz += s_sum[w][0]; szz += s_sum[w][1]; srr += s_sum[w][2];
            }
            if (cnt >= n_best) {                                   // fewer inliers: skip
                // r2_score on the inliers: 1 - SS_res / SS_tot (a constant target scores 1 if reproduced, else 0)
                const double sstot = szz - sz * sz / (double)cnt;
                double score;
                if (sstot > 0.0) score = 1.0 - srr / sstot;
                else score = srr == 0.0 ? 1.0 : 0.0;
                if (!(cnt == n_best && score < score_best)) {      // same count but worse score: skip
                    n_best = cnt;
                    score_best = score;
                    have_best = 1;
                    s_best = s_plane;
                    max_tr = fmin(max_tr, rz_dynamic_max_trials(n_best, n, stop_probability));
                }
            }
        }
        // thread 0 writes s_plane / s_state of the next trial only after this point; everybody else has read them
        // before the barrier above, and reads them again only after the next one
    }

    __shared__ int s_have;
    if (tid == 0) s_have = have_best;
    __syncthreads();
    const bool have = s_have != 0;
    const RzPlane b = s_best;
    for (int i = tid; i < n; i += RZ_THREADS) {
        uint8_t f = 2;
        const double* q = Q + (size_t)i * 3;
        if (have) f = rz_residual(b, q[0], q[1], q[2]) <= thr ? 1 : 0;
        flags[lo + i] = f;
    }
    if (tid == 0) {
        pch_ransac_tile r;
        memset(&r, 0, sizeof(r));
        r.n_points = n;
        r.n_trials = trials;
        r.n_inliers = have_best ? n_best : 0;
        r.status = have_best ? 0 : 2;         // 2: no valid consensus set (scikit-learn raises ValueError)
        r.anchor[0] = b.x0; r.anchor[1] = b.y0; r.anchor[2] = b.z0;
        r.slope[0] = b.a; r.slope[1] = b.b;
        r.score = score_best;
        res[t] = r;
    }
}

extern "C" int pch_ransac_tiles(const double* tile_points_dev, const int64_t* bounds_dev, int32_t n_tiles,
                                double distance_threshold, int32_t max_trials, double stop_probability, uint64_t seed,
                                const int32_t* triples_dev, int32_t min_tile_points, uint8_t* flags_dev,
                                pch_ransac_tile* tiles_out_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n_tiles >= 0 && max_trials >= 0 && min_tile_points >= 3, "pch_ransac_tiles: bad arguments");
    PCH_CHECK_ARG(distance_threshold >= 0 && stop_probability >= 0 && stop_probability <= 1, "pch_ransac_tiles: bad threshold / probability");
    if (n_tiles == 0) return PCH_OK;
    PCH_CHECK_ARG(tile_points_dev && bounds_dev && flags_dev && tiles_out_dev, "pch_ransac_tiles: null pointer");
    PCH_LAUNCH(st, "k_ransac_tiles", k_ransac_tiles<<<(unsigned)n_tiles, RZ_THREADS, 0, st>>>(
        tile_points_dev, bounds_dev, n_tiles, distance_threshold, max_trials, stop_probability, seed, triples_dev,
        min_tile_points, flags_dev, tiles_out_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// np.vstack of every tile's ground / non-ground rows, tile by tile, each in `points[tile_mask]` order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RZ_THREADS)
k_ransac_split(const double* __restrict__ P, const uint8_t* __restrict__ flags, const int64_t* __restrict__ bounds, int32_t n_tiles,
               const int64_t* __restrict__ ground_off, const int64_t* __restrict__ other_off, double* __restrict__ ground_out,
               double* __restrict__ other_out) {
    const int t = blockIdx.x;
    if (t >= n_tiles) return;
    const int64_t lo = bounds[t], hi = bounds[t + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_w[2][RZ_WARPS];
    int64_t g_base = ground_off[t], o_base = other_off[t];
    for (int64_t base = lo; base < hi; base += RZ_THREADS) {
        const int64_t i = base + tid;
        const int f = i < hi ? (int)flags[i] : 2;
        const unsigned mg = __ballot_sync(0xffffffffu, f == 1), mo = __ballot_sync(0xffffffffu, f == 0);
        if (lane == 0) {
            s_w[0][warp] = __popc(mg);
            s_w[1][warp] = __popc(mo);
        }
        __syncthreads();
        int g_before = 0, o_before = 0, g_tot = 0, o_tot = 0;
#pragma unroll
        for (int w = 0; w < RZ_WARPS; ++w) {
            if (w < warp) { g_before += s_w[0][w]; o_before += s_w[1][w]; }
            g_tot += s_w[0][w];
            o_tot += s_w[1][w];
        }
        const unsigned below = (1u << lane) - 1u;
        if (f == 1) {
            const int64_t d = g_base + g_before + __popc(mg & below);
            ground_out[d * 3] = P[i * 3]; ground_out[d * 3 + 1] = P[i * 3 + 1]; ground_out[d * 3 + 2] = P[i * 3 + 2];
        } else if (f == 0) {
            const int64_t d = o_base + o_before + __popc(mo & below);
            other_out[d * 3] = P[i * 3]; other_out[d * 3 + 1] = P[i * 3 + 1]; other_out[d * 3 + 2] = P[i * 3 + 2];
        }
        g_base += g_tot;
        o_base += o_tot;
        __syncthreads();
    }
}

extern "C" int pch_ransac_split(const double* tile_points_dev, const uint8_t* flags_dev, const int64_t* bounds_dev, int32_t n_tiles,
                                const int64_t* ground_off_dev, const int64_t* other_off_dev, double* ground_out_dev,
                                double* other_out_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n_tiles >= 0, "pch_ransac_split: bad arguments");
    if (n_tiles == 0) return PCH_OK;
    PCH_CHECK_ARG(tile_points_dev && flags_dev && bounds_dev && ground_off_dev && other_off_dev && ground_out_dev && other_out_dev,
                  "pch_ransac_split: null pointer");
    PCH_LAUNCH(st, "k_ransac_split", k_ransac_split<<<(unsigned)n_tiles, RZ_THREADS, 0, st>>>(
        tile_points_dev, flags_dev, bounds_dev, n_tiles, ground_off_dev, other_off_dev, ground_out_dev, other_out_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
