// Geoid height shift (PROJ `+proj=vgridshift` on a GTX grid) and EPSG:4547 -> EPSG:4326
// (inverse Gauss-Krueger, PROJ extended tmerc) per point, float64.
// Reference call sites: utils/elevation_converter.py:29-31,48,57-68 (multiplier +1, per tower);
// crs.py:25-35 (EGM96, multiplier -1); utils/table_match_gim.py:72-75,232 and
// test/005test.py:37,55 (Transformer.from_crs("EPSG:4547","EPSG:4326",always_xy=True), per point).
// The fused LAS kernel streams raw records through the bulk-TMA tile ring AND stages the geoid grid
// window of the tile in shared memory with row-wise bulk TMA copies.
#include "pch_common.cuh"
#include "pch_tiles.cuh"
#include <string.h>
#include <stdlib.h>
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#define PCH_GEOID_NODATA (-88.8888f)

__device__ __forceinline__ double geo_nan() { return __longlong_as_double(0x7ff8000000000000ll); }

// Bilinear geoid height; `node(r, c)` fetches the float32 grid node.  Mirrors oracle/geoid.py
// operation for operation (left-to-right float64 sums, no FMA contraction).
template <class NODE>
__device__ __forceinline__ double geoid_bilinear(const pch_geoid_grid& g, double lat, double lon, NODE node) {
    double dl = __dsub_rn(lon, g.ll_lon);
    dl = __dsub_rn(dl, __dmul_rn(floor(__ddiv_rn(dl, 360.0)), 360.0));
    const double gx = __ddiv_rn(dl, g.dlon);
    const double gy = __ddiv_rn(__dsub_rn(lat, g.ll_lat), g.dlat);
    long long ix = (long long)floor(gx), iy = (long long)floor(gy);
    const double fx = __dsub_rn(gx, (double)ix), fy = __dsub_rn(gy, (double)iy);
    const bool bad = !(gy >= 0.0) || gy > (double)(g.rows - 1) || (!g.is_global && gx > (double)(g.cols - 1)) ||
                     !(gx == gx);
    if (bad) return geo_nan();
    if (iy < 0) iy = 0;
    if (iy > g.rows - 1) iy = g.rows - 1;
    if (g.is_global) { ix %= g.cols; if (ix < 0) ix += g.cols; }
    else { if (ix < 0) ix = 0; if (ix > g.cols - 1) ix = g.cols - 1; }
    long long ix2 = ix + 1;
    if (ix2 >= g.cols) ix2 = g.is_global ? 0 : g.cols - 1;
    long long iy2 = iy + 1 < g.rows ? iy + 1 : g.rows - 1;
    const float f00 = node((int)iy, (int)ix), f01 = node((int)iy, (int)ix2);
    const float f10 = node((int)iy2, (int)ix), f11 = node((int)iy2, (int)ix2);
    const double ofx = __dsub_rn(1.0, fx), ofy = __dsub_rn(1.0, fy);
    const double w00 = __dmul_rn(ofx, ofy), w01 = __dmul_rn(fx, ofy), w10 = __dmul_rn(ofx, fy), w11 = __dmul_rn(fx, fy);
    // PROJ's nodata rule (grids.cpp, vertical grid value): nodata corners are left out, the sum is divided by
    // the weight of the valid ones, and only a cell with no valid corner has no value.
    double n = 0.0, tw = 0.0;
    int nw = 0;
    if (f00 != PCH_GEOID_NODATA) { n = __dadd_rn(n, __dmul_rn(w00, (double)f00)); tw = __dadd_rn(tw, w00); ++nw; }
    if (f01 != PCH_GEOID_NODATA) { n = __dadd_rn(n, __dmul_rn(w01, (double)f01)); tw = __dadd_rn(tw, w01); ++nw; }
    if (f10 != PCH_GEOID_NODATA) { n = __dadd_rn(n, __dmul_rn(w10, (double)f10)); tw = __dadd_rn(tw, w10); ++nw; }
    if (f11 != PCH_GEOID_NODATA) { n = __dadd_rn(n, __dmul_rn(w11, (double)f11)); tw = __dadd_rn(tw, w11); ++nw; }
    if (nw == 0) return geo_nan();
    if (nw != 4) n = __ddiv_rn(n, tw);
    return n;
}

__global__ void k_geoid_shift(const double* __restrict__ lat, const double* __restrict__ lon, const double* __restrict__ h,
                              int64_t n, const float* __restrict__ grid, pch_geoid_grid g, double mult,
                              double* __restrict__ out_h, double* __restrict__ out_n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double N = geoid_bilinear(g, lat[i], lon[i], [&](int r, int c) { return __ldg(&grid[(size_t)r * g.pitch + c]); });
        if (out_n) out_n[i] = N;
        if (out_h) out_h[i] = __dadd_rn(h[i], __dmul_rn(mult, N));
    }
}

// ------------------------------------------------------------------------------------------------
// inverse transverse Mercator (Krueger series to n^6, complex Clenshaw summation)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gk_inverse_point(const pch_tm_params& p, double x, double y, double& lon_deg, double& lat_deg) {
    const double xi = (y - p.fn) / (p.k0 * p.rect_radius);
    const double eta = (x - p.fe) / (p.k0 * p.rect_radius);
    double s2, c2;
    sincos(2.0 * xi, &s2, &c2);
    const double sh2 = sinh(2.0 * eta), ch2 = cosh(2.0 * eta);
    // theta = 2(xi + i eta): cos(theta) = c2*ch2 - i s2*sh2 ; sin(theta) = s2*ch2 + i c2*sh2
    const double cr = c2 * ch2, ci = -s2 * sh2;
    const double sr = s2 * ch2, si = c2 * sh2;
    // Clenshaw: b_k = beta_k + 2 cos(theta) b_{k+1} - b_{k+2};  sum_{j} beta_j sin(j theta) = b_1 sin(theta)
    double b1r = 0.0, b1i = 0.0, b2r = 0.0, b2i = 0.0;
#pragma unroll
    for (int k = 5; k >= 0; --k) {
        const double tr = 2.0 * (cr * b1r - ci * b1i) - b2r + p.beta[k];
        const double ti = 2.0 * (cr * b1i + ci * b1r) - b2i;
        b2r = b1r; b2i = b1i;
        b1r = tr; b1i = ti;
    }
    const double dxi = b1r * sr - b1i * si;
    const double deta = b1r * si + b1i * sr;
    const double xip = xi - dxi, etap = eta - deta;
    double sx, cx;
    sincos(xip, &sx, &cx);
    const double she = sinh(etap);
    const double taup = sx / sqrt(she * she + cx * cx);
    const double lam = atan2(she, cx);
    double tau = taup;
    const double e = p.ecc, e2m = 1.0 - e * e;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const double t1 = sqrt(1.0 + tau * tau);
        const double sigma = sinh(e * atanh(e * tau / t1));
        const double taui = tau * sqrt(1.0 + sigma * sigma) - sigma * t1;
        const double dtau = (taup - taui) / sqrt(1.0 + taui * taui) * (1.0 + e2m * tau * tau) / (e2m * t1);
        tau += dtau;
    }
    const double rad2deg = 57.295779513082320876798154814105;
    lon_deg = p.lon0_deg + lam * rad2deg;
    lat_deg = atan(tau) * rad2deg;
}

__global__ void k_gk_inverse(const double* __restrict__ x, const double* __restrict__ y, int64_t n, pch_tm_params p,
                             double* __restrict__ lon, double* __restrict__ lat) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double lo, la;
        gk_inverse_point(p, x[i], y[i], lo, la);
        lon[i] = lo;
        lat[i] = la;
    }
}

static unsigned geo_grid(int64_t n, int threads, int per_sm) {
    int64_t b = pch_ceil_div(n > 0 ? n : 1, threads);
    int64_t cap = (int64_t)pch_sm_count() * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

static int check_grid(const pch_geoid_grid* g) {
    PCH_CHECK_ARG(g != nullptr, "null grid descriptor");
    PCH_CHECK_ARG(g->rows >= 2 && g->cols >= 2 && g->pitch >= g->cols, "bad grid shape");
    PCH_CHECK_ARG(g->dlat > 0.0 && g->dlon > 0.0, "grid steps must be > 0");
    return PCH_OK;
}

extern "C" int pch_geoid_shift(const double* lat, const double* lon, const double* h, int64_t n, const float* grid,
                               const pch_geoid_grid* g, double multiplier, double* out_h, double* out_n,
                               pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0, "n must be >= 0");
    int rc = check_grid(g);
    if (rc) return rc;
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(lat && lon && grid && (out_n || (h && out_h)), "null pointer");
    PCH_LAUNCH(st, "k_geoid_shift", k_geoid_shift<<<geo_grid(n, 256, 8), 256, 0, st>>>(lat, lon, h, n, grid, *g, multiplier, out_h, out_n));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_gk_inverse(const double* x, const double* y, int64_t n, const pch_tm_params* p, double* lon,
                              double* lat, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && p != nullptr, "bad arguments");
    PCH_CHECK_ARG(p->rect_radius > 0.0 && p->k0 > 0.0, "bad projection constants");
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(x && y && lon && lat, "null pointer");
    PCH_LAUNCH(st, "k_gk_inverse", k_gk_inverse<<<geo_grid(n, 128, 16), 128, 0, st>>>(x, y, n, *p, lon, lat));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// fused: raw LAS records -> (lon, lat, H) per point
// ------------------------------------------------------------------------------------------------
struct GeoWindow {
    int32_t row0, col0, rows, cols;  // cols and col0 multiples of 4 (16-byte bulk copies); rows*cols floats in smem
};

struct GeoAffine {
    double s[3], o[3];
};

template <int ALIGN, bool STAGED>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 1)
k_las_geodetic(const uint8_t* __restrict__ rec, PchTileGeom tg, GeoAffine a, pch_tm_params tm, const float* __restrict__ grid,
               pch_geoid_grid g, GeoWindow w, double mult, int use_crs, double* __restrict__ out /*(n,3)*/,
               const __grid_constant__ CUtensorMap tmap, int use_tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    const size_t tile_bytes = 128 + (size_t)PCH_STAGES * tg.stage_bytes;
    float* s_win = reinterpret_cast<float*>(smem + tile_bytes + 128);      // 128-byte aligned: a tensor-map destination
    uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + tile_bytes);
    if (STAGED) {
        if (threadIdx.x == 0) {
            pch_mbar_init(wbar, 1);
            pch_fence_mbar_init();
            const uint32_t row_bytes = (uint32_t)w.cols * 4u;
            pch_mbar_arrive_expect_tx(wbar, row_bytes * (uint32_t)w.rows);
            if (use_tmap) {
                // the whole window of the geoid grid in ONE 2-D tensor-map copy (cp.async.bulk.tensor.2d -> SASS UTMALDG):
                // box = (cols, rows) at coordinate (col0, row0) of the (pitch x rows) float32 grid
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(pch_smem_u32(s_win)), "l"(&tmap), "r"(w.col0), "r"(w.row0), "r"(pch_smem_u32(wbar))
                             : "memory");
            } else {
                for (int r = 0; r < w.rows; ++r)
                    pch_tma_load_1d(s_win + (size_t)r * w.cols, grid + (size_t)(w.row0 + r) * g.pitch + w.col0, row_bytes, wbar);
            }
        }
        __syncthreads();
        pch_mbar_wait(wbar, 0);
    }
    pch_stream_tiles(rec, tg, smem, [&](const PchTile& t) {
        for (int r = threadIdx.x; r < t.count; r += PCH_TILE_THREADS) {
            int X, Y, Z;
            pch_load_xyz<ALIGN>(t.base + (size_t)r * tg.rec_len, X, Y, Z);
            const double x = pch_scaled(X, a.s[0], a.o[0]);
            const double y = pch_scaled(Y, a.s[1], a.o[1]);
            const double h = pch_scaled(Z, a.s[2], a.o[2]);
            double lon = x, lat = y;
            if (use_crs) gk_inverse_point(tm, x, y, lon, lat);
            double N;
            if (STAGED) {
                N = geoid_bilinear(g, lat, lon, [&](int rr, int cc) -> float {
                    const int lr = rr - w.row0, lc = cc - w.col0;
                    if (lr >= 0 && lr < w.rows && lc >= 0 && lc < w.cols) return s_win[lr * w.cols + lc];
                    return __ldg(&grid[(size_t)rr * g.pitch + cc]);  // outside the staged window: still exact
                });
            } else {
                N = geoid_bilinear(g, lat, lon, [&](int rr, int cc) -> float { return __ldg(&grid[(size_t)rr * g.pitch + cc]); });
            }
            double* o = out + (t.r0 + r) * 3;
            o[0] = lon;
            o[1] = lat;
            o[2] = __dadd_rn(h, __dmul_rn(mult, N));
        }
    });
}

extern "C" int pch_las_geodetic(const uint8_t* rec, int64_t n, int32_t rec_len, const double* scales, const double* offsets,
                                const pch_tm_params* tm, const float* grid, const pch_geoid_grid* g, int32_t win_row0,
                                int32_t win_col0, int32_t win_rows, int32_t win_cols, double multiplier, double* out,
                                pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && rec_len >= 12 && rec_len <= 256, "bad n/rec_len");
    PCH_CHECK_ARG(scales && offsets, "null scales/offsets");
    int rc = check_grid(g);
    if (rc) return rc;
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(rec && grid && out, "null pointer");
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(rec) & 15) == 0 && (reinterpret_cast<uintptr_t>(grid) & 15) == 0,
                  "record and grid buffers must be 16-byte aligned");
    GeoAffine a;
    for (int i = 0; i < 3; ++i) { a.s[i] = scales[i]; a.o[i] = offsets[i]; }
    pch_tm_params tmv;
    int use_crs = tm != nullptr;
    if (tm) tmv = *tm; else memset(&tmv, 0, sizeof(tmv));
    GeoWindow w{win_row0, win_col0, win_rows, win_cols};
    bool staged = win_rows > 0 && win_cols > 0;
    if (staged) {
        PCH_CHECK_ARG(win_row0 >= 0 && win_col0 >= 0 && win_row0 + win_rows <= g->rows && win_col0 + win_cols <= g->pitch,
                      "geoid window outside the grid");
        PCH_CHECK_ARG((win_col0 % 4) == 0 && (win_cols % 4) == 0 && (g->pitch % 4) == 0,
                      "geoid window columns/pitch must be multiples of 4 floats");
        PCH_CHECK_ARG((size_t)win_rows * win_cols * 4 <= 64 * 1024 && (size_t)win_rows * win_cols * 4 < (1u << 20),
                      "geoid window larger than 64 KiB");
    }
    PchTileGeom tg = pch_tile_geom(n, rec_len, n);
    size_t smem = pch_tile_smem_bytes(tg) + 128 + (staged ? (size_t)win_rows * win_cols * 4 : 0);
    // 2-D tensor map of the grid, box = the window (each box side <= 256 elements); without it (older driver, larger
    // window) the window is staged row by row with 1-D bulk copies
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    int use_tmap = 0;
    if (staged && win_rows <= 256 && win_cols <= 256 && !getenv("PCH_GEO_NO_TENSORMAP")) {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[2] = {(cuuint64_t)g->pitch, (cuuint64_t)g->rows};
            const cuuint64_t gstride[1] = {(cuuint64_t)g->pitch * 4};
            const cuuint32_t box[2] = {(cuuint32_t)win_cols, (cuuint32_t)win_rows};
            const cuuint32_t estr[2] = {1, 1};
            if (((encode_fn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)grid, gdim, gstride, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                use_tmap = 1;
        } else {
            (void)cudaGetLastError();
        }
    }
    int grid_dim = pch_tile_grid(tg, smem * 2 <= 220 * 1024 ? 2 : 1);
    int al = pch_rec_align(rec_len);
#define LAUNCH_GEO(A, S)                                                                                         \
    do {                                                                                                         \
        PCH_CUDA(cudaFuncSetAttribute(k_las_geodetic<A, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        PCH_LAUNCH(st, "k_las_geodetic", k_las_geodetic<A, S><<<grid_dim, PCH_TILE_THREADS, smem, st>>>(rec, tg, a, tmv, grid, *g, w, multiplier, use_crs, out, tmap, use_tmap)); \
    } while (0)
    if (staged) {
        if (al == 4) LAUNCH_GEO(4, true); else if (al == 2) LAUNCH_GEO(2, true); else LAUNCH_GEO(1, true);
    } else {
        if (al == 4) LAUNCH_GEO(4, false); else if (al == 2) LAUNCH_GEO(2, false); else LAUNCH_GEO(1, false);
    }
#undef LAUNCH_GEO
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


// ------------------------------------------------------------------------------------------------
// tower-level matching: great-circle distance matrix between GIM towers and the
// converted point-cloud towers — utils/table_match_gim.py:17-34 (haversine, R = 6371 km) evaluated
// for every (i, j) of match_towers' double loop (:168-192).
// ------------------------------------------------------------------------------------------------
__global__ void k_haversine_matrix(const double* __restrict__ lat1, const double* __restrict__ lon1, int64_t n1,
                                   const double* __restrict__ lat2, const double* __restrict__ lon2, int64_t n2,
                                   double* __restrict__ out) {
    const double d2r = 0.017453292519943295;   // math.radians
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; t < n1 * n2; t += stride) {
        const int64_t i = t / n2, j = t - i * n2;
        const double la1 = lat1[i] * d2r, lo1 = lon1[i] * d2r, la2 = lat2[j] * d2r, lo2 = lon2[j] * d2r;
        const double dlat = la2 - la1, dlon = lo2 - lo1;
        const double s1 = sin(dlat / 2), s2 = sin(dlon / 2);
        const double a = s1 * s1 + cos(la1) * cos(la2) * s2 * s2;
        const double c = 2 * atan2(sqrt(a), sqrt(1 - a));
        out[t] = 6371.0 * c * 1000;
    }
}

extern "C" int pch_haversine_matrix(const double* lat1, const double* lon1, int64_t n1, const double* lat2,
                                    const double* lon2, int64_t n2, double* out, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n1 >= 0 && n2 >= 0, "negative size");
    if (n1 == 0 || n2 == 0) return PCH_OK;
    PCH_CHECK_ARG(lat1 && lon1 && lat2 && lon2 && out, "null pointer");
    PCH_LAUNCH(st, "k_haversine_matrix", k_haversine_matrix<<<geo_grid(n1 * n2, 128, 8), 128, 0, st>>>(lat1, lon1, n1, lat2, lon2, n2, out));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
