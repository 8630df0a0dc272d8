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

extern "C" int pch_f32_centroid(const float* xyz, int64_t m, float* sums3, float* centroid3, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 1, "centroid of an empty cloud");
    PCH_CHECK_ARG(xyz && sums3 && centroid3, "null pointer");
    PCH_LAUNCH(st, "k_seq_sum_serial", k_seq_sum_serial<<<1, 32, 0, st>>>(xyz, m, sums3));
    PCH_LAUNCH_CHECK();
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
// order-preserving compaction of points with z_shifted > thr
// ------------------------------------------------------------------------------------------------
#define CP_THREADS 256
#define CP_ROWS 8
#define CP_TILE (CP_THREADS * CP_ROWS)

extern "C" size_t pch_compact_workspace_bytes(int64_t m) { return 256 + (size_t)(pch_ceil_div(m > 0 ? m : 1, CP_TILE)) * 8; }

// keep[i] = zs[i] > thr  (mode 0)   or   keep_mask[i] != 0 (mode 1, zs == nullptr)
__global__ void __launch_bounds__(CP_THREADS)
k_compact(const float* __restrict__ xyz, const float* __restrict__ zs, const uint8_t* __restrict__ keep_mask,
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
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        int64_t i = start + wbase + j * 32 + lane;
        bool keep = false;
        if (i < m) keep = zs ? (zs[i] > thr) : (keep_mask[i] != 0);
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
    if (tid == 0) {
        s_off = pch_lookback_u64(status, tile, 0, total, err);
        if (tile == n_tiles - 1) *count_out = (long long)(s_off + total);
    }
    __syncthreads();
    const uint64_t off = s_off + wprefix;
    if (!out_xyz && !out_src) return;
    const float cx = centroid ? centroid[0] : 0.f, cy = centroid ? centroid[1] : 0.f, cz = centroid ? centroid[2] : 0.f;
#pragma unroll
    for (int j = 0; j < CP_ROWS; ++j) {
        if (!(keep_bits & (1u << j))) continue;
        int64_t i = start + wbase + j * 32 + lane;
        uint64_t o = off + rank[j];
        if (out_xyz) {
            out_xyz[o * 3 + 0] = __fsub_rn(xyz[i * 3 + 0], cx);
            out_xyz[o * 3 + 1] = __fsub_rn(xyz[i * 3 + 1], cy);
            out_xyz[o * 3 + 2] = __fsub_rn(xyz[i * 3 + 2], cz);
        }
        if (out_src) out_src[o] = (int32_t)i;
    }
}

extern "C" int pch_compact_points(const float* xyz, const float* zs, const uint8_t* keep_mask, int64_t m,
                                  const float* centroid3, float thr, float* out_xyz, int32_t* out_src,
                                  uint8_t* out_mask, int64_t* count_dev, void* workspace, size_t workspace_bytes,
                                  pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    PCH_CHECK_ARG(count_dev && workspace, "null pointer");
    PCH_CHECK_ARG(m < (1ll << 31), "more than 2^31-1 points per compaction");
    PCH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int64_t), st));
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && (zs || keep_mask), "null pointer");
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
    PCH_LAUNCH(st, "k_compact", k_compact<<<(unsigned)tiles, CP_THREADS, 0, st>>>(xyz, zs, keep_mask, m, centroid3, thr, out_xyz, out_src, out_mask,
                                                      (long long*)count_dev, status, counter, err));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// grid min-z ground model (north_star extension; no reference implementation exists)
//   cell = floor((xy - min_xy)/cell) in float32; ground_z = min z of the cell; keep = z-ground_z > hag
// ------------------------------------------------------------------------------------------------
__global__ void k_grid_min(const float* __restrict__ xyz, int64_t m, float minx, float miny, float cell, int ny,
                           int64_t n_cells, uint32_t* __restrict__ cell_min) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
        long long ix = (long long)floorf(__fdiv_rn(__fsub_rn(x, minx), cell));
        long long iy = (long long)floorf(__fdiv_rn(__fsub_rn(y, miny), cell));
        long long cid = ix * ny + iy;
        if (cid < 0 || cid >= n_cells) continue;
        atomicMin(&cell_min[cid], pch_f32_to_ordered(z));
    }
}

__global__ void k_grid_label(const float* __restrict__ xyz, int64_t m, float minx, float miny, float cell, int ny,
                             int64_t n_cells, const uint32_t* __restrict__ cell_min, float hag,
                             uint8_t* __restrict__ keep, float* __restrict__ ground_z) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        float x = xyz[i * 3 + 0], y = xyz[i * 3 + 1], z = xyz[i * 3 + 2];
        long long ix = (long long)floorf(__fdiv_rn(__fsub_rn(x, minx), cell));
        long long iy = (long long)floorf(__fdiv_rn(__fsub_rn(y, miny), cell));
        long long cid = ix * ny + iy;
        float g = z;
        if (cid >= 0 && cid < n_cells) g = pch_ordered_to_f32(cell_min[cid]);
        if (ground_z) ground_z[i] = g;
        keep[i] = (__fsub_rn(z, g) > hag) ? 1 : 0;
    }
}

extern "C" int pch_grid_min_ground(const float* xyz, int64_t m, float minx, float miny, float cell, int32_t nx,
                                   int32_t ny, float hag, uint32_t* cell_min, uint8_t* keep, float* ground_z,
                                   pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0 && nx >= 1 && ny >= 1 && cell > 0.f, "bad grid");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && cell_min && keep, "null pointer");
    int64_t n_cells = (int64_t)nx * ny;
    PCH_CUDA(cudaMemsetAsync(cell_min, 0xff, (size_t)n_cells * 4, st));
    unsigned grid = grid_for(m, 256, 8);
    PCH_LAUNCH(st, "k_grid_min", k_grid_min<<<grid, 256, 0, st>>>(xyz, m, minx, miny, cell, ny, n_cells, cell_min));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_grid_label", k_grid_label<<<grid, 256, 0, st>>>(xyz, m, minx, miny, cell, ny, n_cells, cell_min, hag, keep, ground_z));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// componentwise float32 min/max of an (m,3) array -> out6 (minx,miny,minz,maxx,maxy,maxz)
__global__ void k_minmax_f32(const float* __restrict__ xyz, int64_t m, uint32_t* __restrict__ out6) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
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
