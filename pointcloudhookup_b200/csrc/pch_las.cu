// LAS attribute decode / encode kernels (laspy's scaled views, restated on raw record bytes).
// Reference call sites: ui/import_PC.py:28,45-48,61-65; utils/tower_extraction.py:60-62,243-257;
// ui/extract.py:114-115,361-362.
#include "pch_common.cuh"
#include "pch_tiles.cuh"

static int check_rec_args(const void* rec, int64_t n, int32_t rec_len) {
    PCH_CHECK_ARG(n >= 0, "n must be >= 0 (got %lld)", (long long)n);
    PCH_CHECK_ARG(rec_len >= 12 && rec_len <= 256, "record length %d outside [12, 256]", rec_len);
    PCH_CHECK_ARG(n == 0 || rec != nullptr, "null record pointer");
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(rec) & 15) == 0, "record buffer must be 16-byte aligned");
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// chunk min/max on the int32 lattice
// ------------------------------------------------------------------------------------------------
__global__ void k_init_minmax(int32_t* mm, int64_t n_chunks) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_chunks * 6) mm[i] = ((i % 6) < 3) ? INT_MAX : INT_MIN;
}

template <int ALIGN>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 2)
k_chunk_minmax(const uint8_t* __restrict__ rec, PchTileGeom g, int32_t* __restrict__ mm, int4* __restrict__ xyz16) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_part[PCH_TILE_THREADS / 32][6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pch_stream_tiles(rec, g, smem, [&](const PchTile& t) {
        int mnx = INT_MAX, mny = INT_MAX, mnz = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, mxz = INT_MIN;
        for (int r = tid; r < t.count; r += PCH_TILE_THREADS) {
            int X, Y, Z;
            pch_load_xyz<ALIGN>(t.base + (size_t)r * g.rec_len, X, Y, Z);
            mnx = min(mnx, X); mny = min(mny, Y); mnz = min(mnz, Z);
            mxx = max(mxx, X); mxy = max(mxy, Y); mxz = max(mxz, Z);
            // optional 16-byte aligned copy of the lattice coordinates: every later pass (keys, reduce
            // gathers) reads this instead of the 2-byte aligned AoS records
            if (xyz16) xyz16[t.r0 + r] = make_int4(X, Y, Z, 0);
        }
        mnx = pch_warp_min(mnx); mny = pch_warp_min(mny); mnz = pch_warp_min(mnz);
        mxx = pch_warp_max(mxx); mxy = pch_warp_max(mxy); mxz = pch_warp_max(mxz);
        if (lane == 0) {
            s_part[warp][0] = mnx; s_part[warp][1] = mny; s_part[warp][2] = mnz;
            s_part[warp][3] = mxx; s_part[warp][4] = mxy; s_part[warp][5] = mxz;
        }
        __syncthreads();
        if (tid < 6) {
            int v = s_part[0][tid];
            for (int w = 1; w < PCH_TILE_THREADS / 32; ++w) v = (tid < 3) ? min(v, s_part[w][tid]) : max(v, s_part[w][tid]);
            if (tid < 3) atomicMin(&mm[t.chunk * 6 + tid], v);
            else atomicMax(&mm[t.chunk * 6 + tid], v);
        }
        // the streamer's trailing __syncthreads protects s_part for the next tile
    });
}

template <class K, class... Args>
static int launch_tiles(const char* name, K kernel, const PchTileGeom& g, int ctas_per_sm, cudaStream_t st,
                        const uint8_t* rec, Args... args) {
    size_t smem = pch_tile_smem_bytes(g);
    PCH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = pch_tile_grid(g, ctas_per_sm);
    PCH_LAUNCH(st, name, kernel<<<grid, PCH_TILE_THREADS, smem, st>>>(rec, g, args...));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

#define PCH_DISPATCH_ALIGN(rec_len, KERNEL, ...)                                  \
    do {                                                                          \
        int a__ = pch_rec_align(rec_len);                                         \
        int rc__;                                                                 \
        if (a__ == 4) rc__ = launch_tiles(#KERNEL, KERNEL<4>, __VA_ARGS__);       \
        else if (a__ == 2) rc__ = launch_tiles(#KERNEL, KERNEL<2>, __VA_ARGS__);  \
        else rc__ = launch_tiles(#KERNEL, KERNEL<1>, __VA_ARGS__);                \
        if (rc__ != PCH_OK) return rc__;                                          \
    } while (0)

extern "C" int pch_las_chunk_minmax(const uint8_t* rec, int64_t n, int32_t rec_len, int64_t chunk_size,
                                    int32_t* mm, int32_t* xyz16, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rec_args(rec, n, rec_len);
    if (rc) return rc;
    PCH_CHECK_ARG(chunk_size > 0, "chunk_size must be > 0");
    PCH_CHECK_ARG(mm != nullptr, "null minmax pointer");
    if (n == 0) return PCH_OK;
    PchTileGeom g = pch_tile_geom(n, rec_len, chunk_size);
    int64_t n_chunks = pch_ceil_div(n, g.chunk_size);
    PCH_LAUNCH(st, "k_init_minmax", k_init_minmax<<<(unsigned)pch_ceil_div(n_chunks * 6, 256), 256, 0, st>>>(mm, n_chunks));
    PCH_LAUNCH_CHECK();
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(xyz16) & 15) == 0, "xyz16 must be 16-byte aligned");
    PCH_DISPATCH_ALIGN(rec_len, k_chunk_minmax, g, 2, st, rec, mm, (int4*)xyz16);
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// decode to (n,3) float64 / float32
// ------------------------------------------------------------------------------------------------
struct PchAffine {
    double sx, sy, sz, ox, oy, oz;
};

template <int ALIGN, typename OUT>
__device__ __forceinline__ void decode_tile(const PchTile& t, const PchTileGeom& g, const PchAffine& a, OUT* __restrict__ out) {
    for (int r = threadIdx.x; r < t.count; r += PCH_TILE_THREADS) {
        int X, Y, Z;
        pch_load_xyz<ALIGN>(t.base + (size_t)r * g.rec_len, X, Y, Z);
        OUT* o = out + (t.r0 + r) * 3;
        o[0] = (OUT)pch_scaled(X, a.sx, a.ox);   // (float) cast rounds to nearest even like astype
        o[1] = (OUT)pch_scaled(Y, a.sy, a.oy);
        o[2] = (OUT)pch_scaled(Z, a.sz, a.oz);
    }
}

template <int ALIGN>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 2)
k_decode_f64(const uint8_t* __restrict__ rec, PchTileGeom g, PchAffine a, double* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    pch_stream_tiles(rec, g, smem, [&](const PchTile& t) { decode_tile<ALIGN, double>(t, g, a, out); });
}
template <int ALIGN>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 2)
k_decode_f32(const uint8_t* __restrict__ rec, PchTileGeom g, PchAffine a, float* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    pch_stream_tiles(rec, g, smem, [&](const PchTile& t) { decode_tile<ALIGN, float>(t, g, a, out); });
}

static int make_affine(const double* scales, const double* offsets, PchAffine& a) {
    PCH_CHECK_ARG(scales && offsets, "null scales/offsets");
    a.sx = scales[0]; a.sy = scales[1]; a.sz = scales[2];
    a.ox = offsets[0]; a.oy = offsets[1]; a.oz = offsets[2];
    return PCH_OK;
}

extern "C" int pch_las_decode_f64(const uint8_t* rec, int64_t n, int32_t rec_len, const double* scales,
                                  const double* offsets, double* xyz, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rec_args(rec, n, rec_len);
    if (rc) return rc;
    PchAffine a;
    if ((rc = make_affine(scales, offsets, a))) return rc;
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz != nullptr, "null output");
    PchTileGeom g = pch_tile_geom(n, rec_len, n);
    PCH_DISPATCH_ALIGN(rec_len, k_decode_f64, g, 2, st, rec, a, xyz);
    return PCH_OK;
}

extern "C" int pch_las_decode_f32(const uint8_t* rec, int64_t n, int32_t rec_len, const double* scales,
                                  const double* offsets, float* xyz, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rec_args(rec, n, rec_len);
    if (rc) return rc;
    PchAffine a;
    if ((rc = make_affine(scales, offsets, a))) return rc;
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz != nullptr, "null output");
    PchTileGeom g = pch_tile_geom(n, rec_len, n);
    PCH_DISPATCH_ALIGN(rec_len, k_decode_f32, g, 2, st, rec, a, xyz);
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// quantise (las.x = arr) and encode (LasData.write)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pch_quantise1(double v, double scale, double offset) {
    return __double2int_rn(__ddiv_rn(__dsub_rn(v, offset), scale));
}

__global__ void k_quantise(const double* __restrict__ xyz, int64_t m, PchAffine a, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m * 3; i += stride) {
        int ax = (int)(i % 3);
        double s = ax == 0 ? a.sx : (ax == 1 ? a.sy : a.sz);
        double o = ax == 0 ? a.ox : (ax == 1 ? a.oy : a.oz);
        out[i] = pch_quantise1(xyz[i], s, o);
    }
}

extern "C" int pch_las_quantise(const double* xyz, int64_t m, const double* scales, const double* offsets,
                                int32_t* lattice, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    PchAffine a;
    int rc = make_affine(scales, offsets, a);
    if (rc) return rc;
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && lattice, "null pointer");
    int64_t blocks = pch_ceil_div(m * 3, 256);
    int64_t cap = (int64_t)pch_sm_count() * 16;
    if (blocks > cap) blocks = cap;
    PCH_LAUNCH(st, "k_quantise", k_quantise<<<(unsigned)blocks, 256, 0, st>>>(xyz, m, a, lattice));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

#define ENC_TILE 1024
// Builds ENC_TILE zeroed records in shared memory, drops X,Y,Z in, then streams the tile out with
// 16-byte stores (ENC_TILE*rec_len is a multiple of 16, and every tile starts 16-byte aligned).
__global__ void __launch_bounds__(256)
k_encode(const int32_t* __restrict__ lat, int64_t m, int32_t rec_len, uint8_t* __restrict__ out, int32_t* __restrict__ mm6) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ int s_part[8][6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int mnx = INT_MAX, mny = INT_MAX, mnz = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, mxz = INT_MIN;
    const int64_t n_tiles = (m + ENC_TILE - 1) / ENC_TILE;
    for (int64_t T = blockIdx.x; T < n_tiles; T += gridDim.x) {
        const int64_t r0 = T * ENC_TILE;
        const int cnt = (int)min((int64_t)ENC_TILE, m - r0);
        const int bytes = cnt * rec_len;
        const int words16 = (bytes + 15) / 16;
        uint4* sm4 = reinterpret_cast<uint4*>(sm);
        for (int i = tid; i < words16; i += 256) sm4[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        for (int r = tid; r < cnt; r += 256) {
            const int32_t* p = lat + (r0 + r) * 3;
            int X = p[0], Y = p[1], Z = p[2];
            mnx = min(mnx, X); mny = min(mny, Y); mnz = min(mnz, Z);
            mxx = max(mxx, X); mxy = max(mxy, Y); mxz = max(mxz, Z);
            uint8_t* q = sm + (size_t)r * rec_len;
            uint32_t v[3] = {(uint32_t)X, (uint32_t)Y, (uint32_t)Z};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                q[4 * k + 0] = (uint8_t)(v[k]);
                q[4 * k + 1] = (uint8_t)(v[k] >> 8);
                q[4 * k + 2] = (uint8_t)(v[k] >> 16);
                q[4 * k + 3] = (uint8_t)(v[k] >> 24);
            }
        }
        __syncthreads();
        uint8_t* dst = out + r0 * rec_len;  // 16-byte aligned: ENC_TILE*rec_len % 16 == 0
        const int full16 = bytes / 16;
        uint4* dst4 = reinterpret_cast<uint4*>(dst);
        for (int i = tid; i < full16; i += 256) dst4[i] = sm4[i];
        for (int i = full16 * 16 + tid; i < bytes; i += 256) dst[i] = sm[i];
        __syncthreads();
    }
    if (mm6) {
        mnx = pch_warp_min(mnx); mny = pch_warp_min(mny); mnz = pch_warp_min(mnz);
        mxx = pch_warp_max(mxx); mxy = pch_warp_max(mxy); mxz = pch_warp_max(mxz);
        if (lane == 0) {
            s_part[warp][0] = mnx; s_part[warp][1] = mny; s_part[warp][2] = mnz;
            s_part[warp][3] = mxx; s_part[warp][4] = mxy; s_part[warp][5] = mxz;
        }
        __syncthreads();
        if (tid < 6) {
            int v = s_part[0][tid];
            for (int w = 1; w < 8; ++w) v = (tid < 3) ? min(v, s_part[w][tid]) : max(v, s_part[w][tid]);
            if (tid < 3) atomicMin(&mm6[tid], v);
            else atomicMax(&mm6[tid], v);
        }
    }
}

extern "C" int pch_las_encode(const int32_t* lattice, int64_t m, int32_t rec_len, uint8_t* rec_out,
                              int32_t* mm6, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    PCH_CHECK_ARG(rec_len >= 12 && rec_len <= 128, "record length %d outside [12, 128]", rec_len);
    if (mm6) {
        PCH_LAUNCH(st, "k_init_minmax", k_init_minmax<<<1, 32, 0, st>>>(mm6, 1));
        PCH_LAUNCH_CHECK();
    }
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(lattice && rec_out, "null pointer");
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(rec_out) & 15) == 0, "output records must be 16-byte aligned");
    size_t smem = (size_t)ENC_TILE * rec_len + 16;
    PCH_CUDA(cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t tiles = pch_ceil_div(m, ENC_TILE);
    int64_t grid = (int64_t)pch_sm_count() * 2;
    if (grid > tiles) grid = tiles;
    PCH_LAUNCH(st, "k_encode", k_encode<<<(unsigned)grid, 256, smem, st>>>(lattice, m, rec_len, rec_out, mm6));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// per-tower box crop (test/kuangxuan.py:69-79) and preview subsample
// (pyGUI_towers_test.py:174-177; ui/vtk_widget.py:115-118)
//
// The reference loops over towers and builds one boolean mask over ALL points per tower.  Here one
// streaming pass over the raw records tests every point against every box (boxes in shared memory,
// union-box early out) and emits a word  box << 32 | point index  per hit; sorting the (small) word list
// restores `points[mask]` order for every tower at once.
// ------------------------------------------------------------------------------------------------
#define CROP_SMEM_BOXES 64   // 3 KB: keeps two CTAs of the 3-stage record ring resident per SM

struct CropArgs {
    const double* boxes;   // [n_boxes][6] xmin,ymin,zmin,xmax,ymax,zmax (inclusive, float64 compares)
    int32_t n_boxes;
    int32_t box0;          // id of boxes[0]
    uint64_t* words;
    long long capacity;
    unsigned long long* total;   // hits so far (may exceed capacity: the caller re-runs with more room)
};

template <int ALIGN>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 2)
k_box_crop(const uint8_t* __restrict__ rec, PchTileGeom g, PchAffine a, CropArgs c) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ double s_box[CROP_SMEM_BOXES][6];
    __shared__ double s_union[6];
    __shared__ uint32_t s_wsum[PCH_TILE_THREADS / 32];
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < c.n_boxes * 6; i += PCH_TILE_THREADS) s_box[i / 6][i % 6] = c.boxes[i];
    __syncthreads();
    if (tid < 6) {
        double v = s_box[0][tid];
        for (int b = 1; b < c.n_boxes; ++b) v = tid < 3 ? fmin(v, s_box[b][tid]) : fmax(v, s_box[b][tid]);
        s_union[tid] = v;
    }
    __syncthreads();
    constexpr int MAXR = 4;   // tile_records <= 1024 = 4 records per thread
    pch_stream_tiles(rec, g, smem, [&](const PchTile& t) {
        double px[MAXR], py[MAXR], pz[MAXR];
        uint32_t hits = 0;
        bool cand[MAXR];
#pragma unroll
        for (int j = 0; j < MAXR; ++j) {
            const int r = tid + j * PCH_TILE_THREADS;
            cand[j] = false;
            if (r < t.count) {
                int X, Y, Z;
                pch_load_xyz<ALIGN>(t.base + (size_t)r * g.rec_len, X, Y, Z);
                px[j] = pch_scaled(X, a.sx, a.ox); py[j] = pch_scaled(Y, a.sy, a.oy); pz[j] = pch_scaled(Z, a.sz, a.oz);
                cand[j] = px[j] >= s_union[0] && px[j] <= s_union[3] && py[j] >= s_union[1] && py[j] <= s_union[4] &&
                          pz[j] >= s_union[2] && pz[j] <= s_union[5];
            }
            if (cand[j])
                for (int b = 0; b < c.n_boxes; ++b)
                    hits += (px[j] >= s_box[b][0] && px[j] <= s_box[b][3] && py[j] >= s_box[b][1] && py[j] <= s_box[b][4] &&
                             pz[j] >= s_box[b][2] && pz[j] <= s_box[b][5]) ? 1u : 0u;
        }
        // block-exclusive scan of the per-thread hit counts, one global reservation per tile
        uint32_t incl = hits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wpre = 0, tile_hits = 0;
#pragma unroll
        for (int w = 0; w < PCH_TILE_THREADS / 32; ++w) {
            const uint32_t v = s_wsum[w];
            if (w < warp) wpre += v;
            tile_hits += v;
        }
        if (tile_hits == 0) return;     // uniform for the CTA: the streamer's trailing barrier still runs
        if (tid == 0) s_base = atomicAdd(c.total, (unsigned long long)tile_hits);
        __syncthreads();
        unsigned long long o = s_base + wpre + (incl - hits);
        if (hits == 0) return;
#pragma unroll
        for (int j = 0; j < MAXR; ++j) {
            if (!cand[j]) continue;
            const uint64_t idx = (uint64_t)(t.r0 + tid + j * PCH_TILE_THREADS);
            for (int b = 0; b < c.n_boxes; ++b)
                if (px[j] >= s_box[b][0] && px[j] <= s_box[b][3] && py[j] >= s_box[b][1] && py[j] <= s_box[b][4] &&
                    pz[j] >= s_box[b][2] && pz[j] <= s_box[b][5]) {
                    if ((long long)o < c.capacity) c.words[o] = ((uint64_t)(uint32_t)(c.box0 + b) << 32) | idx;
                    ++o;
                }
        }
    });
}

extern "C" int pch_las_box_crop(const uint8_t* rec, int64_t n, int32_t rec_len, const double* scales,
                                const double* offsets, const double* boxes_dev, int32_t n_boxes, uint64_t* words_dev,
                                int64_t capacity, int64_t* total_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rec_args(rec, n, rec_len);
    if (rc) return rc;
    PchAffine a;
    if ((rc = make_affine(scales, offsets, a))) return rc;
    PCH_CHECK_ARG(n < (1ll << 32), "more than 2^32-1 points per crop");
    PCH_CHECK_ARG(n_boxes >= 0 && capacity >= 0 && total_dev, "bad n_boxes/capacity/total");
    PCH_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), st));
    if (n == 0 || n_boxes == 0) return PCH_OK;
    PCH_CHECK_ARG(boxes_dev && (words_dev || capacity == 0), "null pointer");
    PchTileGeom g = pch_tile_geom(n, rec_len, n);
    for (int32_t b0 = 0; b0 < n_boxes; b0 += CROP_SMEM_BOXES) {
        CropArgs c;
        c.boxes = boxes_dev + (size_t)b0 * 6;
        c.n_boxes = n_boxes - b0 < CROP_SMEM_BOXES ? n_boxes - b0 : CROP_SMEM_BOXES;
        c.box0 = b0;
        c.words = words_dev;
        c.capacity = capacity;
        c.total = (unsigned long long*)total_dev;
        PCH_DISPATCH_ALIGN(rec_len, k_box_crop, g, 2, st, rec, a, c);
    }
    return PCH_OK;
}

// first word >= (b << 32) for b = 0..n_boxes in the sorted word list: tower b owns [bounds[b], bounds[b+1])
__global__ void k_word_bounds(const uint64_t* __restrict__ words, int64_t m, int32_t n_boxes, int64_t* __restrict__ bounds) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_boxes) return;
    const uint64_t key = (uint64_t)(uint32_t)b << 32;
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (words[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    bounds[b] = lo;
}

extern "C" int pch_word_bounds(const uint64_t* sorted_words_dev, int64_t m, int32_t n_boxes, int64_t* bounds_dev,
                               pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0 && n_boxes >= 0 && bounds_dev && (m == 0 || sorted_words_dev), "bad arguments");
    PCH_LAUNCH(st, "k_word_bounds", k_word_bounds<<<(unsigned)pch_ceil_div(n_boxes + 1, 128), 128, 0, st>>>(sorted_words_dev, m, n_boxes, bounds_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// points[index] decoded to float64 for a list of words (low 32 bits = point index)
template <int ALIGN>
__global__ void __launch_bounds__(256)
k_gather_decode_f64(const uint8_t* __restrict__ rec, int32_t rec_len, PchAffine a, const uint64_t* __restrict__ words,
                    int64_t m, double* __restrict__ out) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < m; j += stride) {
        const uint64_t i = words[j] & 0xffffffffull;
        int X, Y, Z;
        pch_load_xyz<ALIGN>(rec + (size_t)i * rec_len, X, Y, Z);
        out[j * 3 + 0] = pch_scaled(X, a.sx, a.ox);
        out[j * 3 + 1] = pch_scaled(Y, a.sy, a.oy);
        out[j * 3 + 2] = pch_scaled(Z, a.sz, a.oz);
    }
}

extern "C" int pch_las_gather_f64(const uint8_t* rec, int64_t n, int32_t rec_len, const double* scales,
                                  const double* offsets, const uint64_t* words_dev, int64_t m, double* out_dev,
                                  pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rec_args(rec, n, rec_len);
    if (rc) return rc;
    PchAffine a;
    if ((rc = make_affine(scales, offsets, a))) return rc;
    PCH_CHECK_ARG(m >= 0, "m must be >= 0");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(n > 0 && words_dev && out_dev, "null pointer / empty source");
    int64_t blocks = pch_ceil_div(m, 256);
    int64_t cap = (int64_t)pch_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const int al = pch_rec_align(rec_len);
    if (al == 4) { PCH_LAUNCH(st, "k_gather_decode_f64", k_gather_decode_f64<4><<<(unsigned)blocks, 256, 0, st>>>(rec, rec_len, a, words_dev, m, out_dev)); }
    else if (al == 2) { PCH_LAUNCH(st, "k_gather_decode_f64", k_gather_decode_f64<2><<<(unsigned)blocks, 256, 0, st>>>(rec, rec_len, a, words_dev, m, out_dev)); }
    else { PCH_LAUNCH(st, "k_gather_decode_f64", k_gather_decode_f64<1><<<(unsigned)blocks, 256, 0, st>>>(rec, rec_len, a, words_dev, m, out_dev)); }
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// Preview subsample: k distinct point indices out of n.  seed == 0: evenly spaced, idx_j = floor(j*n/k).
// seed != 0: idx_j = P(j) for a keyed bijection P of [0, n) (4-round balanced Feistel network on the smallest
// even number of bits covering n, cycle-walked back into range) = a sample without replacement.
__device__ __forceinline__ uint32_t pch_mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

__global__ void k_sample_indices(int64_t n, int64_t k, uint64_t seed, int half_bits, uint64_t* __restrict__ words) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= k) return;
    if (seed == 0) {
        words[j] = ((uint64_t)j * (uint64_t)n) / (uint64_t)k;   // j, n < 2^32: no overflow
        return;
    }
    const uint32_t mask = half_bits >= 32 ? 0xffffffffu : ((1u << half_bits) - 1u);
    uint64_t v = (uint64_t)j;
    do {
        uint32_t l = (uint32_t)(v >> half_bits) & mask, r = (uint32_t)v & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t key = (uint32_t)(seed >> (16 * (round & 1))) ^ (0x9e3779b9u * (uint32_t)(round + 1)) ^ (uint32_t)(seed >> 32);
            const uint32_t f = pch_mix32(r ^ key) & mask;
            const uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        v = ((uint64_t)l << half_bits) | (uint64_t)r;
    } while (v >= (uint64_t)n);
    words[j] = v;
}

extern "C" int pch_sample_indices(int64_t n, int64_t k, uint64_t seed, uint64_t* words_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && k >= 0 && k <= n && n < (1ll << 32), "need 0 <= k <= n < 2^32");
    if (k == 0) return PCH_OK;
    PCH_CHECK_ARG(words_dev, "null pointer");
    int bits = 2;
    while (bits < 64 && (1ull << bits) < (uint64_t)n) bits += 2;
    PCH_LAUNCH(st, "k_sample_indices", k_sample_indices<<<(unsigned)pch_ceil_div(k, 256), 256, 0, st>>>(n, k, seed, bits / 2, words_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
