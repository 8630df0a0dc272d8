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
