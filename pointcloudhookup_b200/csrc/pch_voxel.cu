// Voxel-grid downsample = open3d PointCloud.voxel_down_sample as called by
// ui/import_PC.py:8-13 (process_chunk) / ui/Sampling.py:10-18, on the raw LAS records of each
// `chunk_size` slice (ui/import_PC.py:45-50):
//   plan   : origin = chunk min bound - 0.5*voxel, index ranges -> packed key layout
//   keys   : idx = floor((p - origin)/voxel) in float64 (true divide), key = ix|iy|iz|index-in-chunk
//   (sort) : pch_sort_u64_segmented over the key bits, per chunk
//   reduce : per run of equal voxel, float64 running sum in input order, mean = sum/count,
//            optional re-quantisation (ui/import_PC.py:61-63) and float32 read-back
//            (utils/tower_extraction.py:60-62) fused into the same pass.
#include "pch_common.cuh"
#include "pch_tiles.cuh"
#include "pch_sort.cuh"

struct PchAffine3 {
    double s[3], o[3];
    double r[3];   // RN(1/s[i]) for pch_div_by
    int fast;      // every scale passed pch_recip_ok
};
struct VoxelDiv {
    double v, r;   // voxel size, RN(1/voxel)
    int fast;
};
static VoxelDiv make_voxel_div(double voxel) {
    VoxelDiv d;
    d.v = voxel;
    d.r = 1.0 / voxel;
    d.fast = pch_recip_ok(voxel) ? 1 : 0;
#ifdef PCH_NO_FASTDIV
    d.fast = 0;
#endif
    return d;
}
// floor((p - origin) / voxel) for the three axes: one range guard, then either three reciprocal sequences or
// three true divides (the divide code exists once per kernel)
__device__ __forceinline__ void pch_voxel_index3(double dx, double dy, double dz, const VoxelDiv& d, uint64_t& ix,
                                                 uint64_t& iy, uint64_t& iz) {
    double qx, qy, qz;
    if (d.fast && pch_div_inrange3(dx, dy, dz)) {
        qx = pch_div_by_nocheck(dx, d.v, d.r); qy = pch_div_by_nocheck(dy, d.v, d.r); qz = pch_div_by_nocheck(dz, d.v, d.r);
    } else {
        qx = __ddiv_rn(dx, d.v); qy = __ddiv_rn(dy, d.v); qz = __ddiv_rn(dz, d.v);
    }
    ix = (uint64_t)(long long)floor(qx); iy = (uint64_t)(long long)floor(qy); iz = (uint64_t)(long long)floor(qz);
}

__device__ __forceinline__ int bits_for(long long v) {  // bits needed to store values 0..v
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b;
}

__global__ void k_voxel_plan(const int32_t* __restrict__ mm, int64_t n_chunks, int64_t chunk_size, PchAffine3 a,
                             double voxel, double* __restrict__ origins, pch_voxel_plan* __restrict__ plan) {
    __shared__ long long s_max[3];
    if (threadIdx.x < 3) s_max[threadIdx.x] = 0;
    __syncthreads();
    const double half = __dmul_rn(voxel, 0.5);
    for (int64_t c = threadIdx.x; c < n_chunks; c += blockDim.x) {
        for (int ax = 0; ax < 3; ++ax) {
            double lo = pch_scaled(mm[c * 6 + ax], a.s[ax], a.o[ax]);
            double hi = pch_scaled(mm[c * 6 + 3 + ax], a.s[ax], a.o[ax]);
            if (hi < lo) { double t = lo; lo = hi; hi = t; }  // negative scale
            double org = __dsub_rn(lo, half);
            origins[c * 3 + ax] = org;
            double ref = __ddiv_rn(__dsub_rn(hi, org), voxel);
            long long im = (long long)floor(ref);
            atomicMax(&s_max[ax], im);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pch_voxel_plan p;
        p.bits_x = bits_for(s_max[0]);
        p.bits_y = bits_for(s_max[1]);
        p.bits_z = bits_for(s_max[2]);
        p.bits_idx = bits_for(chunk_size - 1);
        p.key_bits = p.bits_x + p.bits_y + p.bits_z;
        p.n_passes = (p.key_bits + 7) / 8;
        bool ok = (p.key_bits + p.bits_idx <= 64) && s_max[0] < 2147483647ll && s_max[1] < 2147483647ll &&
                  s_max[2] < 2147483647ll;
        p.status = ok ? PCH_OK : PCH_ERR_RANGE;
        p.reserved = 0;
        *plan = p;
    }
}

static int make_affine3(const double* scales, const double* offsets, PchAffine3& a) {
    a.fast = 1;
    if (!scales && !offsets) {  // float64 point input: identity
        for (int i = 0; i < 3; ++i) { a.s[i] = 1.0; a.o[i] = 0.0; a.r[i] = 1.0; }
        return PCH_OK;
    }
    PCH_CHECK_ARG(scales && offsets, "null scales/offsets");
    for (int i = 0; i < 3; ++i) {
        a.s[i] = scales[i];
        a.o[i] = offsets[i];
        PCH_CHECK_ARG(scales[i] != 0.0, "zero LAS scale");
        a.r[i] = 1.0 / scales[i];
        if (!pch_recip_ok(scales[i])) a.fast = 0;
    }
#ifdef PCH_NO_FASTDIV
    a.fast = 0;
#endif
    return PCH_OK;
}

extern "C" int pch_voxel_plan_build(const int32_t* mm, int64_t n_chunks, int64_t chunk_size, const double* scales,
                                    const double* offsets, double voxel, double* origins, pch_voxel_plan* plan,
                                    pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n_chunks >= 1 && chunk_size >= 1, "n_chunks and chunk_size must be >= 1");
    PCH_CHECK_ARG(voxel > 0.0, "voxel_size must be > 0 (open3d raises otherwise)");
    PCH_CHECK_ARG(mm && origins && plan, "null pointer");
    PchAffine3 a;
    int rc = make_affine3(scales, offsets, a);
    if (rc) return rc;
    PCH_LAUNCH(st, "k_voxel_plan", k_voxel_plan<<<1, 256, 0, st>>>(mm, n_chunks, chunk_size, a, voxel, origins, plan));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
struct KeyLayout {
    int sh_x, sh_y, sh_z;  // left shifts of ix, iy, iz
};

template <int ALIGN>
__global__ void __launch_bounds__(PCH_TILE_THREADS, 2)
k_voxel_keys(const uint8_t* __restrict__ rec, PchTileGeom g, PchAffine3 a, VoxelDiv voxel,
             const double* __restrict__ origins, KeyLayout kl, uint64_t* __restrict__ keys, int4* __restrict__ xyz16) {
    extern __shared__ __align__(128) uint8_t smem[];
    pch_stream_tiles(rec, g, smem, [&](const PchTile& t) {
        const double ox = origins[t.chunk * 3 + 0], oy = origins[t.chunk * 3 + 1], oz = origins[t.chunk * 3 + 2];
        const int64_t chunk_start = t.chunk * g.chunk_size;
        for (int r = threadIdx.x; r < t.count; r += PCH_TILE_THREADS) {
            int X, Y, Z;
            pch_load_xyz<ALIGN>(t.base + (size_t)r * g.rec_len, X, Y, Z);
            double x = pch_scaled(X, a.s[0], a.o[0]);
            double y = pch_scaled(Y, a.s[1], a.o[1]);
            double z = pch_scaled(Z, a.s[2], a.o[2]);
            // (p - origin) / voxel with a correctly rounded divide, then floor -> int (open3d)
            uint64_t ix, iy, iz;
            pch_voxel_index3(__dsub_rn(x, ox), __dsub_rn(y, oy), __dsub_rn(z, oz), voxel, ix, iy, iz);
            uint64_t local = (uint64_t)(t.r0 + r - chunk_start);
            keys[t.r0 + r] = (ix << kl.sh_x) | (iy << kl.sh_y) | (iz << kl.sh_z) | local;
            // 16-byte aligned copy of the lattice coordinates: the reduce pass gathers ONE aligned
            // vector per point instead of four loads from a 2-byte aligned 34-byte record
            if (xyz16) xyz16[t.r0 + r] = make_int4(X, Y, Z, 0);
        }
    });
}

extern "C" int pch_voxel_keys(const uint8_t* rec, int64_t n, int32_t rec_len, int64_t chunk_size, const double* scales,
                              const double* offsets, double voxel, const double* origins, const pch_voxel_plan* plan,
                              uint64_t* keys, int32_t* xyz16, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && rec_len >= 12 && rec_len <= 256, "bad n/rec_len");
    PCH_CHECK_ARG(chunk_size > 0 && voxel > 0.0, "chunk_size and voxel_size must be > 0");
    PCH_CHECK_ARG(plan != nullptr, "null plan");
    if (plan->status != PCH_OK || plan->key_bits + plan->bits_idx > 64) {
        pch_set_error("voxel index range needs %d+%d bits > 64: voxel_size too small for this chunk extent",
                      plan->key_bits, plan->bits_idx);
        return PCH_ERR_RANGE;
    }
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(rec && origins && keys, "null pointer");
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(rec) & 15) == 0, "record buffer must be 16-byte aligned");
    PchAffine3 a;
    int rc = make_affine3(scales, offsets, a);
    if (rc) return rc;
    PchTileGeom g = pch_tile_geom(n, rec_len, chunk_size);
    KeyLayout kl;
    kl.sh_z = plan->bits_idx;
    kl.sh_y = kl.sh_z + plan->bits_z;
    kl.sh_x = kl.sh_y + plan->bits_y;
    size_t smem = pch_tile_smem_bytes(g);
    int grid = pch_tile_grid(g, 2);
    int al = pch_rec_align(rec_len);
#define LAUNCH_KEYS(A)                                                                                       \
    do {                                                                                                     \
        PCH_CUDA(cudaFuncSetAttribute(k_voxel_keys<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        PCH_LAUNCH(st, "k_voxel_keys", k_voxel_keys<A><<<grid, PCH_TILE_THREADS, smem, st>>>(rec, g, a, make_voxel_div(voxel), origins, kl, keys, (int4*)xyz16));          \
    } while (0)
    if (al == 4) LAUNCH_KEYS(4);
    else if (al == 2) LAUNCH_KEYS(2);
    else LAUNCH_KEYS(1);
#undef LAUNCH_KEYS
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// keys from the packed lattice copy (plain coalesced 16-byte loads; no second pass over the records)
__global__ void __launch_bounds__(256)
k_voxel_keys16(const int4* __restrict__ xyz16, int64_t n, int64_t chunk, PchAffine3 a, VoxelDiv voxel,
               const double* __restrict__ origins, KeyLayout kl, uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int4 v = __ldg(xyz16 + i);
        const int64_t c = i / chunk;
        const double x = pch_scaled(v.x, a.s[0], a.o[0]);
        const double y = pch_scaled(v.y, a.s[1], a.o[1]);
        const double z = pch_scaled(v.z, a.s[2], a.o[2]);
        uint64_t ix, iy, iz;
        pch_voxel_index3(__dsub_rn(x, origins[c * 3 + 0]), __dsub_rn(y, origins[c * 3 + 1]),
                         __dsub_rn(z, origins[c * 3 + 2]), voxel, ix, iy, iz);
        keys[i] = (ix << kl.sh_x) | (iy << kl.sh_y) | (iz << kl.sh_z) | (uint64_t)(i - c * chunk);
    }
}

extern "C" int pch_voxel_keys_xyz16(const int32_t* xyz16, int64_t n, int64_t chunk_size, const double* scales,
                                    const double* offsets, double voxel, const double* origins,
                                    const pch_voxel_plan* plan, uint64_t* keys, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && chunk_size > 0 && voxel > 0.0 && plan, "bad arguments");
    if (plan->status != PCH_OK || plan->key_bits + plan->bits_idx > 64) {
        pch_set_error("voxel index range needs %d+%d bits > 64: voxel_size too small for this chunk extent",
                      plan->key_bits, plan->bits_idx);
        return PCH_ERR_RANGE;
    }
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz16 && origins && keys, "null pointer");
    PCH_CHECK_ARG((reinterpret_cast<uintptr_t>(xyz16) & 15) == 0, "xyz16 must be 16-byte aligned");
    PchAffine3 a;
    int rc = make_affine3(scales, offsets, a);
    if (rc) return rc;
    if (chunk_size > n) chunk_size = n;
    KeyLayout kl;
    kl.sh_z = plan->bits_idx;
    kl.sh_y = kl.sh_z + plan->bits_z;
    kl.sh_x = kl.sh_y + plan->bits_y;
    int64_t blocks = pch_ceil_div(n, 256);
    int64_t cap = (int64_t)pch_sm_count() * 16;
    PCH_LAUNCH(st, "k_voxel_keys16", k_voxel_keys16<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(
                                         (const int4*)xyz16, n, chunk_size, a, make_voxel_div(voxel), origins, kl, keys));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// Device-planned variant for the fused stage (pch_voxel_downsample_las): the key layout is read from the plan in
// DEVICE memory (no host round trip after the plan kernel), a CTA walks a contiguous range of RS_TILE-point
// tiles (a tile lies in one chunk, so the origin is loaded once per chunk instead of one 64-bit division per
// point), and the digit histograms of every radix pass are built on the way (pch_sort.cuh) — the sort then
// starts at its scan and never re-reads the keys for k_hist.
__global__ void __launch_bounds__(256)
k_voxel_keys16_plan(const int4* __restrict__ xyz16, SortGeom sg, PchAffine3 a, VoxelDiv voxel,
                    const double* __restrict__ origins, const pch_voxel_plan* __restrict__ dplan,
                    uint64_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * 256];
    const int tid = threadIdx.x;
    const pch_voxel_plan plan = *dplan;
    if (plan.status != PCH_OK || ((plan.key_bits + 7) >> 3) > sg.hist_passes) return;
    const int sh_z = plan.bits_idx, sh_y = sh_z + plan.bits_z, sh_x = sh_y + plan.bits_y;
    for (int i = tid; i < RS_MAX_PASSES * 256; i += 256) sh[i] = 0;
    __syncthreads();
    const int64_t tiles = sg.total_tickets;
    const int64_t per = (tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = t0 + per < tiles ? t0 + per : tiles;
    int64_t cur = -1;
    double ox = 0.0, oy = 0.0, oz = 0.0;
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t c = tile / sg.tiles_per_seg, lt = tile - c * sg.tiles_per_seg;
        const int64_t cstart = c * sg.seg_size;
        int64_t cend = cstart + sg.seg_size;
        if (cend > sg.n) cend = sg.n;
        const int64_t start = cstart + lt * RS_TILE;
        if (start >= cend) continue;
        const int cnt = (int)(cend - start > RS_TILE ? RS_TILE : cend - start);
        if (c != cur) {
            if (cur >= 0) {
                __syncthreads();
                pch_sort_hist_flush(sh, hist, cur, sg.hist_passes, plan.key_bits);
                __syncthreads();
            }
            cur = c;
            ox = origins[c * 3 + 0]; oy = origins[c * 3 + 1]; oz = origins[c * 3 + 2];
        }
        const uint64_t local0 = (uint64_t)(start - cstart);
        // four 16-byte loads per thread are issued before the first one is consumed (the profile of the plain loop
        // showed 56 % of the samples waiting on one load at a time)
        for (int i0 = tid; i0 < cnt; i0 += 256 * 4) {
            int4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 256;
                v[u] = i < cnt ? __ldg(xyz16 + start + i) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 256;
                if (i >= cnt) break;
                const double x = pch_scaled(v[u].x, a.s[0], a.o[0]);
                const double y = pch_scaled(v[u].y, a.s[1], a.o[1]);
                const double z = pch_scaled(v[u].z, a.s[2], a.o[2]);
                uint64_t ix, iy, iz;
                pch_voxel_index3(__dsub_rn(x, ox), __dsub_rn(y, oy), __dsub_rn(z, oz), voxel, ix, iy, iz);
                const uint64_t key = (ix << sh_x) | (iy << sh_y) | (iz << sh_z) | (local0 + (uint64_t)i);
                keys[start + i] = key;
                pch_sort_hist_add(sh, key, plan.bits_idx, plan.key_bits);
            }
        }
    }
    __syncthreads();
    if (cur >= 0) pch_sort_hist_flush(sh, hist, cur, sg.hist_passes, plan.key_bits);
}

// ------------------------------------------------------------------------------------------------
// segmented in-order reduction
// ------------------------------------------------------------------------------------------------
#ifndef VR_THREADS
#define VR_THREADS 256
#endif
#ifndef VR_ROWS
#define VR_ROWS 8
#endif
#define VR_TILE (VR_THREADS * VR_ROWS)
#define VR_WARPS (VR_THREADS / 32)
#define VR_RECIP 64
#ifndef VR_MINB
#define VR_MINB 4
#endif

struct ReduceGeom {
    int64_t n, chunk_size, tiles_per_chunk, total_tiles;
    int32_t bits_idx, rec_len;
};

extern "C" size_t pch_voxel_reduce_workspace_bytes(int64_t n, int64_t chunk_size) {
    if (n <= 0) return 256;
    if (chunk_size <= 0 || chunk_size > n) chunk_size = n;
    int64_t n_chunks = pch_ceil_div(n, chunk_size);
    int64_t tiles = n_chunks * pch_ceil_div(chunk_size, VR_TILE);
    return 256 + (size_t)tiles * 8;
}

// ALIGN > 0: `rec` is a raw LAS byte stream (record alignment ALIGN); ALIGN == 0: `rec` is an
// (n,3) float64 array (process_chunk's arbitrary point input).
//
// One voxel = one run of equal key in the sorted chunk.  Per tile of VR_TILE sorted slots:
//   1. every thread loads its slots' keys and gathers their points (all loads in flight at once);
//   2. head flags -> ballots -> the tile's head LIST (s_heads[r] = slot of the r-th run), tile count published;
//   3. the runs are folded VR_THREADS at a time, run r by thread r % VR_THREADS: every lane has a run (no
//      idle non-head lanes), the fold code exists once, and output rows r, r+1, ... are written by
//      consecutive lanes.  A run is summed IN INPUT ORDER in float64 (open3d's AccumulatedPoint), then
//      mean = sum/count and the LAS re-quantisation;
//   4. when only lattice / float32 outputs are wanted (the fused pipeline) the folds run BEFORE the
//      look-back wait and park their lattice triple in s_xyz[r] (slot r is never read again: round k only
//      reads slots >= k*VR_THREADS, and writes slots < (k+1)*VR_THREADS after a barrier).
template <int ALIGN>
__global__ void __launch_bounds__(VR_THREADS, VR_MINB)
k_voxel_reduce(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ keys_odd,
               const pch_voxel_plan* __restrict__ dplan, ReduceGeom g, const uint8_t* __restrict__ rec,
               const int4* __restrict__ xyz16, const int32_t* __restrict__ vidx /* (n,3) or NULL: wide keys */,
               PchAffine3 a,
               double* __restrict__ mean_out, int32_t* __restrict__ lat_out, float* __restrict__ f32_out,
               float* __restrict__ z32_out, unsigned long long* __restrict__ chunk_counts, long long* __restrict__ total_out,
               uint64_t* __restrict__ status, uint32_t* __restrict__ counter, int* __restrict__ err) {
    __shared__ uint64_t s_keys[VR_TILE + 1];  // [0] = key preceding the tile
    __shared__ int s_xyz[ALIGN > 0 ? VR_TILE : 1][3];  // lattice coordinates of the tile's points (LAS source)
    __shared__ uint16_t s_heads[VR_TILE + 2];
    __shared__ uint32_t s_wcount[VR_WARPS];
    __shared__ double s_recip[VR_RECIP + 1];   // RN(1/count) for the common small voxel populations
    __shared__ uint64_t s_tile_off;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bi = g.bits_idx;
    if (dplan) {
        // device-planned run: key layout and the buffer that holds the sorted keys (odd pass count -> the
        // ping-pong partner) come from the plan; an unusable plan is reported through total_out = -1
        const int st_plan = dplan->status, kb = dplan->key_bits;
        if (st_plan != PCH_OK) {
            if (blockIdx.x == 0 && tid == 0 && total_out) *total_out = -1;
            return;
        }
        bi = dplan->bits_idx;
        if (((kb + 7) >> 3) & 1) keys = keys_odd;
    }
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    if (tid <= VR_RECIP) s_recip[tid] = tid ? __ddiv_rn(1.0, (double)tid) : 0.0;
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= g.total_tiles) return;
    const int64_t chunk = tile / g.tiles_per_chunk;
    const int64_t lt = tile - chunk * g.tiles_per_chunk;
    const int64_t cstart = chunk * g.chunk_size;
    int64_t cend = cstart + g.chunk_size;
    if (cend > g.n) cend = g.n;
    const int64_t start = cstart + lt * VR_TILE;
    const int cnt = (int)min((int64_t)VR_TILE, cend - start);
    const uint64_t idx_mask = bi >= 64 ? ~0ull : ((1ull << bi) - 1ull);

    {
        // first all key loads, then all point gathers, so the whole tile's loads are in flight at once
        uint64_t kk[VR_ROWS];
#pragma unroll
        for (int j = 0; j < VR_ROWS; ++j) {
            const int i = tid + j * VR_THREADS;
            kk[j] = i < cnt ? keys[start + i] : 0ull;
        }
        int gx[VR_ROWS], gy[VR_ROWS], gz[VR_ROWS];
#pragma unroll
        for (int j = 0; j < VR_ROWS; ++j) {
            const int i = tid + j * VR_THREADS;
            gx[j] = gy[j] = gz[j] = 0;
            if (ALIGN > 0 && i < cnt) {
                const int64_t src = cstart + (int64_t)(kk[j] & idx_mask);
                if (xyz16) {
                    const int4 v = __ldg(xyz16 + src);
                    gx[j] = v.x; gy[j] = v.y; gz[j] = v.z;
                } else {
                    pch_load_xyz<(ALIGN > 0 ? ALIGN : 1)>(rec + (size_t)src * g.rec_len, gx[j], gy[j], gz[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < VR_ROWS; ++j) {
            const int i = tid + j * VR_THREADS;
            if (i < cnt) {
                s_keys[i + 1] = kk[j];
                if (ALIGN > 0) { s_xyz[i][0] = gx[j]; s_xyz[i][1] = gy[j]; s_xyz[i][2] = gz[j]; }
            }
        }
    }
    if (tid == 0) s_keys[0] = lt > 0 ? keys[start - 1] : ~0ull;
    __syncthreads();

    // head flags in warp-blocked order: warp w owns [w*32*VR_ROWS, ...), row j = 32 consecutive slots
    uint32_t row_rank[VR_ROWS];
    uint32_t head_bits = 0;
    uint32_t wtotal = 0;
    const int wbase = warp * (32 * VR_ROWS);
#pragma unroll
    for (int j = 0; j < VR_ROWS; ++j) {
        int i = wbase + j * 32 + lane;
        bool head = false;
        if (i < cnt) {
            uint64_t k = s_keys[i + 1] >> bi;
            if (vidx) {   // wide keys: the word only carries one axis; compare the full (ix,iy,iz) triples
                head = (i == 0 && lt == 0);
                if (!head) {
                    const int32_t* va = vidx + (cstart + (int64_t)(s_keys[i + 1] & idx_mask)) * 3;
                    const int32_t* vb = vidx + (cstart + (int64_t)(s_keys[i] & idx_mask)) * 3;
                    head = va[0] != vb[0] || va[1] != vb[1] || va[2] != vb[2];
                }
            } else {
                head = (i == 0 && lt == 0) || (k != (s_keys[i] >> bi));
            }
        }
        uint32_t b = __ballot_sync(0xffffffffu, head);
        row_rank[j] = wtotal + __popc(b & ((1u << lane) - 1u));
        wtotal += __popc(b);
        if (head) head_bits |= 1u << j;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();
    uint32_t wprefix = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < VR_WARPS; ++w) {
        uint32_t c = s_wcount[w];
        if (w < warp) wprefix += c;
        tile_total += c;
    }
#pragma unroll
    for (int j = 0; j < VR_ROWS; ++j)
        if (head_bits & (1u << j)) s_heads[wprefix + row_rank[j]] = (uint16_t)(wbase + j * 32 + lane);
    if (tid == 0) {
        s_heads[tile_total] = (uint16_t)cnt;
        pch_lookback_publish_u64(status, tile, 0, tile_total);   // publish early: successors never wait on our compute
        if (chunk_counts && tile_total) atomicAdd(&chunk_counts[chunk], (unsigned long long)tile_total);
    }
    const bool deferred = (ALIGN > 0) && (mean_out == nullptr);
    if (!deferred && tid == 0) {
        s_tile_off = pch_lookback_walk_u64(status, tile, 0, tile_total, err);
        if (tile == g.total_tiles - 1 && total_out) *total_out = (long long)(s_tile_off + tile_total);
    }
    __syncthreads();

    auto store_row = [&](uint64_t m, int qx, int qy, int qz) {
        if (lat_out) {
            lat_out[m * 3 + 0] = qx; lat_out[m * 3 + 1] = qy; lat_out[m * 3 + 2] = qz;
        }
        const float fz = (float)pch_scaled(qz, a.s[2], a.o[2]);
        if (f32_out) {
            f32_out[m * 3 + 0] = (float)pch_scaled(qx, a.s[0], a.o[0]);
            f32_out[m * 3 + 1] = (float)pch_scaled(qy, a.s[1], a.o[1]);
            f32_out[m * 3 + 2] = fz;
        }
        if (z32_out) z32_out[m] = fz;
    };

    const int rounds = ((int)tile_total + VR_THREADS - 1) / VR_THREADS;
#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
        const int r = rd * VR_THREADS + tid;
        const bool valid = r < (int)tile_total;
        double mx = 0.0, my = 0.0, mz = 0.0;
        int qx = 0, qy = 0, qz = 0;
        if (valid) {
            const int i0 = s_heads[r], i1 = s_heads[r + 1];
            double sx = 0.0, sy = 0.0, sz = 0.0;
            long long cntp = i1 - i0;
#pragma unroll 1
            for (int li = i0; li < i1; ++li) {
                if (ALIGN > 0) {
                    sx = __dadd_rn(sx, pch_scaled(s_xyz[ALIGN > 0 ? li : 0][0], a.s[0], a.o[0]));
                    sy = __dadd_rn(sy, pch_scaled(s_xyz[ALIGN > 0 ? li : 0][1], a.s[1], a.o[1]));
                    sz = __dadd_rn(sz, pch_scaled(s_xyz[ALIGN > 0 ? li : 0][2], a.s[2], a.o[2]));
                } else {
                    const double* q = reinterpret_cast<const double*>(rec) + (size_t)(cstart + (int64_t)(s_keys[li + 1] & idx_mask)) * 3;
                    sx = __dadd_rn(sx, q[0]); sy = __dadd_rn(sy, q[1]); sz = __dadd_rn(sz, q[2]);
                }
            }
            if (r == (int)tile_total - 1 && start + cnt < cend) {
                // the tile's last run may continue into the following tiles of the chunk
                const uint64_t hk = s_keys[i0 + 1];
                const int32_t* vhead = vidx ? vidx + (cstart + (int64_t)(hk & idx_mask)) * 3 : nullptr;
                for (int64_t p = start + cnt; p < cend; ++p) {
                    const uint64_t k = keys[p];
                    const int64_t src = cstart + (int64_t)(k & idx_mask);
                    if (vidx) {
                        const int32_t* vn = vidx + src * 3;
                        if (vn[0] != vhead[0] || vn[1] != vhead[1] || vn[2] != vhead[2]) break;
                    } else if ((k >> bi) != (hk >> bi)) break;
                    if (ALIGN > 0) {
                        int X, Y, Z;
                        if (xyz16) {
                            const int4 v = __ldg(xyz16 + src);
                            X = v.x; Y = v.y; Z = v.z;
                        } else {
                            pch_load_xyz<(ALIGN > 0 ? ALIGN : 1)>(rec + (size_t)src * g.rec_len, X, Y, Z);
                        }
                        sx = __dadd_rn(sx, pch_scaled(X, a.s[0], a.o[0]));
                        sy = __dadd_rn(sy, pch_scaled(Y, a.s[1], a.o[1]));
                        sz = __dadd_rn(sz, pch_scaled(Z, a.s[2], a.o[2]));
                    } else {
                        const double* q = reinterpret_cast<const double*>(rec) + (size_t)src * 3;
                        sx = __dadd_rn(sx, q[0]); sy = __dadd_rn(sy, q[1]); sz = __dadd_rn(sz, q[2]);
                    }
                    ++cntp;
                }
            }
            const double dn = (double)cntp;
            // same values as true divides (pch_div_by), at a fraction of the FP64 issue slots; one range guard
            // per stage, and a single copy of the true-divide code for everything the guards turn away
            bool fast = a.fast && cntp <= VR_RECIP && pch_div_inrange3(sx, sy, sz);
            if (fast) {
                const double yn = s_recip[cntp];
                mx = pch_div_by_nocheck(sx, dn, yn); my = pch_div_by_nocheck(sy, dn, yn); mz = pch_div_by_nocheck(sz, dn, yn);
                const double dx = __dsub_rn(mx, a.o[0]), dy = __dsub_rn(my, a.o[1]), dz = __dsub_rn(mz, a.o[2]);
                fast = pch_div_inrange3(dx, dy, dz);
                if (fast) {
                    qx = __double2int_rn(pch_div_by_nocheck(dx, a.s[0], a.r[0]));
                    qy = __double2int_rn(pch_div_by_nocheck(dy, a.s[1], a.r[1]));
                    qz = __double2int_rn(pch_div_by_nocheck(dz, a.s[2], a.r[2]));
                }
            }
            if (!fast) {
                mx = __ddiv_rn(sx, dn); my = __ddiv_rn(sy, dn); mz = __ddiv_rn(sz, dn);
                qx = __double2int_rn(__ddiv_rn(__dsub_rn(mx, a.o[0]), a.s[0]));
                qy = __double2int_rn(__ddiv_rn(__dsub_rn(my, a.o[1]), a.s[1]));
                qz = __double2int_rn(__ddiv_rn(__dsub_rn(mz, a.o[2]), a.s[2]));
            }
        }
        if (deferred) {
            __syncthreads();   // every read of this round (slots >= rd*VR_THREADS) is done before slots < (rd+1)*VR_THREADS change
            if (valid) { s_xyz[ALIGN > 0 ? r : 0][0] = qx; s_xyz[ALIGN > 0 ? r : 0][1] = qy; s_xyz[ALIGN > 0 ? r : 0][2] = qz; }
        } else if (valid) {
            const uint64_t m = s_tile_off + (uint64_t)r;
            if (mean_out) {
                mean_out[m * 3 + 0] = mx; mean_out[m * 3 + 1] = my; mean_out[m * 3 + 2] = mz;
            }
            if (ALIGN > 0) store_row(m, qx, qy, qz);
        }
    }
    if (!deferred) return;
    __syncthreads();
    if (tid == 0) {
        s_tile_off = pch_lookback_walk_u64(status, tile, 0, tile_total, err);
        if (tile == g.total_tiles - 1 && total_out) *total_out = (long long)(s_tile_off + tile_total);
    }
    __syncthreads();
    const uint64_t tile_off = s_tile_off;
    for (int r = tid; r < (int)tile_total; r += VR_THREADS)
        store_row(tile_off + (uint64_t)r, s_xyz[ALIGN > 0 ? r : 0][0], s_xyz[ALIGN > 0 ? r : 0][1], s_xyz[ALIGN > 0 ? r : 0][2]);
}

static int voxel_reduce_impl(const uint64_t* keys, const uint64_t* keys_odd, const pch_voxel_plan* dplan, int64_t n,
                             int64_t chunk_size, int32_t bits_idx,
                             const uint8_t* rec, int32_t rec_len, const int32_t* xyz16, const int32_t* vidx,
                             const double* scales, const double* offsets,
                             double* mean_out, int32_t* lat_out, float* f32_out, float* z32_out,
                             int64_t* chunk_counts, int64_t* total_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    PCH_CHECK_ARG(n >= 0 && chunk_size > 0, "bad n/chunk_size");
    PCH_CHECK_ARG(rec_len == 0 || (rec_len >= 12 && rec_len <= 256), "bad record length");
    PCH_CHECK_ARG(bits_idx >= 0 && bits_idx <= 63, "bad bits_idx");
    PCH_CHECK_ARG(rec_len != 0 || (lat_out == nullptr && f32_out == nullptr && z32_out == nullptr),
                  "float64 point input has no LAS lattice: only mean_dev is available");
    PchAffine3 a;
    int rc = make_affine3(scales, offsets, a);
    if (rc) return rc;
    if (chunk_size > n) chunk_size = n > 0 ? n : 1;
    int64_t n_chunks = pch_ceil_div(n > 0 ? n : 1, chunk_size);
    if (chunk_counts) PCH_CUDA(cudaMemsetAsync(chunk_counts, 0, sizeof(int64_t) * n_chunks, st));
    if (total_out) PCH_CUDA(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(keys && (rec || (xyz16 && rec_len != 0)) && workspace, "null pointer");
    PCH_CHECK_ARG(!xyz16 || (reinterpret_cast<uintptr_t>(xyz16) & 15) == 0, "xyz16 must be 16-byte aligned");
    ReduceGeom g;
    g.n = n; g.chunk_size = chunk_size; g.bits_idx = bits_idx; g.rec_len = rec_len;
    g.tiles_per_chunk = pch_ceil_div(chunk_size, VR_TILE);
    int64_t last = n - (n_chunks - 1) * chunk_size;
    g.total_tiles = (n_chunks - 1) * g.tiles_per_chunk + pch_ceil_div(last, VR_TILE);
    size_t need = 256 + (size_t)g.total_tiles * 8;
    if (workspace_bytes < need) {
        pch_set_error("voxel_reduce workspace too small: %zu < %zu", workspace_bytes, need);
        return PCH_ERR_WORKSPACE;
    }
    PCH_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    int* err = (int*)workspace;
    uint32_t* counter = (uint32_t*)((uint8_t*)workspace + 64);
    uint64_t* status = (uint64_t*)((uint8_t*)workspace + 256);
    int al = rec_len == 0 ? 0 : pch_rec_align(rec_len);
#define LAUNCH_RED(A)                                                                                          \
    PCH_LAUNCH(st, "k_voxel_reduce", k_voxel_reduce<A><<<(unsigned)g.total_tiles, VR_THREADS, 0, st>>>(                                         \
        keys, keys_odd, dplan, g, rec, (const int4*)xyz16, vidx, a, mean_out, lat_out, f32_out, z32_out, (unsigned long long*)chunk_counts, (long long*)total_out, \
        status, counter, err))
    if (al == 4) LAUNCH_RED(4);
    else if (al == 2) LAUNCH_RED(2);
    else if (al == 1) LAUNCH_RED(1);
    else LAUNCH_RED(0);
#undef LAUNCH_RED
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_voxel_reduce(const uint64_t* keys, int64_t n, int64_t chunk_size, int32_t bits_idx,
                                const uint8_t* rec, int32_t rec_len, const int32_t* xyz16, const int32_t* vidx,
                                const double* scales, const double* offsets,
                                double* mean_out, int32_t* lat_out, float* f32_out, float* z32_out,
                                int64_t* chunk_counts, int64_t* total_out, void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    return voxel_reduce_impl(keys, nullptr, nullptr, n, chunk_size, bits_idx, rec, rec_len, xyz16, vidx, scales, offsets,
                             mean_out, lat_out, f32_out, z32_out, chunk_counts, total_out, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// The whole voxel stage of the LAS path as ONE host call with no host round trip inside
// (ui/import_PC.py:45-63 for all chunks at once): chunk extrema + the 16-byte lattice copy -> plan (stays in
// device memory) -> keys + radix histograms -> scan + the radix passes the plan needs (launched for the
// widest key the chunk size allows; surplus passes exit at once) -> in-order reduce, which reads the plan to
// know which ping-pong buffer holds the sorted keys.  The caller reads back the 64-byte head of the
// workspace once: [0] error word of the look-backs, [8] M (or -1: the index range does not fit one key word,
// take the wide-key path), [32..64) the plan.
// ------------------------------------------------------------------------------------------------
extern "C" int pch_las_chunk_minmax(const uint8_t* rec, int64_t n, int32_t rec_len, int64_t chunk_size, int32_t* mm,
                                    int32_t* xyz16, pch_stream_t stream);

struct VoxelFusedWs {
    size_t head, sort, sort_bytes, reduce, reduce_bytes, total;
    int passes;
};
static VoxelFusedWs voxel_fused_ws(int64_t n, int64_t chunk_size) {
    VoxelFusedWs w;
    if (chunk_size <= 0 || chunk_size > n) chunk_size = n > 0 ? n : 1;
    int bits_idx = 0;
    for (int64_t v = chunk_size - 1; v > 0; v >>= 1) ++bits_idx;
    w.passes = (64 - bits_idx + 7) / 8;
    if (w.passes > RS_MAX_PASSES) w.passes = RS_MAX_PASSES;
    if (w.passes < 1) w.passes = 1;
    SortGeom sg = pch_sort_geom(n, chunk_size, 0, 0, w.passes);
    w.head = 0;
    w.sort = 256;
    w.sort_bytes = pch_sort_ws(sg, nullptr).bytes;
    w.reduce = w.sort + pch_align_up(w.sort_bytes, 256);
    w.reduce_bytes = pch_voxel_reduce_workspace_bytes(n, chunk_size);
    w.total = w.reduce + pch_align_up(w.reduce_bytes, 256);
    return w;
}

extern "C" size_t pch_voxel_downsample_las_workspace_bytes(int64_t n, int64_t chunk_size) {
    if (n <= 0) return 256;
    return voxel_fused_ws(n, chunk_size).total;
}

__global__ void k_voxel_head(const pch_voxel_plan* __restrict__ plan, const int* __restrict__ sort_err,
                             const int* __restrict__ red_err, uint8_t* __restrict__ head) {
    if (threadIdx.x == 0) {
        *reinterpret_cast<int*>(head) = (*sort_err) | (*red_err);
        *reinterpret_cast<pch_voxel_plan*>(head + 32) = *plan;
    }
}

extern "C" int pch_voxel_downsample_las(const uint8_t* rec, int64_t n, int32_t rec_len, int64_t chunk_size,
                                        const double* scales, const double* offsets, double voxel,
                                        int32_t* xyz16, uint64_t* keys, uint64_t* tmp, int32_t* minmax, double* origins,
                                        pch_voxel_plan* plan_dev,
                                        double* mean_out, int32_t* lat_out, float* f32_out, float* z32_out,
                                        int64_t* chunk_counts, void* workspace, size_t workspace_bytes,
                                        pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 1 && rec_len >= 12 && rec_len <= 256, "bad n/rec_len");
    PCH_CHECK_ARG(chunk_size > 0 && voxel > 0.0, "chunk_size and voxel_size must be > 0");
    PCH_CHECK_ARG(rec && xyz16 && keys && tmp && minmax && origins && plan_dev && workspace, "null pointer");
    if (chunk_size > n) chunk_size = n;
    const int64_t n_chunks = pch_ceil_div(n, chunk_size);
    VoxelFusedWs w = voxel_fused_ws(n, chunk_size);
    if (workspace_bytes < w.total) {
        pch_set_error("voxel_downsample_las workspace too small: %zu < %zu", workspace_bytes, w.total);
        return PCH_ERR_WORKSPACE;
    }
    uint8_t* base = (uint8_t*)workspace;
    PCH_CUDA(cudaMemsetAsync(base, 0, 256, st));
    int rc = pch_las_chunk_minmax(rec, n, rec_len, chunk_size, minmax, xyz16, stream);
    if (rc) return rc;
    rc = pch_voxel_plan_build(minmax, n_chunks, chunk_size, scales, offsets, voxel, origins, plan_dev, stream);
    if (rc) return rc;
    PchAffine3 a;
    if ((rc = make_affine3(scales, offsets, a))) return rc;
    SortGeom sg = pch_sort_geom(n, chunk_size, 0, 0, w.passes);
    if ((rc = pch_sort_prepare(sg, base + w.sort, w.sort_bytes, st))) return rc;
    SortWs sw = pch_sort_ws(sg, base + w.sort);
    {
        int64_t grid = (int64_t)pch_sm_count() * 8;
        if (grid > sg.total_tickets) grid = sg.total_tickets;
        PCH_LAUNCH(st, "k_voxel_keys16", k_voxel_keys16_plan<<<(unsigned)grid, 256, 0, st>>>(
                                             (const int4*)xyz16, sg, a, make_voxel_div(voxel), origins, plan_dev, keys, sw.hist));
        PCH_LAUNCH_CHECK();
    }
    if ((rc = pch_sort_run(keys, tmp, sg, plan_dev, w.passes, true, base + w.sort, w.sort_bytes, st))) return rc;
    int64_t* total = reinterpret_cast<int64_t*>(base + 8);
    rc = voxel_reduce_impl(keys, tmp, plan_dev, n, chunk_size, 0, rec, rec_len, xyz16, nullptr, scales, offsets, mean_out, lat_out,
                           f32_out, z32_out, chunk_counts, total, base + w.reduce, w.reduce_bytes, st);
    if (rc) return rc;
    PCH_LAUNCH(st, "k_voxel_head", k_voxel_head<<<1, 32, 0, st>>>(plan_dev, (const int*)(base + w.sort), (const int*)(base + w.reduce), base));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


// ------------------------------------------------------------------------------------------------
// float64 point arrays: process_chunk(points_chunk, voxel_size) (ui/import_PC.py:8-13,
// ui/Sampling.py:10-18) accepts ANY (n,3) array, not only LAS-lattice points.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_to_ordered(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_f64(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

__global__ void k_f64_minmax_init(unsigned long long* mm, int64_t n_chunks) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_chunks * 6) mm[i] = ((i % 6) < 3) ? ~0ull : 0ull;
}

__global__ void __launch_bounds__(256)
k_f64_chunk_minmax(const double* __restrict__ xyz, int64_t n, int64_t chunk, unsigned long long* __restrict__ mm) {
    const int64_t slabs = (chunk + 4095) / 4096;
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    for (int64_t s = blockIdx.x; s < n_chunks * slabs; s += gridDim.x) {
        int64_t c = s / slabs, ls = s - c * slabs;
        int64_t lo = c * chunk + ls * 4096, hi = min(min(lo + 4096, (c + 1) * chunk), n);
        unsigned long long mn[3] = {~0ull, ~0ull, ~0ull}, mx[3] = {0, 0, 0};
        for (int64_t i = lo + threadIdx.x; i < hi; i += 256)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                unsigned long long u = f64_to_ordered(xyz[i * 3 + a]);
                mn[a] = min(mn[a], u);
                mx[a] = max(mx[a], u);
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mn[a] = min(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
                mx[a] = max(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            }
            if ((threadIdx.x & 31) == 0 && lo < hi) {
                atomicMin(&mm[c * 6 + a], mn[a]);
                atomicMax(&mm[c * 6 + 3 + a], mx[a]);
            }
        }
    }
}

__global__ void k_voxel_plan_f64(const unsigned long long* __restrict__ mm, int64_t n_chunks, int64_t chunk_size,
                                 double voxel, double* __restrict__ origins, pch_voxel_plan* __restrict__ plan) {
    __shared__ long long s_max[3];
    if (threadIdx.x < 3) s_max[threadIdx.x] = 0;
    __syncthreads();
    const double half = __dmul_rn(voxel, 0.5);
    for (int64_t c = threadIdx.x; c < n_chunks; c += blockDim.x)
        for (int ax = 0; ax < 3; ++ax) {
            double lo = ordered_to_f64(mm[c * 6 + ax]), hi = ordered_to_f64(mm[c * 6 + 3 + ax]);
            double org = __dsub_rn(lo, half);
            origins[c * 3 + ax] = org;
            atomicMax(&s_max[ax], (long long)floor(__ddiv_rn(__dsub_rn(hi, org), voxel)));
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        pch_voxel_plan p;
        p.bits_x = bits_for(s_max[0]); p.bits_y = bits_for(s_max[1]); p.bits_z = bits_for(s_max[2]);
        p.bits_idx = bits_for(chunk_size - 1);
        p.key_bits = p.bits_x + p.bits_y + p.bits_z;
        p.n_passes = (p.key_bits + 7) / 8;
        bool ok = (p.key_bits + p.bits_idx <= 64) && s_max[0] < 2147483647ll && s_max[1] < 2147483647ll &&
                  s_max[2] < 2147483647ll;
        p.status = ok ? PCH_OK : PCH_ERR_RANGE;
        p.reserved = 0;
        *plan = p;
    }
}

__global__ void k_voxel_keys_f64(const double* __restrict__ xyz, int64_t n, int64_t chunk, double voxel,
                                 const double* __restrict__ origins, KeyLayout kl, uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t c = i / chunk;
        uint64_t ix = (uint64_t)(long long)floor(__ddiv_rn(__dsub_rn(xyz[i * 3 + 0], origins[c * 3 + 0]), voxel));
        uint64_t iy = (uint64_t)(long long)floor(__ddiv_rn(__dsub_rn(xyz[i * 3 + 1], origins[c * 3 + 1]), voxel));
        uint64_t iz = (uint64_t)(long long)floor(__ddiv_rn(__dsub_rn(xyz[i * 3 + 2], origins[c * 3 + 2]), voxel));
        keys[i] = (ix << kl.sh_x) | (iy << kl.sh_y) | (iz << kl.sh_z) | (uint64_t)(i - c * chunk);
    }
}

extern "C" int pch_voxel_plan_build_f64(const double* xyz, int64_t n, int64_t chunk_size, double voxel,
                                        uint64_t* minmax_scratch, double* origins, pch_voxel_plan* plan,
                                        pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 1 && chunk_size >= 1, "n and chunk_size must be >= 1");
    PCH_CHECK_ARG(voxel > 0.0, "voxel_size must be > 0 (open3d raises otherwise)");
    PCH_CHECK_ARG(xyz && minmax_scratch && origins && plan, "null pointer");
    if (chunk_size > n) chunk_size = n;
    int64_t n_chunks = pch_ceil_div(n, chunk_size);
    PCH_LAUNCH(st, "k_f64_minmax_init", k_f64_minmax_init<<<(unsigned)pch_ceil_div(n_chunks * 6, 256), 256, 0, st>>>((unsigned long long*)minmax_scratch, n_chunks));
    int64_t blocks = pch_ceil_div(n, 4096);
    int64_t cap = (int64_t)pch_sm_count() * 16;
    PCH_LAUNCH(st, "k_f64_chunk_minmax", k_f64_chunk_minmax<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(xyz, n, chunk_size, (unsigned long long*)minmax_scratch));
    PCH_LAUNCH(st, "k_voxel_plan_f64", k_voxel_plan_f64<<<1, 256, 0, st>>>((const unsigned long long*)minmax_scratch, n_chunks, chunk_size, voxel, origins, plan));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_voxel_keys_f64(const double* xyz, int64_t n, int64_t chunk_size, double voxel, const double* origins,
                                  const pch_voxel_plan* plan, uint64_t* keys, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && chunk_size > 0 && voxel > 0.0 && plan, "bad arguments");
    if (plan->status != PCH_OK || plan->key_bits + plan->bits_idx > 64) {
        pch_set_error("voxel index range needs %d+%d bits > 64: voxel_size too small", plan->key_bits, plan->bits_idx);
        return PCH_ERR_RANGE;
    }
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && origins && keys, "null pointer");
    if (chunk_size > n) chunk_size = n;
    KeyLayout kl;
    kl.sh_z = plan->bits_idx;
    kl.sh_y = kl.sh_z + plan->bits_z;
    kl.sh_x = kl.sh_y + plan->bits_y;
    int64_t blocks = pch_ceil_div(n, 256);
    int64_t cap = (int64_t)pch_sm_count() * 8;
    PCH_LAUNCH(st, "k_voxel_keys_f64", k_voxel_keys_f64<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(xyz, n, chunk_size, voxel, origins, kl, keys));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


// ------------------------------------------------------------------------------------------------
// wide keys: when bits(ix)+bits(iy)+bits(iz)+bits(index) > 64 the voxel index does not fit one sort
// word.  The triples are stored once; the chunk is then sorted by three stable rounds (z, y, x), each
// round sorting words  axis_index << bits_idx | index  built in the order of the previous round (LSD
// over axes), and the reduce pass compares triples instead of word prefixes.
// ------------------------------------------------------------------------------------------------
__global__ void k_voxel_index3(const double* __restrict__ xyz, int64_t n, int64_t chunk, double voxel,
                               const double* __restrict__ origins, int32_t* __restrict__ vidx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t c = i / chunk;
#pragma unroll
        for (int a = 0; a < 3; ++a)
            vidx[i * 3 + a] = (int32_t)(long long)floor(__ddiv_rn(__dsub_rn(xyz[i * 3 + a], origins[c * 3 + a]), voxel));
    }
}

__global__ void k_wide_words(const uint64_t* __restrict__ prev /* NULL = identity order */, const int32_t* __restrict__ vidx,
                             int64_t n, int64_t chunk, int bits_idx, int axis, uint64_t* __restrict__ out) {
    const uint64_t idx_mask = (1ull << bits_idx) - 1ull;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t c = i / chunk;
        const uint64_t idx = prev ? (prev[i] & idx_mask) : (uint64_t)(i - c * chunk);
        out[i] = ((uint64_t)(uint32_t)vidx[(c * chunk + (int64_t)idx) * 3 + axis] << bits_idx) | idx;
    }
}

extern "C" int pch_voxel_index3_f64(const double* xyz, int64_t n, int64_t chunk_size, double voxel, const double* origins,
                                    int32_t* vidx, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && chunk_size > 0 && voxel > 0.0, "bad arguments");
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(xyz && origins && vidx, "null pointer");
    if (chunk_size > n) chunk_size = n;
    int64_t blocks = pch_ceil_div(n, 256), cap = (int64_t)pch_sm_count() * 8;
    PCH_LAUNCH(st, "k_voxel_index3", k_voxel_index3<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(xyz, n, chunk_size, voxel, origins, vidx));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_voxel_wide_words(const uint64_t* prev, const int32_t* vidx, int64_t n, int64_t chunk_size, int32_t bits_idx,
                                    int32_t axis, uint64_t* out, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && chunk_size > 0 && bits_idx >= 0 && bits_idx <= 32 && axis >= 0 && axis < 3, "bad arguments");
    if (n == 0) return PCH_OK;
    PCH_CHECK_ARG(vidx && out, "null pointer");
    if (chunk_size > n) chunk_size = n;
    int64_t blocks = pch_ceil_div(n, 256), cap = (int64_t)pch_sm_count() * 8;
    PCH_LAUNCH(st, "k_wide_words", k_wide_words<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(prev, vidx, n, chunk_size, bits_idx, axis, out));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// ------------------------------------------------------------------------------------------------
// self-test of pch_div_by against the hardware-sequence IEEE divide (used by the parity tests)
// ------------------------------------------------------------------------------------------------
__global__ void k_selftest_fastdiv(const double* __restrict__ a, int64_t n, double b, double y,
                                   unsigned long long* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long mine = 0;
    for (; i < n; i += stride)
        mine += __double_as_longlong(pch_div_by(a[i], b, y)) != __double_as_longlong(__ddiv_rn(a[i], b));
    if (mine) atomicAdd(bad, mine);
}

extern "C" int pch_selftest_fastdiv(const double* a_dev, int64_t n, double b, int64_t* mismatches_dev,
                                    pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && mismatches_dev && (n == 0 || a_dev), "bad arguments");
    PCH_CHECK_ARG(pch_recip_ok(b), "divisor %g is not eligible for the reciprocal path", b);
    PCH_CUDA(cudaMemsetAsync(mismatches_dev, 0, sizeof(int64_t), st));
    if (n == 0) return PCH_OK;
    int64_t blocks = pch_ceil_div(n, 256);
    int64_t cap = (int64_t)pch_sm_count() * 8;
    PCH_LAUNCH(st, "k_selftest_fastdiv", k_selftest_fastdiv<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(
                                              a_dev, n, b, 1.0 / b, (unsigned long long*)mismatches_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
