// Chunked DBSCAN with scikit-learn's exact semantics (utils/tower_extraction.py:96-122):
//   the filtered points are cut into consecutive 50 000-point chunks; every chunk is clustered
//   independently with DBSCAN(eps, min_samples, algorithm='ball_tree'); labels are offset per chunk.
// sklearn rules restated (probed against scikit-learn 1.9.0):
//   * float32 input promoted to float64; neighbour iff ((dx*dx + dy*dy) + dz*dz) <= eps*eps
//   * core iff #neighbours INCLUDING the point itself >= min_samples
//   * cluster ids in order of each cluster's smallest core index; a border point takes the
//     smallest id among clusters owning a core point within eps; everything else is -1
//   * per chunk offset: labels += clusters found in all earlier chunks.
// All chunks are processed in ONE batch of launches (a single chunk cannot fill a B200).
//
// Acceleration structure (exact, never changes the predicate): per chunk a grid of cell side
// l = eps/sqrt(3)*(1-1e-7), so any two points of one cell are neighbours (a cell with >= min_samples
// points is all-core and all core points of a cell share a cluster), and neighbours of a point lie
// within +-2 cells per axis.  Clusters are a union-find over CELLS that hold a core point.
#include "pch_common.cuh"
#include "pch_sort.cuh"

struct DbPlan {
    int32_t bits_x, bits_y, bits_z, bits_idx, key_bits, n_passes, status, reserved;
};

struct DbGeom {
    int64_t G, chunk, n_chunks;
    double cell, eps2;
    int32_t min_pts;
    int32_t sh_x, sh_y, sh_z, bits_idx;  // key layout
    int32_t bits_x, bits_y, bits_z, key_bits;
    const DbPlan* dplan;                 // non-NULL: the key layout is read from this plan in DEVICE memory
    int64_t own_lo, own_hi;              // original indices that count in the per-cluster statistics (halo points do not)
};

// Device-planned runs (pch_dbscan): every kernel that needs the key layout takes it from the plan the plan
// kernel left in device memory, so the host never waits for it.  false = the plan is unusable (cell grid too
// wide for one key word); the kernels then do nothing and the host learns it from the scalar block.
__device__ __forceinline__ bool db_resolve(DbGeom& g) {
    if (g.dplan) {
        const DbPlan p = *g.dplan;
        if (p.status != PCH_OK) return false;
        g.bits_idx = p.bits_idx;
        g.bits_x = p.bits_x; g.bits_y = p.bits_y; g.bits_z = p.bits_z;
        g.key_bits = p.key_bits;
        g.sh_z = p.bits_idx; g.sh_y = g.sh_z + p.bits_z; g.sh_x = g.sh_y + p.bits_y;
    }
    return true;
}

static unsigned db_grid(int64_t n, int threads, int per_sm = 8) {
    int64_t b = pch_ceil_div(n > 0 ? n : 1, threads);
    int64_t cap = (int64_t)pch_sm_count() * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

// ---------------------------------------------------------------- D1: chunk bounds + plan
__global__ void k_db_bounds_init(uint32_t* mm, int64_t n_chunks) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_chunks * 6) mm[i] = ((i % 6) < 3) ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) k_db_bounds(const float* __restrict__ P, int64_t G, int64_t chunk, uint32_t* __restrict__ mm) {
    // one block handles a 4096-point slab inside one chunk; a thread takes runs of 4 consecutive points
    // (48 bytes = three 16-byte loads when the run starts on a 16-byte boundary)
    const int64_t slabs_per_chunk = (chunk + 4095) / 4096;
    const int64_t n_chunks = (G + chunk - 1) / chunk;
    const bool base16 = (reinterpret_cast<uintptr_t>(P) & 15) == 0;
    for (int64_t s = blockIdx.x; s < n_chunks * slabs_per_chunk; s += gridDim.x) {
        int64_t c = s / slabs_per_chunk, ls = s - c * slabs_per_chunk;
        int64_t lo = c * chunk + ls * 4096, hi = min(min(lo + 4096, (c + 1) * chunk), G);
        uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0, 0, 0};
        for (int64_t i0 = lo + (int64_t)threadIdx.x * 4; i0 < hi; i0 += 256 * 4) {
            if (base16 && (i0 & 3) == 0 && i0 + 4 <= hi) {
                const float4* src = reinterpret_cast<const float4*>(P + i0 * 3);
                float v[12];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 f = __ldg(src + k);
                    v[4 * k + 0] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
                }
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const uint32_t u = pch_f32_to_ordered(v[k]);
                    mn[k % 3] = min(mn[k % 3], u);
                    mx[k % 3] = max(mx[k % 3], u);
                }
            } else {
                for (int64_t i = i0; i < min(i0 + 4, hi); ++i)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const uint32_t u = pch_f32_to_ordered(P[i * 3 + a]);
                        mn[a] = min(mn[a], u);
                        mx[a] = max(mx[a], u);
                    }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            uint32_t lo_ = __reduce_min_sync(0xffffffffu, mn[a]);
            uint32_t hi_ = __reduce_max_sync(0xffffffffu, mx[a]);
            if ((threadIdx.x & 31) == 0 && lo < hi) {
                atomicMin(&mm[c * 6 + a], lo_);
                atomicMax(&mm[c * 6 + 3 + a], hi_);
            }
        }
    }
}

__device__ __forceinline__ int db_bits_for(long long v) {
    int b = 0;
    while (v > 0) { ++b; v >>= 1; }
    return b;
}

__device__ __forceinline__ long long db_cell_of(float v, float mn, double cell) {
    return (long long)floor(__ddiv_rn(__dsub_rn((double)v, (double)mn), cell));
}

__global__ void k_db_plan(const uint32_t* __restrict__ mm, int64_t n_chunks, int64_t chunk, double cell, DbPlan* plan) {
    __shared__ long long s_max[3];
    if (threadIdx.x < 3) s_max[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n_chunks; c += blockDim.x)
        for (int a = 0; a < 3; ++a) {
            float lo = pch_ordered_to_f32(mm[c * 6 + a]), hi = pch_ordered_to_f32(mm[c * 6 + 3 + a]);
            atomicMax(&s_max[a], db_cell_of(hi, lo, cell));
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        DbPlan p;
        // +2 guard cells on the low side are not needed (indices are >= 0); neighbours beyond the
        // max index simply do not exist
        p.bits_x = db_bits_for(s_max[0]); p.bits_y = db_bits_for(s_max[1]); p.bits_z = db_bits_for(s_max[2]);
        p.bits_idx = db_bits_for(chunk - 1);
        p.key_bits = p.bits_x + p.bits_y + p.bits_z;
        p.n_passes = (p.key_bits + 7) / 8;
        p.status = (p.key_bits + p.bits_idx <= 62) ? PCH_OK : PCH_ERR_RANGE;
        p.reserved = 0;
        *plan = p;
    }
}

// ---------------------------------------------------------------- D2: cell keys
// floor((v - mn) / cell) for the three axes with one range guard: the reciprocal sequence returns exactly what
// the true divide returns (pch_div_by), so the cell grid is the same either way
__device__ __forceinline__ void db_cell3(float x, float y, float z, float mnx, float mny, float mnz, double cell,
                                         double rcell, int fast, uint64_t& cx, uint64_t& cy, uint64_t& cz) {
    const double dx = __dsub_rn((double)x, (double)mnx), dy = __dsub_rn((double)y, (double)mny),
                 dz = __dsub_rn((double)z, (double)mnz);
    double qx, qy, qz;
    if (fast && pch_div_inrange3(dx, dy, dz)) {
        qx = pch_div_by_nocheck(dx, cell, rcell); qy = pch_div_by_nocheck(dy, cell, rcell); qz = pch_div_by_nocheck(dz, cell, rcell);
    } else {
        qx = __ddiv_rn(dx, cell); qy = __ddiv_rn(dy, cell); qz = __ddiv_rn(dz, cell);
    }
    cx = (uint64_t)(long long)floor(qx); cy = (uint64_t)(long long)floor(qy); cz = (uint64_t)(long long)floor(qz);
}

// four consecutive points per thread: three 16-byte loads in, two 16-byte stores out
__global__ void __launch_bounds__(256)
k_db_keys(const float* __restrict__ P, DbGeom g, const uint32_t* __restrict__ mm, double rcell, int fast,
          uint64_t* __restrict__ keys) {
    if (!db_resolve(g)) return;
    int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_quads = (g.G + 3) / 4;
    const bool aligned = ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(keys)) & 15) == 0;
    for (; q < n_quads; q += stride) {
        const int64_t i0 = q * 4;
        float v[12];
        if (aligned && i0 + 4 <= g.G) {
            const float4* src = reinterpret_cast<const float4*>(P + i0 * 3);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 f = __ldg(src + k);
                v[4 * k + 0] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) v[k] = (i0 * 3 + k < g.G * 3) ? P[i0 * 3 + k] : 0.f;
        }
        uint64_t out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t i = i0 + e;
            const int64_t c = (i < g.G ? i : g.G - 1) / g.chunk;
            const float mnx = pch_ordered_to_f32(mm[c * 6 + 0]), mny = pch_ordered_to_f32(mm[c * 6 + 1]),
                        mnz = pch_ordered_to_f32(mm[c * 6 + 2]);
            uint64_t cx, cy, cz;
            db_cell3(v[e * 3 + 0], v[e * 3 + 1], v[e * 3 + 2], mnx, mny, mnz, g.cell, rcell, fast, cx, cy, cz);
            out[e] = (cx << g.sh_x) | (cy << g.sh_y) | (cz << g.sh_z) | (uint64_t)(i - c * g.chunk);
        }
        if (aligned && i0 + 4 <= g.G) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(keys + i0);
            dst[0] = make_ulonglong2(out[0], out[1]);
            dst[1] = make_ulonglong2(out[2], out[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e < g.G) keys[i0 + e] = out[e];
        }
    }
}

// ---------------------------------------------------------------- D2b: cells from sorted keys
#define DC_THREADS 256
#define DC_ROWS 8
#define DC_TILE (DC_THREADS * DC_ROWS)

struct DbCells {
    float4* spts;            // [G] sorted points, w = bit pattern of the in-chunk original index
    int32_t* pt_cell;        // [G] cell index of sorted position
    int32_t* inv_pos;        // [G] original global index -> sorted position
    int32_t* cell_start;     // [U+1]
    uint64_t* cell_key;      // [U] key >> bits_idx
    int32_t* chunk_cell0;    // [n_chunks+1] first cell of each chunk
    long long* n_cells;      // [1]
};

__global__ void __launch_bounds__(DC_THREADS)
k_db_cells(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ keys_odd, const float* __restrict__ P, DbGeom g, DbCells o, int64_t tiles_per_chunk,
           int64_t total_tiles, uint64_t* __restrict__ status, uint32_t* __restrict__ counter, int* __restrict__ err) {
    if (!db_resolve(g)) return;
    if (g.dplan && (((g.key_bits + 7) >> 3) & 1)) keys = keys_odd;   // odd pass count: the sorted keys are in the ping-pong partner
    __shared__ uint64_t s_keys[DC_TILE + 1];
    __shared__ uint32_t s_wcount[DC_THREADS / 32];
    __shared__ uint64_t s_off;
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= total_tiles) return;
    const int64_t c = tile / tiles_per_chunk, lt = tile - c * tiles_per_chunk;
    const int64_t cstart = c * g.chunk, cend = min(cstart + g.chunk, g.G);
    const int64_t start = cstart + lt * DC_TILE;
    const int cnt = (int)min((int64_t)DC_TILE, cend - start);
    const int bi = g.bits_idx;
    const uint64_t idx_mask = (1ull << bi) - 1ull;
    for (int i = tid; i < cnt; i += DC_THREADS) s_keys[i + 1] = keys[start + i];
    if (tid == 0) s_keys[0] = lt > 0 ? keys[start - 1] : ~0ull;
    __syncthreads();
    uint32_t incl[DC_ROWS];
    uint32_t head_bits = 0, wtotal = 0;
    const int wbase = warp * (32 * DC_ROWS);
#pragma unroll
    for (int j = 0; j < DC_ROWS; ++j) {
        int i = wbase + j * 32 + lane;
        bool head = false;
        if (i < cnt) head = (i == 0 && lt == 0) || ((s_keys[i + 1] >> bi) != (s_keys[i] >> bi));
        uint32_t b = __ballot_sync(0xffffffffu, head);
        incl[j] = wtotal + __popc(b & ((2u << lane) - 1u));  // heads up to and including this item
        wtotal += __popc(b);
        if (head) head_bits |= 1u << j;
    }
    if (lane == 0) s_wcount[warp] = wtotal;
    __syncthreads();
    uint32_t wprefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < DC_THREADS / 32; ++w) {
        uint32_t cc = s_wcount[w];
        if (w < warp) wprefix += cc;
        total += cc;
    }
    if (tid == 0) pch_lookback_publish_u64(status, tile, 0, total);
    // gather the points BEFORE the look-back walk (the chain wait overlaps the loads)
    float4 pv[DC_ROWS];
#pragma unroll
    for (int j = 0; j < DC_ROWS; ++j) {
        const int i = wbase + j * 32 + lane;
        pv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < cnt) {
            const uint64_t k = s_keys[i + 1];
            const int64_t orig = cstart + (int64_t)(k & idx_mask);
            pv[j].x = P[orig * 3 + 0]; pv[j].y = P[orig * 3 + 1]; pv[j].z = P[orig * 3 + 2];
            pv[j].w = __int_as_float((int)(k & idx_mask));
        }
    }
    if (tid == 0) {
        s_off = pch_lookback_walk_u64(status, tile, 0, total, err);
        if (tile == total_tiles - 1) {
            *o.n_cells = (long long)(s_off + total);
            o.cell_start[s_off + total] = (int32_t)g.G;
            o.chunk_cell0[g.n_chunks] = (int32_t)(s_off + total);
        }
    }
    __syncthreads();
    const uint64_t off = s_off + wprefix;
#pragma unroll
    for (int j = 0; j < DC_ROWS; ++j) {
        int i = wbase + j * 32 + lane;
        if (i >= cnt) continue;
        const int64_t pos = start + i;
        const uint64_t k = s_keys[i + 1];
        const int64_t cell = (int64_t)(off + incl[j]) - 1;  // cell that item i belongs to
        // NOTE: an item before the first head of this tile belongs to the last cell of the
        // previous tile: off+incl-1 is then exactly that cell (off counts cells before the tile).
        const int64_t orig = cstart + (int64_t)(k & idx_mask);
        o.pt_cell[pos] = (int32_t)cell;
        o.inv_pos[orig] = (int32_t)pos;
        o.spts[pos] = pv[j];
        if (head_bits & (1u << j)) {
            o.cell_start[cell] = (int32_t)pos;
            o.cell_key[cell] = k >> bi;
            if (i == 0 && lt == 0) o.chunk_cell0[c] = (int32_t)cell;
        }
    }
}

// ---------------------------------------------------------------- D2c: neighbour columns
// For cell u and each of the 25 (ox,oy) in [-2,2]^2: first cell index and count (<=5) of the
// existing cells with cz in [cz-2, cz+2] — one binary search per column.
__global__ void k_db_nbr(DbGeom g, const uint64_t* __restrict__ cell_key, const int32_t* __restrict__ cell_start,
                         const int32_t* __restrict__ chunk_cell0, const long long* __restrict__ U_dev,
                         int32_t* __restrict__ nbr_first, uint8_t* __restrict__ nbr_cnt) {
    if (!db_resolve(g)) return;
    const int64_t U = *U_dev;
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int bxy = g.bits_y + g.bits_z;
    const uint64_t mz = (1ull << g.bits_z) - 1ull, my = (1ull << g.bits_y) - 1ull;
    for (; t < U * 25; t += stride) {
        const int64_t u = t / 25;
        const int col = (int)(t - u * 25);
        const int ox = col / 5 - 2, oy = col % 5 - 2;
        const uint64_t k = cell_key[u];
        const long long cx = (long long)(k >> bxy), cy = (long long)((k >> g.bits_z) & my), cz = (long long)(k & mz);
        const long long nx = cx + ox, ny = cy + oy;
        int32_t first = 0;
        int cntc = 0;
        if (nx >= 0 && ny >= 0 && nx < (1ll << g.bits_x) && ny < (1ll << g.bits_y)) {
            const int64_t c = cell_start[u] / g.chunk;
            const long long z0 = cz - 2 < 0 ? 0 : cz - 2;
            const long long z1 = cz + 2 > (long long)mz ? (long long)mz : cz + 2;
            const uint64_t klo = ((uint64_t)nx << bxy) | ((uint64_t)ny << g.bits_z) | (uint64_t)z0;
            const uint64_t khi = ((uint64_t)nx << bxy) | ((uint64_t)ny << g.bits_z) | (uint64_t)z1;
            int32_t lo = chunk_cell0[c], hi = chunk_cell0[c + 1];
            while (lo < hi) {
                int32_t mid = (lo + hi) >> 1;
                if (cell_key[mid] < klo) lo = mid + 1; else hi = mid;
            }
            first = lo;
            const int32_t end = chunk_cell0[c + 1];
            while (first + cntc < end && cell_key[first + cntc] <= khi) ++cntc;
        }
        nbr_first[t] = first;
        nbr_cnt[t] = (uint8_t)cntc;
    }
}

__device__ __forceinline__ double db_dist2(const float4& a, const float4& b) {
    double dx = __dsub_rn((double)a.x, (double)b.x);
    double dy = __dsub_rn((double)a.y, (double)b.y);
    double dz = __dsub_rn((double)a.z, (double)b.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---------------------------------------------------------------- D3: core points
// Phase 1 (thread per point): a point of a cell holding >= min_samples points is core at once (all
// cell-mates are neighbours).  Every other point goes to a work list.  Phase 2 (one WARP per listed
// point): the lanes stride over the points of the 25 neighbour columns — the cells of one column are
// consecutive in the sorted order, so a column is ONE contiguous range — and count hits with ballots,
// leaving as soon as min_samples is reached.
__global__ void __launch_bounds__(256)
k_db_core1(DbGeom g, const int32_t* __restrict__ pt_cell, const int32_t* __restrict__ cell_start,
           uint8_t* __restrict__ core, int32_t* __restrict__ worklist, unsigned int* __restrict__ n_work) {
    if (!db_resolve(g)) return;
    const int lane = threadIdx.x & 31;
    const int64_t Gpad = (g.G + 31) / 32 * 32;
    int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; pos < Gpad; pos += stride) {
        bool pending = false;
        if (pos < g.G) {
            const int32_t u = pt_cell[pos];
            const bool dense = (cell_start[u + 1] - cell_start[u]) >= g.min_pts;
            core[pos] = dense ? 1 : 0;
            pending = !dense;
        }
        const uint32_t b = __ballot_sync(0xffffffffu, pending);
        if (b) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(n_work, (unsigned int)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pending) worklist[base + __popc(b & ((1u << lane) - 1u))] = (int32_t)pos;
        }
    }
}

// neighbour columns nearest first: a core point usually reaches min_samples in its own column
__constant__ int c_db_order[25] = {12, 7, 11, 13, 17, 6, 8, 16, 18, 2, 10, 14, 22, 1, 3, 5, 9, 15, 19, 21, 23, 0, 4, 20, 24};

// squared distance bounds from point p to the geometric box of cell (cx,cy,cz) of a chunk whose grid
// origin is mn: lo = min distance, hi = distance to the farthest corner.  `pad` widens the box so the
// rounding of floor((x-mn)/cell) can never put a point outside "its" box.
__device__ __forceinline__ void db_cell_bounds(const float4& p, const float* mn, double cell, long long cx, long long cy,
                                               long long cz, double& dmin2, double& dmax2) {
    const double pad = 1e-6;
    const double pv[3] = {(double)p.x, (double)p.y, (double)p.z};
    const long long cc[3] = {cx, cy, cz};
    dmin2 = 0.0; dmax2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double lo = (double)mn[a] + (double)cc[a] * cell - pad;
        const double hi = (double)mn[a] + (double)(cc[a] + 1) * cell + pad;
        double tmin = 0.0;
        if (pv[a] < lo) tmin = lo - pv[a]; else if (pv[a] > hi) tmin = pv[a] - hi;
        const double tmax = fmax(fabs(pv[a] - lo), fabs(hi - pv[a]));
        dmin2 += tmin * tmin;
        dmax2 += tmax * tmax;
    }
}

__global__ void __launch_bounds__(256)
k_db_core2(DbGeom g, const float4* __restrict__ spts, const int32_t* __restrict__ pt_cell,
           const int32_t* __restrict__ cell_start, const int32_t* __restrict__ nbr_first,
           const uint8_t* __restrict__ nbr_cnt, const int32_t* __restrict__ worklist,
           const unsigned int* __restrict__ n_work, uint8_t* __restrict__ core,
           const uint64_t* __restrict__ cell_key, const uint32_t* __restrict__ bounds) {
    if (!db_resolve(g)) return;
    const int lane = threadIdx.x & 31;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n = *n_work;
    const int bxy = g.bits_y + g.bits_z;
    const uint64_t mz = (1ull << g.bits_z) - 1ull, my = (1ull << g.bits_y) - 1ull;
    const double lim_lo = g.eps2 * (1.0 + 1e-9), lim_hi = g.eps2 * (1.0 - 1e-9);
    for (; w < n; w += nw) {
        const int32_t pos = worklist[w];
        const int32_t u = pt_cell[pos];
        const float4 p = spts[pos];
        const int64_t ch = pos / g.chunk;
        const float mn[3] = {pch_ordered_to_f32(bounds[ch * 6 + 0]), pch_ordered_to_f32(bounds[ch * 6 + 1]),
                             pch_ordered_to_f32(bounds[ch * 6 + 2])};
        int count = 0;
        for (int oi = 0; oi < 25 && count < g.min_pts; ++oi) {
            const int col = c_db_order[oi];
            const int nc = nbr_cnt[(int64_t)u * 25 + col];
            if (nc == 0) continue;
            const int32_t f = nbr_first[(int64_t)u * 25 + col];
            for (int kc = 0; kc < nc && count < g.min_pts; ++kc) {
                const int32_t v = f + kc;
                const int32_t b = cell_start[v], e = cell_start[v + 1];
                if (v != u) {
                    const uint64_t key = cell_key[v];
                    double dmin2, dmax2;
                    db_cell_bounds(p, mn, g.cell, (long long)(key >> bxy), (long long)((key >> g.bits_z) & my),
                                   (long long)(key & mz), dmin2, dmax2);
                    if (dmin2 > lim_lo) continue;                 // no point of this cell can be within eps
                    if (dmax2 < lim_hi) { count += e - b; continue; }   // every point of this cell is within eps
                } else {
                    count += e - b;                               // own cell: all neighbours by construction
                    continue;
                }
                for (int32_t q0 = b; q0 < e && count < g.min_pts; q0 += 128) {
                    float4 c4[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int32_t q = q0 + r * 32 + lane;
                        c4[r] = q < e ? spts[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int32_t q = q0 + r * 32 + lane;
                        const bool hit = q < e && db_dist2(p, c4[r]) <= g.eps2;
                        count += __popc(__ballot_sync(0xffffffffu, hit));
                    }
                }
            }
        }
        if (lane == 0 && count >= g.min_pts) core[pos] = 1;
    }
}

// ---------------------------------------------------------------- D4a: per-cell core summary
// cell_info[u]: number of core points and their float32 AABB (for pruning cell-pair searches)
struct __align__(16) DbCellInfo {
    float mn[3];
    int32_t n_core;
    float mx[3];
    int32_t parent;  // union-find parent (cell index); valid when n_core > 0
};

__global__ void __launch_bounds__(256)
k_db_cellinfo(const long long* __restrict__ U_dev, const float4* __restrict__ spts, const int32_t* __restrict__ cell_start,
              const uint8_t* __restrict__ core, DbCellInfo* __restrict__ info, int64_t chunk,
              int32_t* __restrict__ cell_mincore, DbGeom g, const uint64_t* __restrict__ cell_key,
              const uint32_t* __restrict__ bounds, unsigned long long* __restrict__ cell_mask) {
    if (!db_resolve(g)) return;
    const int64_t U = *U_dev;
    const int bxy = g.bits_y + g.bits_z;
    const uint64_t mzk = (1ull << g.bits_z) - 1ull, myk = (1ull << g.bits_y) - 1ull;
    // one warp per cell
    const int lane = threadIdx.x & 31;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; w < U; w += nw) {
        const int32_t b = cell_start[w], e = cell_start[w + 1];
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        int n = 0;
        int mincore = INT_MAX;
        const int64_t ch = b / chunk;
        const int32_t chunk_base = (int32_t)(ch * chunk);
        // 4x4x4 occupancy of the cell by its core points (sub-cell side = cell/4): lets the union pass
        // decide most cell pairs with integer geometry instead of point searches
        const uint64_t key = cell_key[w];
        const double lo3[3] = {(double)pch_ordered_to_f32(bounds[ch * 6 + 0]) + (double)(long long)(key >> bxy) * g.cell,
                               (double)pch_ordered_to_f32(bounds[ch * 6 + 1]) + (double)(long long)((key >> g.bits_z) & myk) * g.cell,
                               (double)pch_ordered_to_f32(bounds[ch * 6 + 2]) + (double)(long long)(key & mzk) * g.cell};
        const double inv_h = 4.0 / g.cell;
        unsigned long long mask = 0ull;
        for (int32_t q = b + lane; q < e; q += 32) {
            if (core[q]) {
                float4 p = spts[q];
                int sx = (int)floor(((double)p.x - lo3[0]) * inv_h), sy = (int)floor(((double)p.y - lo3[1]) * inv_h),
                    sz = (int)floor(((double)p.z - lo3[2]) * inv_h);
                sx = min(max(sx, 0), 3); sy = min(max(sy, 0), 3); sz = min(max(sz, 0), 3);
                mask |= 1ull << (sx * 16 + sy * 4 + sz);
                mincore = min(mincore, chunk_base + __float_as_int(p.w));
                mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
                mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
                ++n;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o);
            mincore = min(mincore, __shfl_xor_sync(0xffffffffu, mincore, o));
            mask |= __shfl_xor_sync(0xffffffffu, mask, o);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
            }
        }
        if (lane == 0) {
            DbCellInfo ci;
            ci.mn[0] = mn[0]; ci.mn[1] = mn[1]; ci.mn[2] = mn[2];
            ci.mx[0] = mx[0]; ci.mx[1] = mx[1]; ci.mx[2] = mx[2];
            ci.n_core = n;
            ci.parent = (int32_t)w;
            info[w] = ci;
            cell_mincore[w] = mincore;
            cell_mask[w] = mask;
        }
    }
}

// ---------------------------------------------------------------- D4b: union-find over core cells
__device__ __forceinline__ int32_t uf_find(DbCellInfo* info, int32_t x) {
    int32_t p = ((volatile DbCellInfo*)info)[x].parent;
    while (p != x) {
        int32_t gp = ((volatile DbCellInfo*)info)[p].parent;
        if (gp != p) atomicCAS(&info[x].parent, p, gp);  // path halving (benign race)
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(DbCellInfo* info, int32_t a, int32_t b) {
    while (true) {
        a = uf_find(info, a);
        b = uf_find(info, b);
        if (a == b) return;
        if (a > b) { int32_t t = a; a = b; b = t; }      // smaller index becomes the root
        if (atomicCAS(&info[b].parent, b, a) == b) return;
    }
}

// conservative lower bound of the squared distance from point p to an AABB, in double
__device__ __forceinline__ double db_box_dist2(const float4& p, const float* mn, const float* mx) {
    double d = 0.0;
    const float v[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double t = 0.0;
        if (v[a] < mn[a]) t = (double)mn[a] - (double)v[a];
        else if (v[a] > mx[a]) t = (double)v[a] - (double)mx[a];
        d += t * t;
    }
    return d;
}

#define UN_CAP 256
__global__ void __launch_bounds__(256)
k_db_union(DbGeom g, const long long* __restrict__ U_dev, const float4* __restrict__ spts,
           const int32_t* __restrict__ cell_start, const uint8_t* __restrict__ core,
           const int32_t* __restrict__ nbr_first, const uint8_t* __restrict__ nbr_cnt, DbCellInfo* __restrict__ info,
           const uint64_t* __restrict__ cell_key, int pass, const unsigned long long* __restrict__ cell_mask) {
    if (!db_resolve(g)) return;
    // pass 0: only face/edge/corner-adjacent cells (|offset| <= 1): cheap hits that merge almost every
    // dense region; pass 1: the remaining (distance-2) cells, most of which are then skipped by the
    // "already in one set" test instead of being searched exhaustively.
    const int64_t U = *U_dev;
    const uint64_t zmask = (1ull << g.bits_z) - 1ull;
    // one warp per cell A.  The 25 neighbour columns are inspected by 25 lanes in parallel (the table
    // and key look-ups are a chain of dependent loads: doing them lane-parallel instead of one warp task
    // per column removes most of this kernel's latency); the surviving (A, B) pairs, B > A, are queued
    // in shared memory and then processed by the whole warp one at a time.
    __shared__ float3 s_la[256 / 32][UN_CAP];
    __shared__ int2 s_cand[256 / 32][128];
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double slack = g.eps2 * (1.0 + 1e-12);  // pruning must never drop a true neighbour pair
    for (; w < U; w += nw) {
        const int32_t A = (int32_t)w;
        if (info[A].n_core == 0) continue;
        const long long czA = (long long)(cell_key[A] & zmask);
        const unsigned long long mA = cell_mask[A];
        int n_cand = 0;
        {
            int32_t f = 0;
            int nc = 0, ox_l = 0, oy_l = 0;
            bool near_xy_l = false;
            if (lane < 25) {
                f = nbr_first[(int64_t)A * 25 + lane];
                nc = nbr_cnt[(int64_t)A * 25 + lane];
                ox_l = lane / 5 - 2; oy_l = lane % 5 - 2;
                near_xy_l = ox_l >= -1 && ox_l <= 1 && oy_l >= -1 && oy_l <= 1;
                if (pass == 0 && !near_xy_l) nc = 0;
            }
            for (int k = 0; k < 5; ++k) {
                bool valid = false;
                const int32_t Bc = f + k;
                int dzl = 0;
                if (k < nc && Bc > A && info[Bc].n_core > 0) {
                    dzl = (int)((long long)(cell_key[Bc] & zmask) - czA);
                    const bool near = near_xy_l && dzl >= -1 && dzl <= 1;
                    valid = (pass == 0) == near;
                }
                const uint32_t bal = __ballot_sync(0xffffffffu, valid);
                if (valid)
                    s_cand[warp_in_block][n_cand + __popc(bal & ((1u << lane) - 1u))] =
                        make_int2(Bc, (ox_l + 2) | ((oy_l + 2) << 3) | ((dzl + 2) << 6));
                n_cand += __popc(bal);
            }
            __syncwarp();
        }
        for (int ci = 0; ci < n_cand; ++ci) {
            const int2 cd = s_cand[warp_in_block][ci];
            const int32_t B = cd.x;
            const int ox = (cd.y & 7) - 2, oy = ((cd.y >> 3) & 7) - 2;
            const long long dz = (long long)((cd.y >> 6) & 7) - 2;
            int same = 0;  // already joined? (work saving only; decided by lane 0 so the warp stays uniform)
            if (lane == 0) same = uf_find(info, A) == uf_find(info, B);
            same = __shfl_sync(0xffffffffu, same, 0);
            if (same) continue;
            {
                // Sub-cell geometry (exact, integer): a pair of occupied sub-cells whose FARTHEST corners are
                // within eps proves a core-core neighbour pair (certain hit); if even the NEAREST corners of
                // every occupied pair are beyond eps there is none (certain miss).  Units: sub-cell side
                // h = cell/4 = eps/(4*sqrt(3))*(1-1e-7), (eps/h)^2 = 48.00001; the integer thresholds 47 / 49
                // leave room for the 1e-6 m slack of the cell assignment.
                const unsigned long long mB = cell_mask[B];
                const int d0x = ox * 4, d0y = oy * 4, d0z = (int)dz * 4;
                bool hit = false, maybe = false;
                for (unsigned long long rest = mB; rest && !hit; rest &= rest - 1) {
                    const int b = __ffsll((long long)rest) - 1;
                    const int bx = b >> 4, by = (b >> 2) & 3, bz = b & 3;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int a = lane + 32 * half;
                        if ((mA >> a) & 1ull) {
                            const int dx = abs(d0x + bx - (a >> 4)), dy = abs(d0y + by - ((a >> 2) & 3)), dzz = abs(d0z + bz - (a & 3));
                            const int smax = (dx + 1) * (dx + 1) + (dy + 1) * (dy + 1) + (dzz + 1) * (dzz + 1);
                            const int mx_ = max(dx - 1, 0), my_ = max(dy - 1, 0), mz_ = max(dzz - 1, 0);
                            const int smin = mx_ * mx_ + my_ * my_ + mz_ * mz_;
                            hit = hit || smax <= 47;
                            maybe = maybe || smin < 49;
                        }
                    }
                    hit = __any_sync(0xffffffffu, hit);
                }
                if (hit) {
                    if (lane == 0) uf_union(info, A, B);
                    __syncwarp();
                    continue;
                }
                if (!__any_sync(0xffffffffu, maybe)) continue;   // certain miss
            }
            float amn[3], amx[3], bmn[3], bmx[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                amn[a] = info[A].mn[a]; amx[a] = info[A].mx[a];
                bmn[a] = info[B].mn[a]; bmx[a] = info[B].mx[a];
            }
            // box-box distance
            double bb = 0.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double t = 0.0;
                if (amx[a] < bmn[a]) t = (double)bmn[a] - (double)amx[a];
                else if (bmx[a] < amn[a]) t = (double)amn[a] - (double)bmx[a];
                bb += t * t;
            }
            if (bb > slack) continue;
            const int32_t ab = cell_start[A], ae = cell_start[A + 1];
            const int32_t bbeg = cell_start[B], bend = cell_start[B + 1];
            bool found = false;
            // A is consumed in batches of up to UN_CAP candidates (core points of A within eps of B's
            // core box), staged in shared memory; every batch is tested against the core points of B
            // that are within eps of the batch's own bounding box.
            float3* la = s_la[warp_in_block];
            int32_t a0 = ab;
            while (a0 < ae && !found) {
                int n_la = 0;
                float lmn[3] = {INFINITY, INFINITY, INFINITY}, lmx[3] = {-INFINITY, -INFINITY, -INFINITY};
                for (; a0 < ae && n_la <= UN_CAP - 32; a0 += 32) {
                    const int32_t ai = a0 + lane;
                    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
                    bool act = false;
                    if (ai < ae && core[ai]) {
                        pa = spts[ai];
                        act = db_box_dist2(pa, bmn, bmx) <= slack;
                    }
                    const uint32_t bal = __ballot_sync(0xffffffffu, act);
                    if (act) {
                        la[n_la + __popc(bal & ((1u << lane) - 1u))] = make_float3(pa.x, pa.y, pa.z);
                        lmn[0] = fminf(lmn[0], pa.x); lmn[1] = fminf(lmn[1], pa.y); lmn[2] = fminf(lmn[2], pa.z);
                        lmx[0] = fmaxf(lmx[0], pa.x); lmx[1] = fmaxf(lmx[1], pa.y); lmx[2] = fmaxf(lmx[2], pa.z);
                    }
                    n_la += __popc(bal);
                }
                if (n_la == 0) continue;
#pragma unroll
                for (int o = 16; o; o >>= 1)
#pragma unroll
                    for (int ax = 0; ax < 3; ++ax) {
                        lmn[ax] = fminf(lmn[ax], __shfl_xor_sync(0xffffffffu, lmn[ax], o));
                        lmx[ax] = fmaxf(lmx[ax], __shfl_xor_sync(0xffffffffu, lmx[ax], o));
                    }
                __syncwarp();
                for (int32_t b0 = bbeg; b0 < bend && !found; b0 += 32) {
                    const int32_t bi = b0 + lane;
                    float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
                    bool actb = false;
                    if (bi < bend && core[bi]) {
                        pb = spts[bi];
                        actb = db_box_dist2(pb, lmn, lmx) <= slack;
                    }
                    if (!__any_sync(0xffffffffu, actb)) continue;
                    bool hit = false;
                    for (int j = 0; j < n_la; ++j) {
                        const float3 q = la[j];   // broadcast read
                        const float4 qa = make_float4(q.x, q.y, q.z, 0.f);
                        hit = hit || (actb && db_dist2(qa, pb) <= g.eps2);
                        if ((j & 15) == 15 && __any_sync(0xffffffffu, hit)) break;
                    }
                    if (__any_sync(0xffffffffu, hit)) found = true;
                }
                __syncwarp();
            }
            if (found && lane == 0) uf_union(info, A, B);
            __syncwarp();
        }
    }
}

// between the two union passes: point every cell straight at its root so the "already joined" test of
// the second pass is two loads instead of two pointer chases
__global__ void k_db_compress(const long long* __restrict__ U_dev, DbCellInfo* __restrict__ info) {
    const int64_t U = *U_dev;
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; u < U; u += stride) {
        if (info[u].n_core <= 0) continue;
        int32_t r = (int32_t)u;
        while (true) {
            const int32_t p = ((volatile DbCellInfo*)info)[r].parent;
            if (p == r) break;
            r = p;
        }
        // only ever replaces a parent by an ancestor: safe against concurrent readers in this kernel
        info[u].parent = r;
    }
}

__global__ void k_db_flatten(const long long* __restrict__ U_dev, DbCellInfo* __restrict__ info, int32_t* __restrict__ cell_root) {
    const int64_t U = *U_dev;
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; u < U; u += stride) {
        int32_t r = -1;
        if (info[u].n_core > 0) {
            r = (int32_t)u;
            while (info[r].parent != r) r = info[r].parent;
        }
        cell_root[u] = r;
    }
}

// ---------------------------------------------------------------- D5: cluster heads (min core index)
__global__ void k_db_mincore(const long long* __restrict__ U_dev, const int32_t* __restrict__ cell_mincore,
                             const int32_t* __restrict__ cell_root, int32_t* __restrict__ root_min /*[U], init large*/) {
    const int64_t U = *U_dev;
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; u < U; u += stride) {
        const int32_t r = cell_root[u];
        if (r >= 0) atomicMin(&root_min[r], cell_mincore[u]);
    }
}

__global__ void k_db_heads(const long long* __restrict__ U_dev, const int32_t* __restrict__ cell_root,
                           const int32_t* __restrict__ root_min, uint8_t* __restrict__ is_head /*[G], zeroed*/) {
    const int64_t U = *U_dev;
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; u < U; u += stride)
        if (cell_root[u] == (int32_t)u) is_head[root_min[u]] = 1;
}

// head_list[r] = original index of the r-th cluster head (ascending) -> cluster id r
__global__ void k_db_cluster_ids(const long long* __restrict__ K_dev, const int32_t* __restrict__ head_list,
                                 const int32_t* __restrict__ inv_pos, const int32_t* __restrict__ pt_cell,
                                 const int32_t* __restrict__ cell_root, int32_t* __restrict__ root_label /*[U]*/) {
    const int64_t K = *K_dev;
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < K; r += stride) root_label[cell_root[pt_cell[inv_pos[head_list[r]]]]] = (int32_t)r;
}

// ---------------------------------------------------------------- D7: per-cluster reduction
// acc[k]: count, float32 AABB (ordered-uint encoded while reducing), float64 coordinate sums
struct DbClusterAcc {
    unsigned long long count;
    uint32_t mn[3], mx[3];
    double sum[3];
};

__global__ void k_db_acc_init(int64_t cap, DbClusterAcc* acc) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < cap; k += stride) {
        DbClusterAcc a;
        a.count = 0;
        for (int i = 0; i < 3; ++i) { a.mn[i] = 0xffffffffu; a.mx[i] = 0u; a.sum[i] = 0.0; }
        acc[k] = a;
    }
}

struct DbRun {
    int32_t lab;
    uint32_t cnt;
    uint32_t mn[3], mx[3];
    double sum[3];
};

__device__ __forceinline__ void db_run_reset(DbRun& r, int32_t lab) {
    r.lab = lab; r.cnt = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) { r.mn[a] = 0xffffffffu; r.mx[a] = 0u; r.sum[a] = 0.0; }
}

__device__ __forceinline__ void db_flush(DbClusterAcc* __restrict__ acc, const DbRun& r) {
    if (r.lab < 0 || r.cnt == 0) return;
    DbClusterAcc* a = &acc[r.lab];
    atomicAdd(&a->count, (unsigned long long)r.cnt);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        atomicMin(&a->mn[i], r.mn[i]);
        atomicMax(&a->mx[i], r.mx[i]);
        atomicAdd(&a->sum[i], r.sum[i]);
    }
}

// merge the 32 lane-local partial runs (same label in every lane) and flush once
__device__ __forceinline__ void db_warp_flush(DbClusterAcc* __restrict__ acc, DbRun& r, int lane) {
    if (r.lab < 0) return;   // warp-uniform: the run label is shared by all lanes
    r.cnt = __reduce_add_sync(0xffffffffu, r.cnt);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        r.mn[a] = __reduce_min_sync(0xffffffffu, r.mn[a]);
        r.mx[a] = __reduce_max_sync(0xffffffffu, r.mx[a]);
#pragma unroll
        for (int o = 16; o; o >>= 1) r.sum[a] += __shfl_xor_sync(0xffffffffu, r.sum[a], o);
    }
    if (lane == 0) db_flush(acc, r);
}

#define CR_ROWS 64
#ifndef CR_UNROLL
#define CR_UNROLL 4
#endif

// ---------------------------------------------------------------- D6: labels (core + border + noise)
// core points take their cell's cluster id (thread per point); the non-core points (again the work
// list of D3, minus those that turned out core) are searched by one warp each: smallest cluster id
// among the core points within eps, else -1.
// k_db_labels_core also performs the per-cluster reduction (D7) for the core points in the same pass: in
// cell-sorted order the cluster id is constant over long stretches, so a warp reads rows of 32 consecutive
// sorted points, keeps a lane-local running accumulator while the row's label matches, and merges + flushes
// (one set of atomics) only when the label changes.  Border points add themselves in k_db_labels_border.
__global__ void __launch_bounds__(256)
k_db_labels_core(DbGeom g, const float4* __restrict__ spts, const int32_t* __restrict__ pt_cell,
                 const uint8_t* __restrict__ core, const int32_t* __restrict__ cell_root,
                 const int32_t* __restrict__ root_label, int32_t* __restrict__ labels, int64_t cap,
                 DbClusterAcc* __restrict__ acc) {
    if (!db_resolve(g)) return;
    const int lane = threadIdx.x & 31;
    const int64_t n_rows = (g.G + 31) / 32;
    const int64_t n_groups = (n_rows + CR_ROWS - 1) / CR_ROWS;
    int64_t grp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t ngw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (; grp < n_groups; grp += ngw) {
        DbRun run;
        db_run_reset(run, -1);
        for (int j0 = 0; j0 < CR_ROWS; j0 += CR_UNROLL) {
            if ((grp * CR_ROWS + j0) >= n_rows) break;   // warp-uniform
            // the label of a core point sits behind three dependent gathers (pt_cell -> cell_root -> root_label):
            // issue each level for CR_UNROLL rows at once so that many loads are in flight per lane
            int64_t pos[CR_UNROLL];
            bool isc[CR_UNROLL];
            float4 p[CR_UNROLL];
            int32_t lab_u[CR_UNROLL];
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                pos[u] = (grp * CR_ROWS + j0 + u) * 32 + lane;
                isc[u] = (grp * CR_ROWS + j0 + u) < n_rows && pos[u] < g.G && core[pos[u]];
            }
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                lab_u[u] = isc[u] ? pt_cell[pos[u]] : 0;
                p[u] = isc[u] ? spts[pos[u]] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) lab_u[u] = isc[u] ? cell_root[lab_u[u]] : 0;
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) lab_u[u] = isc[u] ? root_label[lab_u[u]] : -1;
#pragma unroll
            for (int u = 0; u < CR_UNROLL; ++u) {
                if ((grp * CR_ROWS + j0 + u) >= n_rows) break;   // warp-uniform
                int32_t lab = lab_u[u];
                float v[3] = {p[u].x, p[u].y, p[u].z};
                if (isc[u]) {
                    const int64_t c = pos[u] / g.chunk;
                    const int64_t orig = c * g.chunk + __float_as_int(p[u].w);
                    labels[orig] = lab;
                    // statistics only for the clusters the caller made room for, and only for this rank's own points
                    if (lab >= cap || orig < g.own_lo || orig >= g.own_hi) lab = -1;
                }
                const uint32_t valid = __ballot_sync(0xffffffffu, lab >= 0);
                if (valid == 0) continue;
                const int32_t cand = __shfl_sync(0xffffffffu, lab, __ffs(valid) - 1);
                if (run.lab < 0) db_run_reset(run, cand);
                if (!__any_sync(0xffffffffu, lab == run.lab)) {   // the current run ended before this row
                    db_warp_flush(acc, run, lane);
                    db_run_reset(run, cand);
                }
                if (lab == run.lab) {
                    run.cnt += 1;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const uint32_t uo = pch_f32_to_ordered(v[a]);
                        run.mn[a] = min(run.mn[a], uo);
                        run.mx[a] = max(run.mx[a], uo);
                        run.sum[a] += (double)v[a];
                    }
                } else if (lab >= 0) {                            // a second label inside the row: rare, direct atomics
                    DbClusterAcc* a = &acc[lab];
                    atomicAdd(&a->count, 1ull);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t uo = pch_f32_to_ordered(v[k]);
                        atomicMin(&a->mn[k], uo);
                        atomicMax(&a->mx[k], uo);
                        atomicAdd(&a->sum[k], (double)v[k]);
                    }
                }
            }
        }
        db_warp_flush(acc, run, lane);
    }
}

__global__ void __launch_bounds__(256)
k_db_labels_border(DbGeom g, const float4* __restrict__ spts, const int32_t* __restrict__ pt_cell,
                   const int32_t* __restrict__ cell_start, const uint8_t* __restrict__ core,
                   const int32_t* __restrict__ nbr_first, const uint8_t* __restrict__ nbr_cnt,
                   const int32_t* __restrict__ cell_root, const int32_t* __restrict__ root_label,
                   const int32_t* __restrict__ worklist, const unsigned int* __restrict__ n_work,
                   int32_t* __restrict__ labels, int64_t cap, DbClusterAcc* __restrict__ acc) {
    if (!db_resolve(g)) return;
    const int lane = threadIdx.x & 31;
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n = *n_work;
    for (; w < n; w += nw) {
        const int32_t pos = worklist[w];
        if (core[pos]) continue;  // warp-uniform
        const int32_t u = pt_cell[pos];
        const float4 p = spts[pos];
        int32_t best = INT_MAX;
        for (int col = 0; col < 25; ++col) {
            const int nc = nbr_cnt[(int64_t)u * 25 + col];
            if (nc == 0) continue;
            const int32_t f = nbr_first[(int64_t)u * 25 + col];
            const int32_t b = cell_start[f], e = cell_start[f + nc];
            for (int32_t q0 = b; q0 < e; q0 += 32) {
                const int32_t q = q0 + lane;
                if (q < e && core[q]) {
                    const int32_t cl = root_label[cell_root[pt_cell[q]]];
                    if (cl >= 0 && cl < best && db_dist2(p, spts[q]) <= g.eps2) best = cl;
                }
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (lane == 0) {
            const int64_t c = pos / g.chunk;
            const int64_t orig = c * g.chunk + __float_as_int(p.w);
            labels[orig] = best == INT_MAX ? -1 : best;
            if (best != INT_MAX && best < cap && orig >= g.own_lo && orig < g.own_hi) {     // the border point joins its cluster's statistics
                DbClusterAcc* a = &acc[best];
                atomicAdd(&a->count, 1ull);
                const float v[3] = {p.x, p.y, p.z};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t u = pch_f32_to_ordered(v[k]);
                    atomicMin(&a->mn[k], u);
                    atomicMax(&a->mx[k], u);
                    atomicAdd(&a->sum[k], (double)v[k]);
                }
            }
        }
    }
}

__global__ void k_db_acc_finish(int64_t cap, const long long* __restrict__ K_dev, const DbClusterAcc* __restrict__ acc,
                                pch_cluster_stats* __restrict__ out) {
    const int64_t K = K_dev ? min((int64_t)*K_dev, cap) : cap;
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < K; k += stride) {
        pch_cluster_stats s;
        s.count = (int64_t)acc[k].count;
        for (int i = 0; i < 3; ++i) {
            s.min[i] = pch_ordered_to_f32(acc[k].mn[i]);
            s.max[i] = pch_ordered_to_f32(acc[k].mx[i]);
            s.sum[i] = acc[k].sum[i];
        }
        out[k] = s;
    }
}

// ================================================================================================
// host side
// ================================================================================================
static double db_cell_side(double eps) { return eps / sqrt(3.0) * (1.0 - 1e-7); }

extern "C" int pch_dbscan_plan(const float* P, int64_t G, int64_t chunk, double eps, uint32_t* bounds_dev,
                               pch_voxel_plan* plan_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(G >= 1 && chunk >= 1, "G and chunk must be >= 1");
    PCH_CHECK_ARG(eps > 0.0, "eps must be > 0");
    PCH_CHECK_ARG(G < (1ll << 31), "more than 2^31-1 candidate points");
    PCH_CHECK_ARG(P && bounds_dev && plan_dev, "null pointer");
    if (chunk > G) chunk = G;
    int64_t n_chunks = pch_ceil_div(G, chunk);
    PCH_LAUNCH(st, "k_db_bounds_init", k_db_bounds_init<<<(unsigned)pch_ceil_div(n_chunks * 6, 256), 256, 0, st>>>(bounds_dev, n_chunks));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_bounds", k_db_bounds<<<db_grid(G, 4096, 16), 256, 0, st>>>(P, G, chunk, bounds_dev));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_plan", k_db_plan<<<1, 256, 0, st>>>(bounds_dev, n_chunks, chunk, db_cell_side(eps), (DbPlan*)plan_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

struct DbWs {
    size_t total;
    size_t keys, tmp, sortws, sortws_bytes, spts, pt_cell, inv_pos, cell_start, cell_key, chunk_cell0, scalars,
        nbr_first, nbr_cnt, core, info, cell_root, root_min, root_label, is_head, head_list, scan_status, acc,
        worklist, cell_mincore, cell_mask, bounds, plan;
};

extern "C" size_t pch_sort_workspace_bytes(int64_t n, int64_t seg_size, int32_t bit_lo, int32_t bit_hi);
extern "C" int pch_sort_u64_segmented(uint64_t* keys, uint64_t* tmp, int64_t n, int64_t seg_size, int32_t bit_lo,
                                      int32_t bit_hi, void* workspace, size_t workspace_bytes, pch_stream_t stream);
extern "C" size_t pch_compact_workspace_bytes(int64_t m);
extern "C" int pch_compact_points(const float* xyz, const float* zs, const uint8_t* keep_mask, int64_t m,
                                  const float* centroid3, float thr, float* out_xyz, int32_t* out_src,
                                  uint8_t* out_mask, int64_t* count_dev, void* workspace, size_t workspace_bytes,
                                  pch_stream_t stream);

static int db_max_passes(int64_t chunk) {   // radix passes of the widest cell key a chunk of this size allows (62 bits in all)
    int bits_idx = 0;
    for (int64_t v = chunk - 1; v > 0; v >>= 1) ++bits_idx;
    int p = (62 - bits_idx + 7) / 8;
    return p < 1 ? 1 : (p > RS_MAX_PASSES ? RS_MAX_PASSES : p);
}

static DbWs db_ws(int64_t G, int64_t chunk, int sort_passes, int64_t max_clusters) {
    DbWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += pch_align_up(bytes, 256); return o; };
    int64_t n_chunks = pch_ceil_div(G, chunk);
    int64_t tiles = n_chunks * pch_ceil_div(chunk, DC_TILE);
    w.scalars = take(256);  // [0]=err int, [64]=tile counter, [128]=n_cells (i64), [136]=n_clusters (i64)
    w.keys = take((size_t)G * 8);
    w.tmp = take((size_t)G * 8);
    w.sortws_bytes = pch_sort_ws(pch_sort_geom(G, chunk, 0, 0, sort_passes < 1 ? 1 : sort_passes), nullptr).bytes;
    w.sortws = take(w.sortws_bytes);
    w.spts = take((size_t)G * 16);
    w.pt_cell = take((size_t)G * 4);
    w.inv_pos = take((size_t)G * 4);
    w.cell_start = take((size_t)(G + 1) * 4);
    w.cell_key = take((size_t)G * 8);
    w.chunk_cell0 = take((size_t)(n_chunks + 1) * 4);
    w.nbr_first = take((size_t)G * 25 * 4);
    w.nbr_cnt = take((size_t)G * 25);
    w.core = take((size_t)G);
    w.info = take((size_t)G * sizeof(DbCellInfo));
    w.cell_root = take((size_t)G * 4);
    w.root_min = take((size_t)G * 4);
    w.root_label = take((size_t)G * 4);
    w.is_head = take((size_t)G);
    w.head_list = take((size_t)G * 4);
    size_t sc = (size_t)tiles * 8 + 256;
    size_t cw = pch_compact_workspace_bytes(G);
    w.scan_status = take(sc > cw ? sc : cw);
    w.acc = take((size_t)max_clusters * sizeof(DbClusterAcc));
    w.worklist = take((size_t)(G + 32) * 4);
    w.cell_mincore = take((size_t)G * 4);
    w.cell_mask = take((size_t)G * 8);
    w.bounds = take((size_t)n_chunks * 6 * 4);
    w.plan = take(256);
    w.total = off;
    return w;
}

extern "C" size_t pch_dbscan_workspace_bytes(int64_t G, int64_t chunk, const pch_voxel_plan* plan, int64_t max_clusters) {
    if (G <= 0 || !plan) return 256;
    if (chunk > G) chunk = G;
    return db_ws(G, chunk, plan->n_passes, max_clusters).total;
}

// hplan: the plan on the HOST (pch_dbscan_plan + a read-back, the two-call interface) or NULL;
// dplan: the same plan in DEVICE memory (pch_dbscan, no read-back between planning and clustering).
static int dbscan_impl(const float* P, int64_t G, int64_t chunk, double eps, int32_t min_pts,
                       const uint32_t* bounds_dev, const pch_voxel_plan* hplan, const pch_voxel_plan* dplan,
                       int sort_passes, int32_t* labels_dev, int64_t* n_clusters_dev, pch_cluster_stats* stats_dev,
                       int64_t max_clusters, void* workspace, size_t workspace_bytes, pch_stream_t stream,
                       bool core_only = false) {
    cudaStream_t st = (cudaStream_t)stream;
    DbWs w = db_ws(G, chunk, sort_passes, max_clusters);
    if (workspace_bytes < w.total) {
        pch_set_error("dbscan workspace too small: %zu < %zu", workspace_bytes, w.total);
        return PCH_ERR_WORKSPACE;
    }
    uint8_t* base = (uint8_t*)workspace;
    int* err = (int*)(base + w.scalars);
    uint32_t* counter = (uint32_t*)(base + w.scalars + 64);
    long long* U_dev = (long long*)(base + w.scalars + 128);
    unsigned int* n_work = (unsigned int*)(base + w.scalars + 192);
    int32_t* worklist = (int32_t*)(base + w.worklist);
    int32_t* cell_mincore = (int32_t*)(base + w.cell_mincore);
    unsigned long long* cell_mask = (unsigned long long*)(base + w.cell_mask);
    uint64_t* keys = (uint64_t*)(base + w.keys);
    uint64_t* tmp = (uint64_t*)(base + w.tmp);

    DbGeom g;
    g.G = G; g.chunk = chunk; g.n_chunks = pch_ceil_div(G, chunk);
    g.cell = db_cell_side(eps);
    g.eps2 = eps * eps;
    g.min_pts = min_pts;
    g.dplan = (const DbPlan*)dplan;
    g.own_lo = 0; g.own_hi = G;
    g.bits_idx = g.bits_x = g.bits_y = g.bits_z = g.key_bits = g.sh_x = g.sh_y = g.sh_z = 0;
    if (hplan) {
        g.bits_idx = hplan->bits_idx;
        g.bits_x = hplan->bits_x; g.bits_y = hplan->bits_y; g.bits_z = hplan->bits_z;
        g.key_bits = hplan->key_bits;
        g.sh_z = hplan->bits_idx; g.sh_y = g.sh_z + hplan->bits_z; g.sh_x = g.sh_y + hplan->bits_y;
    }

    PCH_CUDA(cudaMemsetAsync(base + w.scalars, 0, 256, st));
    PCH_LAUNCH(st, "k_db_keys", k_db_keys<<<db_grid((G + 3) / 4, 256, 16), 256, 0, st>>>(P, g, bounds_dev, 1.0 / g.cell, pch_recip_ok(g.cell) ? 1 : 0, keys));
    PCH_LAUNCH_CHECK();
    int rc;
    const uint64_t* skeys = keys;
    if (hplan) {
        rc = pch_sort_u64_segmented(keys, tmp, G, chunk, hplan->bits_idx, hplan->bits_idx + hplan->key_bits,
                                    base + w.sortws, w.sortws_bytes, stream);
        if (rc) return rc;
        skeys = (hplan->n_passes % 2) ? tmp : keys;
    } else {
        SortGeom sg = pch_sort_geom(G, chunk, 0, 0, sort_passes);
        if ((rc = pch_sort_prepare(sg, base + w.sortws, w.sortws_bytes, st))) return rc;
        if ((rc = pch_sort_run(keys, tmp, sg, dplan, sort_passes, false, base + w.sortws, w.sortws_bytes, st))) return rc;
    }

    DbCells o;
    o.spts = (float4*)(base + w.spts);
    o.pt_cell = (int32_t*)(base + w.pt_cell);
    o.inv_pos = (int32_t*)(base + w.inv_pos);
    o.cell_start = (int32_t*)(base + w.cell_start);
    o.cell_key = (uint64_t*)(base + w.cell_key);
    o.chunk_cell0 = (int32_t*)(base + w.chunk_cell0);
    o.n_cells = U_dev;
    int64_t tiles_per_chunk = pch_ceil_div(chunk, DC_TILE);
    int64_t last = G - (g.n_chunks - 1) * chunk;
    int64_t total_tiles = (g.n_chunks - 1) * tiles_per_chunk + pch_ceil_div(last, DC_TILE);
    uint64_t* status = (uint64_t*)(base + w.scan_status);
    PCH_CUDA(cudaMemsetAsync(status, 0, (size_t)total_tiles * 8, st));
    PCH_LAUNCH(st, "k_db_cells", k_db_cells<<<(unsigned)total_tiles, DC_THREADS, 0, st>>>(skeys, tmp, P, g, o, tiles_per_chunk, total_tiles, status, counter, err));
    PCH_LAUNCH_CHECK();

    int32_t* nbr_first = (int32_t*)(base + w.nbr_first);
    uint8_t* nbr_cnt = (uint8_t*)(base + w.nbr_cnt);
    uint8_t* core = base + w.core;
    DbCellInfo* info = (DbCellInfo*)(base + w.info);
    int32_t* cell_root = (int32_t*)(base + w.cell_root);
    int32_t* root_min = (int32_t*)(base + w.root_min);
    int32_t* root_label = (int32_t*)(base + w.root_label);
    uint8_t* is_head = base + w.is_head;
    int32_t* head_list = (int32_t*)(base + w.head_list);
    DbClusterAcc* acc = (DbClusterAcc*)(base + w.acc);

    PCH_LAUNCH(st, "k_db_nbr", k_db_nbr<<<db_grid(G, 256), 256, 0, st>>>(g, o.cell_key, o.cell_start, o.chunk_cell0, U_dev, nbr_first, nbr_cnt));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_core1", k_db_core1<<<db_grid(G, 256), 256, 0, st>>>(g, o.pt_cell, o.cell_start, core, worklist, n_work));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_core2", k_db_core2<<<db_grid(G, 8, 16), 256, 0, st>>>(g, o.spts, o.pt_cell, o.cell_start, nbr_first, nbr_cnt,
                                                                            worklist, n_work, core, o.cell_key, bounds_dev));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_cellinfo", k_db_cellinfo<<<db_grid(G, 256 / 32 * 8), 256, 0, st>>>(U_dev, o.spts, o.cell_start, core, info, chunk, cell_mincore, g, o.cell_key, bounds_dev, cell_mask));
    PCH_LAUNCH_CHECK();
    for (int pass = 0; pass < 2; ++pass) {
        PCH_LAUNCH(st, "k_db_union", k_db_union<<<db_grid(G, 64, 16), 256, 0, st>>>(g, U_dev, o.spts, o.cell_start, core, nbr_first,
                                                                                 nbr_cnt, info, o.cell_key, pass, cell_mask));
        PCH_LAUNCH_CHECK();
        if (pass == 0) {
            PCH_LAUNCH(st, "k_db_compress", k_db_compress<<<db_grid(G, 256), 256, 0, st>>>(U_dev, info));
            PCH_LAUNCH_CHECK();
        }
    }
    PCH_LAUNCH(st, "k_db_flatten", k_db_flatten<<<db_grid(G, 256), 256, 0, st>>>(U_dev, info, cell_root));
    PCH_LAUNCH_CHECK();
    PCH_CUDA(cudaMemsetAsync(root_min, 0x7f, (size_t)G * 4, st));
    PCH_CUDA(cudaMemsetAsync(is_head, 0, (size_t)G, st));
    PCH_LAUNCH(st, "k_db_mincore", k_db_mincore<<<db_grid(G, 256), 256, 0, st>>>(U_dev, cell_mincore, cell_root, root_min));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_heads", k_db_heads<<<db_grid(G, 256), 256, 0, st>>>(U_dev, cell_root, root_min, is_head));
    PCH_LAUNCH_CHECK();
    // cluster ids = rank of each head in original order (exclusive scan of head flags, all chunks at once:
    // this IS the reference's running `current_label` offset)
    rc = pch_compact_points(P, nullptr, is_head, G, nullptr, 0.f, nullptr, head_list, nullptr, n_clusters_dev,
                            base + w.scan_status, pch_compact_workspace_bytes(G), stream);
    if (rc) return rc;
    PCH_LAUNCH(st, "k_db_cluster_ids", k_db_cluster_ids<<<db_grid(G, 256), 256, 0, st>>>((const long long*)n_clusters_dev, head_list, o.inv_pos, o.pt_cell,
                                                      cell_root, root_label));
    PCH_LAUNCH_CHECK();
    // labels + per-cluster reduction in one pass over the cell-sorted points (core points: run-merged; border
    // points: per-point atomics from the border search)
    PCH_LAUNCH(st, "k_db_acc_init", k_db_acc_init<<<db_grid(max_clusters, 256), 256, 0, st>>>(max_clusters, acc));
    PCH_LAUNCH_CHECK();
    if (core_only) PCH_CUDA(cudaMemsetAsync(labels_dev, 0xff, (size_t)G * 4, st));   // non-core points stay -1 until pch_dbscan_finish
    PCH_LAUNCH(st, "k_db_labels_core", k_db_labels_core<<<db_grid(G, 8 * 32 * CR_ROWS, 16), 256, 0, st>>>(g, o.spts, o.pt_cell, core, cell_root, root_label,
                                                                                            labels_dev, max_clusters, acc));
    PCH_LAUNCH_CHECK();
    if (core_only) return PCH_OK;
    PCH_LAUNCH(st, "k_db_labels_border", k_db_labels_border<<<db_grid(G, 8, 16), 256, 0, st>>>(g, o.spts, o.pt_cell, o.cell_start, core, nbr_first,
                                                                                    nbr_cnt, cell_root, root_label, worklist, n_work,
                                                                                    labels_dev, max_clusters, acc));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_acc_finish", k_db_acc_finish<<<db_grid(max_clusters, 256), 256, 0, st>>>(max_clusters, (const long long*)n_clusters_dev, acc, stats_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


extern "C" int pch_dbscan_run(const float* P, int64_t G, int64_t chunk, double eps, int32_t min_pts,
                              const uint32_t* bounds_dev, const pch_voxel_plan* plan, int32_t* labels_dev,
                              int64_t* n_clusters_dev, pch_cluster_stats* stats_dev, int64_t max_clusters,
                              void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    PCH_CHECK_ARG(G >= 1 && chunk >= 1 && min_pts >= 1 && eps > 0.0, "bad G/chunk/min_samples/eps");
    PCH_CHECK_ARG(G < (1ll << 31), "more than 2^31-1 candidate points");
    PCH_CHECK_ARG(P && bounds_dev && plan && labels_dev && n_clusters_dev && workspace, "null pointer");
    PCH_CHECK_ARG(max_clusters >= 1 && (stats_dev != nullptr), "stats buffer required");
    if (plan->status != PCH_OK) {
        pch_set_error("DBSCAN cell grid needs %d+%d key bits", plan->key_bits, plan->bits_idx);
        return PCH_ERR_RANGE;
    }
    if (chunk > G) chunk = G;
    return dbscan_impl(P, G, chunk, eps, min_pts, bounds_dev, plan, nullptr, plan->n_passes, labels_dev, n_clusters_dev,
                       stats_dev, max_clusters, workspace, workspace_bytes, stream);
}

// One call, no host round trip inside: chunk bounds -> plan (device) -> clustering.  The first 256 bytes of the
// workspace are the scalar block the caller reads back once: int32 @0 error word, int64 @128 cells, int64 @136
// clusters, uint32 @192 points outside dense cells, pch_voxel_plan @208 (status != 0: the cell grid does not fit
// one key word and nothing was clustered).
__global__ void k_db_head(const DbPlan* __restrict__ plan, uint8_t* __restrict__ scalars, const int* __restrict__ sort_err) {
    if (threadIdx.x == 0) {
        *reinterpret_cast<DbPlan*>(scalars + 208) = *plan;
        if (*sort_err) atomicExch(reinterpret_cast<int*>(scalars), *sort_err);
    }
}

extern "C" size_t pch_dbscan_fused_workspace_bytes(int64_t G, int64_t chunk, int64_t max_clusters) {
    if (G <= 0) return 256;
    if (chunk <= 0 || chunk > G) chunk = G;
    return db_ws(G, chunk, db_max_passes(chunk), max_clusters).total;
}

extern "C" int pch_dbscan(const float* P, int64_t G, int64_t chunk, double eps, int32_t min_pts, int32_t* labels_dev,
                          pch_cluster_stats* stats_dev, int64_t max_clusters, void* workspace, size_t workspace_bytes,
                          pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(G >= 1 && chunk >= 1 && min_pts >= 1 && eps > 0.0, "bad G/chunk/min_samples/eps");
    PCH_CHECK_ARG(G < (1ll << 31), "more than 2^31-1 candidate points");
    PCH_CHECK_ARG(P && labels_dev && workspace, "null pointer");
    PCH_CHECK_ARG(max_clusters >= 1 && (stats_dev != nullptr), "stats buffer required");
    if (chunk > G) chunk = G;
    const int passes = db_max_passes(chunk);
    DbWs w = db_ws(G, chunk, passes, max_clusters);
    if (workspace_bytes < w.total) {
        pch_set_error("dbscan workspace too small: %zu < %zu", workspace_bytes, w.total);
        return PCH_ERR_WORKSPACE;
    }
    uint8_t* base = (uint8_t*)workspace;
    uint32_t* bounds = (uint32_t*)(base + w.bounds);
    pch_voxel_plan* plan_dev = (pch_voxel_plan*)(base + w.plan);
    int rc = pch_dbscan_plan(P, G, chunk, eps, bounds, plan_dev, stream);
    if (rc) return rc;
    int64_t* n_clusters_dev = (int64_t*)(base + w.scalars + 136);
    rc = dbscan_impl(P, G, chunk, eps, min_pts, bounds, nullptr, plan_dev, passes, labels_dev, n_clusters_dev, stats_dev,
                     max_clusters, workspace, workspace_bytes, stream);
    if (rc) return rc;
    PCH_LAUNCH(st, "k_db_head", k_db_head<<<1, 32, 0, st>>>((const DbPlan*)plan_dev, base + w.scalars, (const int*)(base + w.sortws)));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


// ------------------------------------------------------------------------------------------------
// Two-phase clustering for spatial tiles with a halo (DBSCAN row; parity definition = the
// un-chunked variant test/zzzzz.py:79-84 on the concatenated cloud):
//   pch_dbscan_cores : everything up to the cluster ids of the CORE points (labels_dev: local id for core
//                      points, -1 elsewhere).  Between the phases the ranks exchange which local clusters are
//                      the same global cluster (shared halo points) and agree on global ids.
//   pch_dbscan_finish: the local ids are replaced by global ids (map_dev[local] = global id or -1), core labels
//                      are rewritten, border points take the smallest GLOBAL id among the clusters that own a
//                      core point within eps (sklearn's rule on the concatenated cloud), and the per-cluster
//                      statistics count only original indices in [own_lo, own_hi) (halo points belong to the
//                      neighbour).  Uses the workspace of the first phase unchanged.
// ------------------------------------------------------------------------------------------------
extern "C" int pch_dbscan_cores(const float* P, int64_t G, int64_t chunk, double eps, int32_t min_pts, int32_t* labels_dev,
                                int64_t max_clusters, void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(G >= 1 && chunk >= 1 && min_pts >= 1 && eps > 0.0, "bad G/chunk/min_samples/eps");
    PCH_CHECK_ARG(G < (1ll << 31), "more than 2^31-1 candidate points");
    PCH_CHECK_ARG(P && labels_dev && workspace && max_clusters >= 1, "null pointer");
    if (chunk > G) chunk = G;
    const int passes = db_max_passes(chunk);
    DbWs w = db_ws(G, chunk, passes, max_clusters);
    if (workspace_bytes < w.total) {
        pch_set_error("dbscan workspace too small: %zu < %zu", workspace_bytes, w.total);
        return PCH_ERR_WORKSPACE;
    }
    uint8_t* base = (uint8_t*)workspace;
    uint32_t* bounds = (uint32_t*)(base + w.bounds);
    pch_voxel_plan* plan_dev = (pch_voxel_plan*)(base + w.plan);
    int rc = pch_dbscan_plan(P, G, chunk, eps, bounds, plan_dev, stream);
    if (rc) return rc;
    int64_t* n_clusters_dev = (int64_t*)(base + w.scalars + 136);
    rc = dbscan_impl(P, G, chunk, eps, min_pts, bounds, nullptr, plan_dev, passes, labels_dev, n_clusters_dev, nullptr,
                     max_clusters, workspace, workspace_bytes, stream, true);
    if (rc) return rc;
    PCH_LAUNCH(st, "k_db_head", k_db_head<<<1, 32, 0, st>>>((const DbPlan*)plan_dev, base + w.scalars, (const int*)(base + w.sortws)));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

__global__ void k_db_relabel_roots(const long long* __restrict__ U_dev, const long long* __restrict__ K_dev,
                                   const int32_t* __restrict__ cell_root, const int32_t* __restrict__ map,
                                   int32_t* __restrict__ root_label) {
    const int64_t U = *U_dev, K = *K_dev;
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; u < U; u += stride)
        if (cell_root[u] == (int32_t)u) {
            const int32_t l = root_label[u];
            root_label[u] = (l >= 0 && l < K) ? map[l] : -1;
        }
}

extern "C" size_t pch_dbscan_acc_bytes(int64_t n_clusters) { return (size_t)(n_clusters > 0 ? n_clusters : 1) * sizeof(DbClusterAcc); }

extern "C" int pch_dbscan_finish(const float* P, int64_t G, int64_t chunk, double eps, int32_t min_pts,
                                 const int32_t* map_dev, int64_t n_global, int64_t own_lo, int64_t own_hi,
                                 int32_t* labels_dev, pch_cluster_stats* stats_dev, void* acc_dev,
                                 int64_t max_clusters_phase1, void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(G >= 1 && chunk >= 1 && min_pts >= 1 && eps > 0.0, "bad G/chunk/min_samples/eps");
    PCH_CHECK_ARG(P && map_dev && labels_dev && workspace, "null pointer");
    PCH_CHECK_ARG(n_global >= 0 && own_lo >= 0 && own_hi >= own_lo && own_hi <= G, "bad global cluster count / own range");
    PCH_CHECK_ARG(n_global == 0 || (stats_dev && acc_dev), "stats and accumulator buffers required");
    if (chunk > G) chunk = G;
    const int passes = db_max_passes(chunk);
    DbWs w = db_ws(G, chunk, passes, max_clusters_phase1);
    if (workspace_bytes < w.total) {
        pch_set_error("dbscan workspace too small: %zu < %zu", workspace_bytes, w.total);
        return PCH_ERR_WORKSPACE;
    }
    uint8_t* base = (uint8_t*)workspace;
    long long* U_dev = (long long*)(base + w.scalars + 128);
    long long* K_dev = (long long*)(base + w.scalars + 136);
    unsigned int* n_work = (unsigned int*)(base + w.scalars + 192);
    DbGeom g;
    g.G = G; g.chunk = chunk; g.n_chunks = pch_ceil_div(G, chunk);
    g.cell = db_cell_side(eps);
    g.eps2 = eps * eps;
    g.min_pts = min_pts;
    g.dplan = (const DbPlan*)(base + w.plan);
    g.own_lo = own_lo; g.own_hi = own_hi;
    g.bits_idx = g.bits_x = g.bits_y = g.bits_z = g.key_bits = g.sh_x = g.sh_y = g.sh_z = 0;
    const float4* spts = (const float4*)(base + w.spts);
    const int32_t* pt_cell = (const int32_t*)(base + w.pt_cell);
    const int32_t* cell_start = (const int32_t*)(base + w.cell_start);
    const int32_t* nbr_first = (const int32_t*)(base + w.nbr_first);
    const uint8_t* nbr_cnt = (const uint8_t*)(base + w.nbr_cnt);
    const uint8_t* core = base + w.core;
    const int32_t* cell_root = (const int32_t*)(base + w.cell_root);
    int32_t* root_label = (int32_t*)(base + w.root_label);
    const int32_t* worklist = (const int32_t*)(base + w.worklist);
    DbClusterAcc* acc = (DbClusterAcc*)acc_dev;
    const int64_t cap = n_global > 0 ? n_global : 1;
    PCH_LAUNCH(st, "k_db_relabel_roots", k_db_relabel_roots<<<db_grid(G, 256), 256, 0, st>>>(U_dev, K_dev, cell_root, map_dev, root_label));
    PCH_LAUNCH_CHECK();
    PCH_CUDA(cudaMemsetAsync(labels_dev, 0xff, (size_t)G * 4, st));
    if (n_global > 0) {
        PCH_LAUNCH(st, "k_db_acc_init", k_db_acc_init<<<db_grid(cap, 256), 256, 0, st>>>(cap, acc));
        PCH_LAUNCH_CHECK();
    }
    PCH_LAUNCH(st, "k_db_labels_core", k_db_labels_core<<<db_grid(G, 8 * 32 * CR_ROWS, 16), 256, 0, st>>>(g, spts, pt_cell, core, cell_root, root_label,
                                                                                            labels_dev, n_global, acc));
    PCH_LAUNCH_CHECK();
    PCH_LAUNCH(st, "k_db_labels_border", k_db_labels_border<<<db_grid(G, 8, 16), 256, 0, st>>>(g, spts, pt_cell, cell_start, core, nbr_first,
                                                                                    nbr_cnt, cell_root, root_label, worklist, n_work,
                                                                                    labels_dev, n_global, acc));
    PCH_LAUNCH_CHECK();
    if (n_global > 0) {
        PCH_LAUNCH(st, "k_db_acc_finish", k_db_acc_finish<<<db_grid(cap, 256), 256, 0, st>>>(n_global, nullptr, acc, stats_dev));
        PCH_LAUNCH_CHECK();
    }
    return PCH_OK;
}

// smallest value of base + i over the points i in [lo, hi) that carry label k, for every k (atomicMin into
// table_dev[k], which the caller initialises): "smallest core index of a cluster" in a numbering that spans ranks
__global__ void k_label_min_index(const int32_t* __restrict__ labels, int64_t lo, int64_t hi, long long base, int64_t K,
                                  long long* __restrict__ table) {
    // lanes of a warp hold consecutive points, and clusters are long runs of equal labels: the lanes that share a
    // label elect their lowest lane (= their smallest index) and only that one issues the atomic
    const int lane = threadIdx.x & 31;
    int64_t i0 = lo + (blockIdx.x * (int64_t)blockDim.x + threadIdx.x - lane);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i0 < hi; i0 += stride) {
        const int64_t i = i0 + lane;
        int32_t l = -1;
        if (i < hi) {
            l = labels[i];
            if (l >= K) l = -1;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, l);
        if (l >= 0 && (peers & ((1u << lane) - 1u)) == 0) atomicMin(&table[l], base + (long long)(i - lo));
    }
}

extern "C" int pch_label_min_index(const int32_t* labels_dev, int64_t lo, int64_t hi, int64_t base, int64_t n_labels,
                                   int64_t* table_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(lo >= 0 && hi >= lo && n_labels >= 0, "bad range");
    if (hi == lo || n_labels == 0) return PCH_OK;
    PCH_CHECK_ARG(labels_dev && table_dev, "null pointer");
    PCH_LAUNCH(st, "k_label_min_index", k_label_min_index<<<db_grid(hi - lo, 256), 256, 0, st>>>(labels_dev, lo, hi, (long long)base, n_labels, (long long*)table_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

// projection of (n,3) float32 points on a horizontal axis (ux, uy), float64: s = x*ux + y*uy.
//   pch_axis_extent   : minmax_dev[2] (float64) = min and max of s (the caller initialises to +inf / -inf)
//   pch_axis_band_mask: mask_dev[i] = lo <= s_i <= hi, and (labels_dev != NULL) labels_dev[i] >= 0
// Used to pick the halo a tile sends to its neighbour and the zone in which both ranks know a point's core
// status exactly.
__device__ __forceinline__ double axis_s(const float* __restrict__ P, int64_t i, double ux, double uy) {
    return __dadd_rn(__dmul_rn((double)P[i * 3 + 0], ux), __dmul_rn((double)P[i * 3 + 1], uy));
}
__global__ void k_axis_extent(const float* __restrict__ P, int64_t n, double ux, double uy, unsigned long long* __restrict__ mm) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double lo = __longlong_as_double(0x7ff0000000000000ll), hi = -lo;
    for (; i < n; i += stride) {
        const double s = axis_s(P, i, ux, uy);
        lo = fmin(lo, s); hi = fmax(hi, s);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // order-preserving encoding of the doubles so that integer atomics reduce them
        auto enc = [](double d) { unsigned long long u = (unsigned long long)__double_as_longlong(d); return (u >> 63) ? ~u : (u | 0x8000000000000000ull); };
        atomicMin(&mm[0], enc(lo));
        atomicMax(&mm[1], enc(hi));
    }
}
__global__ void k_axis_extent_init(unsigned long long* mm) { mm[0] = ~0ull; mm[1] = 0ull; }
__global__ void k_axis_extent_finish(unsigned long long* mm) {
    for (int k = 0; k < 2; ++k) {
        const unsigned long long u = mm[k];
        mm[k] = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    }
}
__global__ void k_axis_band_mask(const float* __restrict__ P, int64_t n, double ux, double uy, double lo, double hi,
                                 const int32_t* __restrict__ labels, uint8_t* __restrict__ mask) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double s = axis_s(P, i, ux, uy);
        mask[i] = (s >= lo && s <= hi && (!labels || labels[i] >= 0)) ? 1 : 0;
    }
}

extern "C" int pch_axis_extent(const float* xyz_dev, int64_t n, double ux, double uy, double* minmax_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && minmax_dev && (n == 0 || xyz_dev), "bad arguments");
    PCH_LAUNCH(st, "k_axis_extent_init", k_axis_extent_init<<<1, 1, 0, st>>>((unsigned long long*)minmax_dev));
    if (n > 0) PCH_LAUNCH(st, "k_axis_extent", k_axis_extent<<<db_grid(n, 256), 256, 0, st>>>(xyz_dev, n, ux, uy, (unsigned long long*)minmax_dev));
    PCH_LAUNCH(st, "k_axis_extent_finish", k_axis_extent_finish<<<1, 1, 0, st>>>((unsigned long long*)minmax_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_axis_band_mask(const float* xyz_dev, int64_t n, double ux, double uy, double lo, double hi,
                                  const int32_t* labels_dev, uint8_t* mask_dev, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0 && (n == 0 || (xyz_dev && mask_dev)), "bad arguments");
    if (n == 0) return PCH_OK;
    PCH_LAUNCH(st, "k_axis_band_mask", k_axis_band_mask<<<db_grid(n, 256), 256, 0, st>>>(xyz_dev, n, ux, uy, lo, hi, labels_dev, mask_dev));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}


// ------------------------------------------------------------------------------------------------
// cluster-major point order: what `filtered_points[all_labels == label]` (utils/tower_extraction.py:
// 133-134) gathers for every label, done once for all labels: words (label << 32 | index) of the
// labelled points are sorted by label (stable, so the index order inside a cluster is preserved) and
// the rows gathered.  The reference's per-cluster mask scans are O(G*K); this is O(G).
// ------------------------------------------------------------------------------------------------
__global__ void k_label_words(const int32_t* __restrict__ labels, int64_t G, uint64_t* __restrict__ words) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < G; i += stride) {
        const int32_t l = labels[i];
        // noise (-1) gets the largest label value so it sorts to the end
        words[i] = ((uint64_t)(uint32_t)(l < 0 ? 0x7fffffff : l) << 32) | (uint64_t)i;
    }
}

__global__ void k_gather_rows(const float* __restrict__ P, const uint64_t* __restrict__ words, int64_t m,
                              float* __restrict__ out, int32_t* __restrict__ src_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        const int64_t j = (int64_t)(words[i] & 0xffffffffull);
        out[i * 3 + 0] = P[j * 3 + 0];
        out[i * 3 + 1] = P[j * 3 + 1];
        out[i * 3 + 2] = P[j * 3 + 2];
        if (src_idx) src_idx[i] = (int32_t)j;
    }
}

extern "C" int pch_label_words(const int32_t* labels, int64_t G, uint64_t* words, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(G >= 0 && G < (1ll << 31), "bad size");
    if (G == 0) return PCH_OK;
    PCH_CHECK_ARG(labels && words, "null pointer");
    PCH_LAUNCH(st, "k_label_words", k_label_words<<<db_grid(G, 256), 256, 0, st>>>(labels, G, words));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}

extern "C" int pch_gather_rows_f32(const float* P, const uint64_t* words, int64_t m, float* out, int32_t* src_idx,
                                   pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(m >= 0, "bad size");
    if (m == 0) return PCH_OK;
    PCH_CHECK_ARG(P && words && out, "null pointer");
    PCH_LAUNCH(st, "k_gather_rows", k_gather_rows<<<db_grid(m, 256), 256, 0, st>>>(P, words, m, out, src_idx));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
