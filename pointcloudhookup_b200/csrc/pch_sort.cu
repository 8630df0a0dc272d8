// Segmented stable LSD radix sort of 64-bit keys (8-bit digits), one-sweep style:
//   k_hist  : ONE read of the keys builds the digit histograms of every pass, per segment (skipped when the
//             kernel that produced the keys already built them: pch_sort.cuh)
//   k_scan  : exclusive scan of each 256-bin histogram -> digit bases inside the segment
//   k_pass  : per pass, one kernel: tile-local ranking (warp match_any multi-split), chained
//             decoupled look-back across the tiles of a segment, scatter via shared memory.
// Each pass therefore moves 8 B in + 8 B out per key; only the bits that carry voxel/cell index
// are sorted, and segments (voxel chunks / DBSCAN chunks) never mix because every offset is
// segment-relative.  Used for open3d's voxel grouping (ui/import_PC.py:12) and for the DBSCAN
// cell grid (utils/tower_extraction.py:107-112).
#include "pch_sort.cuh"

#ifndef LB_WIN
#define LB_WIN 8
#endif

// ticket -> tile.  Tickets are handed out group by group (RS_GROUP segments), and inside a group tile-index
// major: (lt 0 of every segment of the group), (lt 1 of every segment), ...  A tile's predecessors in its
// own segment therefore always hold smaller tickets (forward progress of the look-back), and the tiles that
// run concurrently belong to different segments.
__device__ __forceinline__ bool tile_of_ticket(const SortGeom& g, int64_t ticket, int64_t& seg, int64_t& lt,
                                               int64_t& seg_start, int64_t& start, int& cnt) {
    const int64_t per_group = (int64_t)RS_GROUP * g.tiles_per_seg;
    const int64_t grp = ticket / per_group;
    const int64_t r = ticket - grp * per_group;
    int64_t sg = g.n_segs - grp * RS_GROUP;
    if (sg > RS_GROUP) sg = RS_GROUP;
    lt = r / sg;
    seg = grp * RS_GROUP + (r - lt * sg);
    seg_start = seg * g.seg_size;
    int64_t seg_end = seg_start + g.seg_size;
    if (seg_end > g.n) seg_end = g.n;
    start = seg_start + lt * RS_TILE;
    const int64_t c = seg_end - start;
    cnt = (int)(c > RS_TILE ? RS_TILE : c);
    return c > 0;
}

__global__ void __launch_bounds__(RS_THREADS)
k_hist(const uint64_t* __restrict__ keys, SortGeom g, const pch_voxel_plan* __restrict__ dplan, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * 256];
    const int tid = threadIdx.x;
    int bit_lo, n_bits;
    if (!pch_sort_range(g, dplan, bit_lo, n_bits)) return;
    const int n_passes = (n_bits + 7) >> 3;
    if (n_passes > g.hist_passes) return;
    for (int i = tid; i < RS_MAX_PASSES * 256; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    // canonical tile order here (no look-back): consecutive tiles of a CTA mostly share a segment, so the
    // table row is flushed once per segment change
    const int64_t tiles = g.total_tickets;
    const int64_t per = (tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * per, t1 = t0 + per < tiles ? t0 + per : tiles;
    int64_t cur_seg = -1;
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t seg = tile / g.tiles_per_seg, lt = tile - seg * g.tiles_per_seg;
        const int64_t seg_start = seg * g.seg_size;
        int64_t seg_end = seg_start + g.seg_size;
        if (seg_end > g.n) seg_end = g.n;
        const int64_t start = seg_start + lt * RS_TILE;
        if (start >= seg_end) continue;
        const int cnt = (int)(seg_end - start > RS_TILE ? RS_TILE : seg_end - start);
        if (seg != cur_seg) {
            if (cur_seg >= 0) {
                __syncthreads();
                pch_sort_hist_flush(sh, hist, cur_seg, g.hist_passes, n_bits);
                __syncthreads();
            }
            cur_seg = seg;
        }
#pragma unroll 4
        for (int i = tid; i < cnt; i += RS_THREADS) pch_sort_hist_add(sh, keys[start + i], bit_lo, n_bits);
    }
    __syncthreads();
    if (cur_seg >= 0) pch_sort_hist_flush(sh, hist, cur_seg, g.hist_passes, n_bits);
}

// exclusive scan of one value per thread over the whole block (threads beyond the 256 digits pass 0)
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*blockDim/32*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    __syncthreads();
    return base + x - v;
}

__global__ void __launch_bounds__(256) k_scan(uint32_t* __restrict__ hist, int64_t rows) {
    __shared__ uint32_t s_warp[8];
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        uint32_t v = hist[r * 256 + threadIdx.x];
        uint32_t e = block_excl_scan_256(v, s_warp);
        hist[r * 256 + threadIdx.x] = e;
    }
}

#define ST_AGG 0x40000000u
#define ST_INCL 0x80000000u
#define ST_VAL 0x3fffffffu

__global__ void __launch_bounds__(RS_THREADS, RS_MINB)
k_pass(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, SortGeom g, const pch_voxel_plan* __restrict__ dplan,
       int pass, const uint32_t* __restrict__ hist, uint32_t* __restrict__ status, uint32_t* __restrict__ counter,
       int* __restrict__ err) {
    __shared__ uint64_t s_keys[RS_TILE];
    __shared__ uint16_t s_whist[RS_WARPS][258];   // a warp ranks 32*RS_KPT = 512 keys: counts fit 16 bits; [256] = spare slot for out-of-range lanes
    __shared__ uint32_t s_dstart[256];
    __shared__ int64_t s_goff[256];
    __shared__ uint32_t s_scan[RS_THREADS / 32];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bit_lo, n_bits;
    if (!pch_sort_range(g, dplan, bit_lo, n_bits)) return;
    if (((n_bits + 7) >> 3) > g.hist_passes) {       // a device plan wider than the launched passes: flag, do not sort
        if (blockIdx.x == 0 && tid == 0) atomicExch(err, 2);
        return;
    }
    if (pass * 8 >= n_bits) return;                  // launched for the widest possible key; this key is narrower
    const int shift = bit_lo + 8 * pass;
    const uint32_t dmask = (1u << (n_bits - 8 * pass < 8 ? n_bits - 8 * pass : 8)) - 1u;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    for (int i = tid; i < RS_WARPS * 258 / 2; i += RS_THREADS) reinterpret_cast<uint32_t*>(&s_whist[0][0])[i] = 0;
    __syncthreads();
    int64_t seg, lt, seg_start, start;
    int cnt;
    if (!tile_of_ticket(g, (int64_t)s_tile, seg, lt, seg_start, start, cnt)) return;
    const int64_t tile = seg * g.tiles_per_seg + lt;          // canonical id: status row
    const int64_t first_tile = seg * g.tiles_per_seg;

    // ---- load (warp-blocked so that the in-tile order is the input order) and rank
    uint64_t key[RS_KPT];
    uint16_t rank[RS_KPT];
    const int wbase = warp * (32 * RS_KPT);
#pragma unroll
    for (int j = 0; j < RS_KPT; ++j) {
        int idx = wbase + j * 32 + lane;
        key[j] = idx < cnt ? in[start + idx] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < RS_KPT; ++j) {
        int idx = wbase + j * 32 + lane;
        uint32_t d = idx < cnt ? ((uint32_t)(key[j] >> shift) & dmask) : 256u;
#ifdef RS_BALLOT
        // peers = lanes holding the same digit, from one ballot per digit bit (+ one for the out-of-range
        // flag): MATCH.ANY has a long, poorly pipelined latency on sm_100, ballots issue back to back
        uint32_t peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
#else
        uint32_t peers = __match_any_sync(0xffffffffu, d);
#endif
        uint32_t before = __popc(peers & ((1u << lane) - 1u));
#ifdef RS_NOBRANCH
        // every lane reads its digit's running count (peers read the same word: a broadcast), then the last
        // peer alone stores the new count: no divergent region, no shuffle.  d == 256 uses the spare slot.
        uint32_t old = s_whist[warp][d];
        __syncwarp();
        if (before + 1 == __popc(peers)) s_whist[warp][d] = (uint16_t)(old + before + 1);
        rank[j] = (uint16_t)(old + before);
        __syncwarp();
#else
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && d < 256u) {
            old = s_whist[warp][d];
            s_whist[warp][d] = (uint16_t)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = (uint16_t)(old + before);
        __syncwarp();
#endif
    }
    __syncthreads();

    // ---- per digit (thread == digit): exclusive over warps, tile total; PUBLISH the aggregate now and
    // do the look-back only after the shared-memory scatter, so predecessors have time to publish too
    uint32_t my_sum = 0;
    if (tid < 256) {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = s_whist[w][d];
            s_whist[w][d] = (uint16_t)sum;
            sum += c;
        }
        my_sum = sum;
        pch_st_volatile_u32(status + ((size_t)tile * 256 + d), (tile == first_tile ? ST_INCL : ST_AGG) | sum);
    }
    {
        const uint32_t ds = block_excl_scan_256(my_sum, s_scan);
        if (tid < 256) s_dstart[tid] = ds;
    }
    __syncthreads();

    // ---- scatter inside shared memory (runs of equal digit become contiguous)
#pragma unroll
    for (int j = 0; j < RS_KPT; ++j) {
        int idx = wbase + j * 32 + lane;
        if (idx < cnt) {
            uint32_t d = (uint32_t)(key[j] >> shift) & dmask;
            s_keys[s_dstart[d] + s_whist[warp][d] + rank[j]] = key[j];
        }
    }

    // ---- look-back for digit `tid`
    if (tid < 256) {
        const int d = tid;
        uint32_t excl = 0;
        if (tile != first_tile) {
            // Walk back over the predecessors LB_WIN at a time: the window's status words are fetched
            // with independent loads (one L2 latency per window instead of one per tile), then consumed
            // in order; a word that is not published yet is re-polled.
            int64_t t = tile - 1;
            bool done = false;
            while (!done) {
                uint32_t w[LB_WIN];
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j) {
                    const int64_t tt = t - j;
                    w[j] = tt >= first_tile ? pch_ld_volatile_u32(status + ((size_t)tt * 256 + d)) : ST_INCL;
                }
#pragma unroll
                for (int j = 0; j < LB_WIN; ++j) {
                    if (done) break;
                    uint32_t spins = 0;
                    while ((w[j] & (ST_AGG | ST_INCL)) == 0) {
                        w[j] = pch_ld_volatile_u32(status + ((size_t)(t - j) * 256 + d));
                        if (++spins > PCH_SPIN_LIMIT) {
                            atomicExch(err, 1);
                            w[j] = ST_INCL;
                        }
                    }
                    excl += w[j] & ST_VAL;
                    if (w[j] & ST_INCL) done = true;
                }
                t -= LB_WIN;
            }
            pch_st_volatile_u32(status + ((size_t)tile * 256 + d), ST_INCL | ((excl + my_sum) & ST_VAL));
        }
        s_goff[d] = seg_start + (int64_t)hist[(seg * g.hist_passes + pass) * 256 + d] + (int64_t)excl - (int64_t)s_dstart[d];
    }
    __syncthreads();
    for (int i = tid; i < cnt; i += RS_THREADS) {
        uint64_t k = s_keys[i];
        uint32_t d = (uint32_t)(k >> shift) & dmask;
        out[s_goff[d] + i] = k;
    }
}

extern "C" size_t pch_sort_workspace_bytes(int64_t n, int64_t seg_size, int32_t bit_lo, int32_t bit_hi) {
    if (n <= 0 || bit_hi <= bit_lo) return 256;
    SortGeom g = pch_sort_geom(n, seg_size, bit_lo, bit_hi - bit_lo, (bit_hi - bit_lo + 7) / 8);
    return pch_sort_ws(g, nullptr).bytes;
}

int pch_sort_prepare(const SortGeom& g, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    SortWs w = pch_sort_ws(g, workspace);
    if (workspace_bytes < w.bytes) {
        pch_set_error("sort workspace too small: %zu < %zu", workspace_bytes, w.bytes);
        return PCH_ERR_WORKSPACE;
    }
    PCH_CHECK_ARG(g.seg_size < (1ll << 30), "segment size must be < 2^30");
    PCH_CHECK_ARG(g.hist_passes >= 1 && g.hist_passes <= RS_MAX_PASSES, "1..%d radix passes", RS_MAX_PASSES);
    PCH_CUDA(cudaMemsetAsync(workspace, 0, w.bytes, st));
    return PCH_OK;
}

int pch_sort_run(uint64_t* keys, uint64_t* tmp, const SortGeom& g, const pch_voxel_plan* dplan, int passes, bool prehist,
                 void* workspace, size_t workspace_bytes, cudaStream_t st) {
    SortWs w = pch_sort_ws(g, workspace);
    if (workspace_bytes < w.bytes) {
        pch_set_error("sort workspace too small: %zu < %zu", workspace_bytes, w.bytes);
        return PCH_ERR_WORKSPACE;
    }
    if (g.total_tickets == 0) return PCH_OK;
    if (!prehist) {
        int64_t hgrid = (int64_t)pch_sm_count() * 8;
        if (hgrid > g.total_tickets) hgrid = g.total_tickets;
        PCH_LAUNCH(st, "k_hist", k_hist<<<(unsigned)hgrid, RS_THREADS, 0, st>>>(keys, g, dplan, w.hist));
        PCH_LAUNCH_CHECK();
    }
    int64_t rows = g.n_segs * g.hist_passes;
    PCH_LAUNCH(st, "k_scan", k_scan<<<(unsigned)(rows < 4096 ? rows : 4096), 256, 0, st>>>(w.hist, rows));
    PCH_LAUNCH_CHECK();
    uint64_t* src = keys;
    uint64_t* dst = tmp;
    for (int p = 0; p < passes; ++p) {
        PCH_LAUNCH(st, "k_pass", k_pass<<<(unsigned)g.total_tickets, RS_THREADS, 0, st>>>(
                                     src, dst, g, dplan, p, w.hist, w.status + (size_t)p * g.total_tickets * 256, w.counters + p, w.err));
        PCH_LAUNCH_CHECK();
        uint64_t* t = src; src = dst; dst = t;
    }
    return PCH_OK;
}

extern "C" int pch_sort_u64_segmented(uint64_t* keys, uint64_t* tmp, int64_t n, int64_t seg_size, int32_t bit_lo,
                                      int32_t bit_hi, void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0, "n must be >= 0");
    PCH_CHECK_ARG(bit_lo >= 0 && bit_hi <= 64 && bit_hi >= bit_lo, "bad bit range [%d,%d)", bit_lo, bit_hi);
    if (n == 0 || bit_hi == bit_lo) return PCH_OK;
    PCH_CHECK_ARG(keys && tmp && workspace, "null pointer");
    const int passes = (bit_hi - bit_lo + 7) / 8;
    SortGeom g = pch_sort_geom(n, seg_size, bit_lo, bit_hi - bit_lo, passes);
    int rc = pch_sort_prepare(g, workspace, workspace_bytes, st);
    if (rc) return rc;
    return pch_sort_run(keys, tmp, g, nullptr, passes, false, workspace, workspace_bytes, st);
}
