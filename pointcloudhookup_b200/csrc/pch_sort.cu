// Segmented stable LSD radix sort of 64-bit keys (8-bit digits), one-sweep style:
//   k_hist  : ONE read of the keys builds the digit histograms of every pass, per segment
//   k_scan  : exclusive scan of each 256-bin histogram -> digit bases inside the segment
//   k_pass  : per pass, one kernel: tile-local ranking (warp match_any multi-split), chained
//             decoupled look-back across the tiles of a segment, scatter via shared memory.
// Each pass therefore moves 8 B in + 8 B out per key; only the bits that carry voxel/cell index
// are sorted, and segments (voxel chunks / DBSCAN chunks) never mix because every offset is
// segment-relative.  Used for open3d's voxel grouping (ui/import_PC.py:12) and for the DBSCAN
// cell grid (utils/tower_extraction.py:107-112).
#include "pch_common.cuh"

#define RS_THREADS 256
#ifndef RS_KPT
#define RS_KPT 16
#endif
#ifndef RS_MINB
#define RS_MINB 4
#endif
#define RS_TILE (RS_THREADS * RS_KPT)
#define RS_WARPS (RS_THREADS / 32)
#define RS_MAX_PASSES 8
#define LB_WIN 8

struct SortGeom {
    int64_t n, seg_size, tiles_per_seg, total_tiles, n_segs;
    int32_t bit_lo, n_passes;
    int32_t pass_bits[RS_MAX_PASSES];
};

static SortGeom sort_geom(int64_t n, int64_t seg_size, int32_t bit_lo, int32_t bit_hi) {
    SortGeom g;
    g.n = n;
    if (seg_size <= 0 || seg_size > n) seg_size = n > 0 ? n : 1;
    g.seg_size = seg_size;
    g.tiles_per_seg = pch_ceil_div(seg_size, RS_TILE);
    g.n_segs = pch_ceil_div(n, seg_size);
    int64_t last = n - (g.n_segs - 1) * seg_size;
    g.total_tiles = n > 0 ? (g.n_segs - 1) * g.tiles_per_seg + pch_ceil_div(last, RS_TILE) : 0;
    g.bit_lo = bit_lo;
    int bits = bit_hi - bit_lo;
    g.n_passes = (bits + 7) / 8;
    for (int p = 0; p < RS_MAX_PASSES; ++p) {
        int b = bits - 8 * p;
        g.pass_bits[p] = b >= 8 ? 8 : (b > 0 ? b : 0);
    }
    return g;
}

// workspace: [0,256): int err; uint32 counters[8] at +64 | hist: n_segs*passes*256 u32 | status: passes*tiles*256 u32
struct SortWs {
    int* err;
    uint32_t* counters;
    uint32_t* hist;
    uint32_t* status;
    size_t bytes;
    size_t zero_bytes;  // everything is zeroed in one memset
};
static SortWs sort_ws(const SortGeom& g, void* base) {
    SortWs w;
    uint8_t* p = (uint8_t*)base;
    w.err = (int*)p;
    w.counters = (uint32_t*)(p + 64);
    size_t off = 256;
    w.hist = (uint32_t*)(p + off);
    off += pch_align_up((size_t)g.n_segs * g.n_passes * 256 * 4, 256);
    w.status = (uint32_t*)(p + off);
    off += pch_align_up((size_t)g.n_passes * g.total_tiles * 256 * 4, 256);
    w.bytes = off;
    w.zero_bytes = off;
    return w;
}

__device__ __forceinline__ void tile_span(const SortGeom& g, int64_t tile, int64_t& seg, int64_t& seg_start,
                                          int64_t& start, int& cnt) {
    seg = tile / g.tiles_per_seg;
    int64_t lt = tile - seg * g.tiles_per_seg;
    seg_start = seg * g.seg_size;
    int64_t seg_end = seg_start + g.seg_size;
    if (seg_end > g.n) seg_end = g.n;
    start = seg_start + lt * RS_TILE;
    int64_t c = seg_end - start;
    cnt = (int)(c > RS_TILE ? RS_TILE : c);
}

__global__ void __launch_bounds__(RS_THREADS)
k_hist(const uint64_t* __restrict__ keys, SortGeom g, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_MAX_PASSES * 256];
    const int tid = threadIdx.x;
    for (int64_t tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int64_t seg, seg_start, start;
        int cnt;
        tile_span(g, tile, seg, seg_start, start, cnt);
        for (int i = tid; i < g.n_passes * 256; i += RS_THREADS) sh[i] = 0;
        __syncthreads();
#pragma unroll 4
        for (int i = tid; i < cnt; i += RS_THREADS) {
            uint64_t k = keys[start + i] >> g.bit_lo;
            for (int p = 0; p < g.n_passes; ++p) {
                uint32_t d = (uint32_t)(k >> (8 * p)) & ((1u << g.pass_bits[p]) - 1u);
                atomicAdd(&sh[p * 256 + d], 1u);
            }
        }
        __syncthreads();
        for (int i = tid; i < g.n_passes * 256; i += RS_THREADS) {
            uint32_t c = sh[i];
            if (c) atomicAdd(&hist[(seg * g.n_passes) * 256 + i], c);
        }
        __syncthreads();
    }
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
k_pass(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, SortGeom g, int pass, int shift, uint32_t dmask,
       const uint32_t* __restrict__ hist, uint32_t* __restrict__ status, uint32_t* __restrict__ counter,
       int* __restrict__ err) {
    __shared__ uint64_t s_keys[RS_TILE];
    __shared__ uint16_t s_whist[RS_WARPS][258];   // a warp ranks 32*RS_KPT = 512 keys: counts fit 16 bits; [256] = spare slot for out-of-range lanes
    __shared__ uint32_t s_dstart[256];
    __shared__ int64_t s_goff[256];
    __shared__ uint32_t s_scan[RS_THREADS / 32];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    for (int i = tid; i < RS_WARPS * 258 / 2; i += RS_THREADS) reinterpret_cast<uint32_t*>(&s_whist[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    if (tile >= g.total_tiles) return;
    int64_t seg, seg_start, start;
    int cnt;
    tile_span(g, tile, seg, seg_start, start, cnt);
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
        s_goff[d] = seg_start + (int64_t)hist[(seg * g.n_passes + pass) * 256 + d] + (int64_t)excl - (int64_t)s_dstart[d];
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
    SortGeom g = sort_geom(n, seg_size, bit_lo, bit_hi);
    return sort_ws(g, nullptr).bytes;
}

extern "C" int pch_sort_u64_segmented(uint64_t* keys, uint64_t* tmp, int64_t n, int64_t seg_size, int32_t bit_lo,
                                      int32_t bit_hi, void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    PCH_CHECK_ARG(n >= 0, "n must be >= 0");
    PCH_CHECK_ARG(bit_lo >= 0 && bit_hi <= 64 && bit_hi >= bit_lo, "bad bit range [%d,%d)", bit_lo, bit_hi);
    if (n == 0 || bit_hi == bit_lo) return PCH_OK;
    PCH_CHECK_ARG(keys && tmp && workspace, "null pointer");
    SortGeom g = sort_geom(n, seg_size, bit_lo, bit_hi);
    PCH_CHECK_ARG(g.seg_size < (1ll << 30), "segment size must be < 2^30");
    SortWs w = sort_ws(g, workspace);
    if (workspace_bytes < w.bytes) {
        pch_set_error("sort workspace too small: %zu < %zu", workspace_bytes, w.bytes);
        return PCH_ERR_WORKSPACE;
    }
    PCH_CUDA(cudaMemsetAsync(workspace, 0, w.zero_bytes, st));
    int64_t hgrid = (int64_t)pch_sm_count() * 8;
    if (hgrid > g.total_tiles) hgrid = g.total_tiles;
    PCH_LAUNCH(st, "k_hist", k_hist<<<(unsigned)hgrid, RS_THREADS, 0, st>>>(keys, g, w.hist));
    PCH_LAUNCH_CHECK();
    int64_t rows = g.n_segs * g.n_passes;
    PCH_LAUNCH(st, "k_scan", k_scan<<<(unsigned)(rows < 4096 ? rows : 4096), 256, 0, st>>>(w.hist, rows));
    PCH_LAUNCH_CHECK();
    uint64_t* src = keys;
    uint64_t* dst = tmp;
    for (int p = 0; p < g.n_passes; ++p) {
        PCH_LAUNCH(st, "k_pass", k_pass<<<(unsigned)g.total_tiles, RS_THREADS, 0, st>>>(src, dst, g, p, g.bit_lo + 8 * p, (1u << g.pass_bits[p]) - 1u, w.hist,
                                                               w.status + (size_t)p * g.total_tiles * 256,
                                                               w.counters + p, w.err));
        PCH_LAUNCH_CHECK();
        uint64_t* t = src; src = dst; dst = t;
    }
    return PCH_OK;
}
