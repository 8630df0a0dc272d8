// Raw-LAS-record tile streamer: persistent CTAs pull tiles of whole records into shared memory
// with 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) through a PCH_STAGES-deep mbarrier ring, so the
// AoS byte stream is read from HBM exactly once, in 16-byte units, regardless of the record length
// (20..67+ bytes, usually 34 = not a power of two and only 2-byte aligned).
#pragma once
#include "pch_common.cuh"

#define PCH_TILE_THREADS 256
#define PCH_STAGES 3

struct PchTileGeom {
    int64_t n;              // records
    int64_t chunk_size;     // tiles never straddle a chunk boundary
    int64_t tiles_per_chunk;
    int64_t total_tiles;
    int64_t total_bytes16;  // n*rec_len rounded up to 16
    int32_t rec_len;
    int32_t tile_records;   // multiple of PCH_TILE_THREADS
    int32_t stage_bytes;    // 16-byte multiple, >= tile_records*rec_len + 32
};

static inline PchTileGeom pch_tile_geom(int64_t n, int32_t rec_len, int64_t chunk_size) {
    PchTileGeom g;
    g.n = n;
    g.rec_len = rec_len;
    if (chunk_size <= 0 || chunk_size > n) chunk_size = n > 0 ? n : 1;
    g.chunk_size = chunk_size;
    int tr = (int)((60 * 1024) / rec_len / PCH_TILE_THREADS) * PCH_TILE_THREADS;
    if (tr > 1024) tr = 1024;
    if (tr < PCH_TILE_THREADS) tr = PCH_TILE_THREADS;  // rec_len > 240: still one record per thread
    g.tile_records = tr;
    g.tiles_per_chunk = pch_ceil_div(chunk_size, tr);
    int64_t n_chunks = pch_ceil_div(n, chunk_size);
    int64_t last = n - (n_chunks - 1) * chunk_size;
    g.total_tiles = n > 0 ? (n_chunks - 1) * g.tiles_per_chunk + pch_ceil_div(last, tr) : 0;
    g.total_bytes16 = (int64_t)pch_align_up((size_t)(n * rec_len), 16);
    g.stage_bytes = (int32_t)pch_align_up((size_t)tr * rec_len + 32, 128);
    return g;
}

static inline size_t pch_tile_smem_bytes(const PchTileGeom& g) { return 128 + (size_t)PCH_STAGES * g.stage_bytes; }

struct PchTile {
    int64_t chunk;      // chunk index
    int64_t r0;         // first record (global index)
    int32_t count;      // records in this tile
    const uint8_t* base;  // shared-memory address of record r0
};

__device__ __forceinline__ void pch_tile_range(const PchTileGeom& g, int64_t T, int64_t& chunk, int64_t& r0, int64_t& r1) {
    chunk = T / g.tiles_per_chunk;
    int64_t lt = T - chunk * g.tiles_per_chunk;
    r0 = chunk * g.chunk_size + lt * g.tile_records;
    int64_t cend = (chunk + 1) * g.chunk_size;
    if (cend > g.n) cend = g.n;
    r1 = r0 + g.tile_records;
    if (r1 > cend) r1 = cend;
}

// Calls f(tile) for every tile owned by this CTA (round-robin over the grid).  f is executed by all
// PCH_TILE_THREADS threads; it may use __syncthreads().  `smem` is the dynamic shared buffer.
template <class F>
__device__ __forceinline__ void pch_stream_tiles(const uint8_t* __restrict__ rec, const PchTileGeom& g, uint8_t* smem, F&& f) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* stages = smem + 128;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < PCH_STAGES; ++s) pch_mbar_init(&bars[s], 1);
        pch_fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int64_t T, int s) {
        int64_t chunk, r0, r1;
        pch_tile_range(g, T, chunk, r0, r1);
        int64_t b0 = r0 * g.rec_len, b1 = r1 * g.rec_len;
        int64_t a0 = b0 & ~(int64_t)15;
        int64_t a1 = (b1 + 15) & ~(int64_t)15;
        if (a1 > g.total_bytes16) a1 = g.total_bytes16;
        uint32_t bytes = (uint32_t)(a1 - a0);
        pch_mbar_arrive_expect_tx(&bars[s], bytes);
        pch_tma_load_1d(stages + (size_t)s * g.stage_bytes, rec + a0, bytes, &bars[s]);
    };
    const int64_t stride = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < PCH_STAGES; ++s) {
            int64_t T = (int64_t)blockIdx.x + s * stride;
            if (T < g.total_tiles) issue(T, s);
        }
    }
    int it = 0;
    for (int64_t T = blockIdx.x; T < g.total_tiles; T += stride, ++it) {
        const int s = it % PCH_STAGES;
        const uint32_t parity = (uint32_t)(it / PCH_STAGES) & 1u;
        pch_mbar_wait(&bars[s], parity);
        PchTile t;
        int64_t r1;
        pch_tile_range(g, T, t.chunk, t.r0, r1);
        t.count = (int32_t)(r1 - t.r0);
        t.base = stages + (size_t)s * g.stage_bytes + (int)((t.r0 * g.rec_len) & 15);
        f(t);
        __syncthreads();  // everyone is done reading stage s before it is refilled
        if (tid == 0) {
            int64_t Tn = T + (int64_t)PCH_STAGES * stride;
            if (Tn < g.total_tiles) issue(Tn, s);
        }
    }
}

// record alignment class of the byte stream (record i at byte i*rec_len from a 16-byte aligned base)
static inline int pch_rec_align(int32_t rec_len) { return (rec_len % 4 == 0) ? 4 : ((rec_len % 2 == 0) ? 2 : 1); }

static inline int pch_tile_grid(const PchTileGeom& g, int ctas_per_sm) {
    int64_t want = (int64_t)pch_sm_count() * ctas_per_sm;
    if (want > g.total_tiles) want = g.total_tiles;
    if (want < 1) want = 1;
    return (int)want;
}
