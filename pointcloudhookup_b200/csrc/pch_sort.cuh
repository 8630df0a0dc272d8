// Internal interface of the segmented radix sort (pch_sort.cu), shared with the kernels that PRODUCE
// keys (pch_voxel.cu, pch_dbscan.cu): a producer can build the per-segment digit histograms of every pass
// while it writes the keys (one read of the keys less), and the bit range that is sorted can live in
// DEVICE memory (a pch_voxel_plan written by the plan kernel), so no host round trip is needed between
// planning and sorting.
#pragma once
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
// segments whose tiles are handed out interleaved (ticket -> (tile-in-segment, segment)): the CTAs resident at
// any moment then work on RS_GROUP different segments, so a tile's predecessors in ITS segment finished long
// ago and the decoupled look-back finds an inclusive prefix one or two tiles back instead of walking through
// a whole wave of tiles that all published their aggregate at the same moment
#ifndef RS_GROUP
#define RS_GROUP 256
#endif

struct SortGeom {
    int64_t n, seg_size, tiles_per_seg, n_segs;
    int64_t total_tickets;      // n_segs * tiles_per_seg (tickets of the ragged last segment beyond its end are void)
    int32_t bit_lo, n_bits;     // host plan; ignored when a device plan is given
    int32_t hist_passes;        // rows per segment in the histogram table (= launched passes)
};

static inline SortGeom pch_sort_geom(int64_t n, int64_t seg_size, int32_t bit_lo, int32_t n_bits, int32_t hist_passes) {
    SortGeom g;
    g.n = n;
    if (seg_size <= 0 || seg_size > n) seg_size = n > 0 ? n : 1;
    g.seg_size = seg_size;
    g.tiles_per_seg = pch_ceil_div(seg_size, RS_TILE);
    g.n_segs = pch_ceil_div(n > 0 ? n : 1, seg_size);
    g.total_tickets = n > 0 ? g.n_segs * g.tiles_per_seg : 0;
    g.bit_lo = bit_lo;
    g.n_bits = n_bits;
    g.hist_passes = hist_passes;
    return g;
}

// workspace: [0,256): int err; uint32 counters[8] at +64 | hist: n_segs*hist_passes*256 u32 |
//            status: hist_passes*total_tickets*256 u32
struct SortWs {
    int* err;
    uint32_t* counters;
    uint32_t* hist;
    uint32_t* status;
    size_t hist_bytes;
    size_t bytes;
};
static inline SortWs pch_sort_ws(const SortGeom& g, void* base) {
    SortWs w;
    uint8_t* p = (uint8_t*)base;
    w.err = (int*)p;
    w.counters = (uint32_t*)(p + 64);
    size_t off = 256;
    w.hist = (uint32_t*)(p + off);
    w.hist_bytes = pch_align_up((size_t)g.n_segs * g.hist_passes * 256 * 4, 256);
    off += w.hist_bytes;
    w.status = (uint32_t*)(p + off);
    off += pch_align_up((size_t)g.hist_passes * g.total_tickets * 256 * 4, 256);
    w.bytes = off;
    return w;
}

// the bit range of this launch: from the device plan when there is one (status != 0 -> nothing to sort)
__device__ __forceinline__ bool pch_sort_range(const SortGeom& g, const pch_voxel_plan* __restrict__ dplan, int& bit_lo,
                                               int& n_bits) {
    bit_lo = g.bit_lo;
    n_bits = g.n_bits;
    if (dplan) {
        if (dplan->status != PCH_OK) return false;
        bit_lo = dplan->bits_idx;
        n_bits = dplan->key_bits;
    }
    return true;
}

// producer side: add one key to the shared-memory digit histograms sh[hist_passes][256]
__device__ __forceinline__ void pch_sort_hist_add(uint32_t* sh, uint64_t key, int bit_lo, int n_bits) {
    uint64_t k = key >> bit_lo;
    for (int b = 0; b < n_bits; b += 8) {
        const int w = n_bits - b < 8 ? n_bits - b : 8;
        atomicAdd(&sh[(b >> 3) * 256 + ((uint32_t)k & ((1u << w) - 1u))], 1u);
        k >>= 8;
    }
}
// producer side: flush a CTA's histograms into the table row of segment `seg` and clear them (all threads)
__device__ __forceinline__ void pch_sort_hist_flush(uint32_t* sh, uint32_t* __restrict__ hist, int64_t seg, int hist_passes,
                                                    int n_bits) {
    const int used = ((n_bits + 7) >> 3) * 256;
    for (int i = threadIdx.x; i < used; i += blockDim.x) {
        const uint32_t c = sh[i];
        if (c) {
            atomicAdd(&hist[(size_t)seg * hist_passes * 256 + i], c);
            sh[i] = 0;
        }
    }
}

// host side (pch_sort.cu): the passes of a sort whose histograms are already in the workspace (prehist) or not;
// `dplan` (device) overrides bit_lo / n_bits; `passes` kernels are launched, those beyond the planned range exit
int pch_sort_prepare(const SortGeom& g, void* workspace, size_t workspace_bytes, cudaStream_t st);   // zeroes histograms + status
int pch_sort_run(uint64_t* keys, uint64_t* tmp, const SortGeom& g, const pch_voxel_plan* dplan, int passes, bool prehist,
                 void* workspace, size_t workspace_bytes, cudaStream_t st);
