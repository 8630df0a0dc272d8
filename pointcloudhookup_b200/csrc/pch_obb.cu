// Minimum-volume oriented bounding box of every candidate cluster, on the device.
// Reference call site: utils/tower_extraction.py:137-139, `trimesh.PointCloud(cluster_points).bounding_box_oriented`
// -> `extents`, `transform`.  trimesh's search (convex hull; for every hull-face normal the minimum-area
// edge-aligned rectangle of the projected hull; smallest volume wins) is evaluated here for EVERY face normal —
// trimesh itself de-duplicates normals on a 0.1 rad grid and keeps the first one in Qhull's facet order, which no
// independent hull can reproduce and which only ever makes its box larger.
//
// One CTA per cluster, three steps, all in float64 on the float32 points:
//   1. hull candidates: S = the extreme points of 256 fixed directions; H1 = hull(S) by gift wrapping; every point
//      clearly inside H1 is interior to the hull and dropped (float32 plane tests with a safety margin), the rest
//      (typically 1-3 % of the cluster) + S are the candidates C.
//   2. hull(C) by gift wrapping: a stack of open directed edges, a hash set of the directed edges already owned by
//      a face, and per edge one block-wide scan for the point all others lie behind (orientation predicate;
//      coplanar ties go to the point farthest from the edge).  Faces (a,b,c) are counter-clockwise seen from outside.
//   3. box search: one warp per face normal n: thickness along n; every hull edge whose two faces look to opposite
//      sides of n is an edge of the projected hull: rectangle aligned with it, area * thickness -> block minimum.
#include "pch_common.cuh"
#include <stdlib.h>

#define OBB_THREADS 256
#define OBB_WARPS (OBB_THREADS / 32)
#define OBB_DIRS 256
#define OBB_MAXF 16384
#define OBB_MAXC 32768
#define OBB_HASH 65536
#define OBB_STACK (3 * OBB_MAXF)

struct ObbWs {
    int32_t* cand;      // [OBB_MAXC]
    int32_t* faces;     // [OBB_MAXF*3]
    double* planes;     // [OBB_MAXF*4] unit outward normal, offset (n.x = d on the plane)
    unsigned long long* eset;  // [OBB_HASH] directed edges (a << 32 | b) + 1, 0 = empty
    int32_t* face_of;   // [OBB_HASH] face that owns the directed edge of the same slot
    int4* stack;        // [OBB_STACK] open directed edges: a, b, vertex opposite to the edge in the face that is known
    int32_t* verts;     // [OBB_MAXC] hull vertices (indices into the cluster rows)
    uint32_t* mark;     // [OBB_MAXC/32+1] scratch bitmap
    int32_t* twin;      // [OBB_MAXF*3] face across directed edge j of face g
};
static const size_t OBB_WS_BYTES = pch_align_up((size_t)OBB_MAXC * 4, 256) + pch_align_up((size_t)OBB_MAXF * 12, 256) +
                                   pch_align_up((size_t)OBB_MAXF * 32, 256) + pch_align_up((size_t)OBB_HASH * 8, 256) +
                                   pch_align_up((size_t)OBB_HASH * 4, 256) + pch_align_up((size_t)OBB_STACK * 16, 256) +
                                   pch_align_up((size_t)OBB_MAXC * 4, 256) + pch_align_up((size_t)(OBB_MAXC / 32 + 1) * 4, 256) +
                                   pch_align_up((size_t)OBB_MAXF * 12, 256);

__host__ __device__ static inline ObbWs obb_ws(uint8_t* base) {
    ObbWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base + off; off += (bytes + 255) / 256 * 256; return p; };
    w.cand = (int32_t*)take((size_t)OBB_MAXC * 4);
    w.faces = (int32_t*)take((size_t)OBB_MAXF * 12);
    w.planes = (double*)take((size_t)OBB_MAXF * 32);
    w.eset = (unsigned long long*)take((size_t)OBB_HASH * 8);
    w.face_of = (int32_t*)take((size_t)OBB_HASH * 4);
    w.stack = (int4*)take((size_t)OBB_STACK * 16);
    w.verts = (int32_t*)take((size_t)OBB_MAXC * 4);
    w.mark = (uint32_t*)take((size_t)(OBB_MAXC / 32 + 1) * 4);
    w.twin = (int32_t*)take((size_t)OBB_MAXF * 12);
    return w;
}

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 ld3(const float* __restrict__ P, int i) { return D3{(double)P[i * 3 + 0], (double)P[i * 3 + 1], (double)P[i * 3 + 2]}; }
__device__ __forceinline__ D3 sub3(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// block-wide "best index" reduction with a caller-supplied comparator better(i, j): true when j beats i.
// -1 = no candidate.  All threads get the winner.
template <class BETTER>
__device__ __forceinline__ int block_best(int mine, BETTER better, int* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, mine, o);
        // both lanes of a pair must agree on the winner: order the pair so the comparison is evaluated identically
        const int lo = (lane & o) ? other : mine, hi = (lane & o) ? mine : other;
        int win;
        if (lo < 0) win = hi;
        else if (hi < 0) win = lo;
        else win = better(lo, hi) ? hi : lo;
        mine = win;
    }
    __syncthreads();
    if (lane == 0) s_red[warp] = mine;
    __syncthreads();
    int best = s_red[0];
    for (int w = 1; w < OBB_WARPS; ++w) {
        const int c = s_red[w];
        if (best < 0) best = c;
        else if (c >= 0 && better(best, c)) best = c;
    }
    __syncthreads();
    return best;
}

#define OBB_SMAX 256    // seed points of the cull (their hull has about 500 faces; every face is one serial wrap step)
#define OBB_ROUNDS 4
#define OBB_SPTS 2048   // candidate coordinates staged in shared memory when they fit (the usual case)
#define OBB_BIG_SPTS 16384   // second launch, one CTA per SM with 192 KB of dynamic shared memory: the few clusters that keep more

// The wrap works on LOCAL ids 0..n-1 (positions in the candidate list); row ids only come back at the end.
struct WrapCtx {
    const float* P;      // cluster rows
    const int32_t* idx;  // candidate list: local id -> row (NULL = identity, all rows)
    const float* S;      // coordinates of the candidates in shared memory, by local id (NULL = read the rows)
    int n;
    double tol3;         // coplanarity threshold on 6*volume
    double tol2;         // collinearity threshold on |cross|^2
};
__device__ __forceinline__ int row_of(const WrapCtx& c, int k) { return c.idx ? c.idx[k] : k; }
__device__ __forceinline__ D3 cpt(const WrapCtx& c, int k) {
    if (c.S) return D3{(double)c.S[k * 3 + 0], (double)c.S[k * 3 + 1], (double)c.S[k * 3 + 2]};
    return ld3(c.P, row_of(c, k));
}
// all threads: stage the candidates' coordinates when they fit
__device__ __forceinline__ void stage_candidates(WrapCtx& c, float* s_pts, int spts) {
    c.S = nullptr;
    if (c.n <= spts) {
        for (int k = threadIdx.x; k < c.n; k += OBB_THREADS) {
            const int r = row_of(c, k);
            s_pts[k * 3 + 0] = c.P[r * 3 + 0]; s_pts[k * 3 + 1] = c.P[r * 3 + 1]; s_pts[k * 3 + 2] = c.P[r * 3 + 2];
        }
        c.S = s_pts;
    }
    __syncthreads();
}

// The next face around the directed hull edge a->b: the point d with the largest rotation angle about the edge,
// measured from the half-plane of the face that is already known (O = a vector from a into that half-plane; for the
// very first edge a direction pointing out of the hull).  All points lie within half a turn of that half-plane, so
// "j is further round than i" is the sign of the triple product (e, I, J): a total order, evaluated pairwise.
// Coplanar ties: same half-plane -> the point farther from the edge line; opposite half-planes (angle 0 against half
// a turn) -> the one on the far side from O.
__device__ int wrap_edge(const WrapCtx& c, int a, int b, D3 O, int* s_red) {
    const D3 A = cpt(c, a), B = cpt(c, b);
    const D3 e = sub3(B, A);
    const double e2 = dot3(e, e);
    const D3 nO = cross3(e, O);
    auto better = [&](int i, int j) -> bool {   // does j beat i?
        const D3 I = sub3(cpt(c, i), A), J = sub3(cpt(c, j), A);
        const D3 nI = cross3(e, I);
        const double v = dot3(nI, J);                       // > 0: j is further round than i
        if (v > c.tol3) return true;
        if (v < -c.tol3) return false;
        const D3 nJ = cross3(e, J);
        if (dot3(nI, nJ) > 0.0) return dot3(nJ, nJ) > dot3(nI, nI);
        return dot3(nJ, nO) < dot3(nI, nO);
    };
    int mine = -1;
    for (int q = threadIdx.x; q < c.n; q += OBB_THREADS) {
        if (q == a || q == b) continue;
        const D3 Q = sub3(cpt(c, q), A);
        const D3 nq = cross3(e, Q);
        if (dot3(nq, nq) <= c.tol2 * e2) continue;          // on the edge line: cannot span a face with it
        if (mine < 0 || better(mine, q)) mine = q;
    }
    return block_best(mine, better, s_red);
}

__device__ __forceinline__ uint32_t edge_slot(unsigned long long key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 48) & (OBB_HASH - 1); }
__device__ bool eset_has(const ObbWs& w, int a, int b) {
    const unsigned long long key = (((unsigned long long)(uint32_t)a << 32) | (uint32_t)b) + 1ull;
    uint32_t s = edge_slot(key);
    for (int p = 0; p < OBB_HASH; ++p) {
        const unsigned long long v = w.eset[s];
        if (v == 0) return false;
        if (v == key) return true;
        s = (s + 1) & (OBB_HASH - 1);
    }
    return false;
}
__device__ void eset_add(const ObbWs& w, int a, int b, int face) {
    const unsigned long long key = (((unsigned long long)(uint32_t)a << 32) | (uint32_t)b) + 1ull;
    uint32_t s = edge_slot(key);
    for (int p = 0; p < OBB_HASH; ++p) {
        const unsigned long long v = w.eset[s];
        if (v == 0 || v == key) { w.eset[s] = key; w.face_of[s] = face; return; }
        s = (s + 1) & (OBB_HASH - 1);
    }
}
__device__ int eset_face(const ObbWs& w, int a, int b) {
    const unsigned long long key = (((unsigned long long)(uint32_t)a << 32) | (uint32_t)b) + 1ull;
    uint32_t s = edge_slot(key);
    for (int p = 0; p < OBB_HASH; ++p) {
        const unsigned long long v = w.eset[s];
        if (v == 0) return -1;
        if (v == key) return w.face_of[s];
        s = (s + 1) & (OBB_HASH - 1);
    }
    return -1;
}

// Gift wrapping of the candidate set.  Returns the number of faces (0 = degenerate input, -1 = capacity).
__device__ int gift_wrap(const WrapCtx& c, const ObbWs& w, int max_faces, int* s_red, int* s_ctl) {
    const int tid = threadIdx.x;
    for (int k = tid; k < OBB_HASH; k += OBB_THREADS) w.eset[k] = 0ull;
    // lowest point (x, then y, then z) and its neighbour on the silhouette of the xy projection
    auto lower = [&](int i, int j) -> bool {
        const D3 a = cpt(c, i), b = cpt(c, j);
        if (b.x != a.x) return b.x < a.x;
        if (b.y != a.y) return b.y < a.y;
        if (b.z != a.z) return b.z < a.z;
        return row_of(c, j) < row_of(c, i);
    };
    int mine = -1;
    for (int q = tid; q < c.n; q += OBB_THREADS)
        if (mine < 0 || lower(mine, q)) mine = q;
    const int p0 = block_best(mine, lower, s_red);
    if (p0 < 0) return 0;
    const D3 A = cpt(c, p0);
    auto right_of = [&](int i, int j) -> bool {   // j beats i when it lies to the right of p0->i in the xy projection
        const D3 I = sub3(cpt(c, i), A), J = sub3(cpt(c, j), A);
        const double cr = I.x * J.y - I.y * J.x;
        const double sc = sqrt((I.x * I.x + I.y * I.y) * (J.x * J.x + J.y * J.y));
        if (cr < -1e-12 * sc) return true;
        if (cr > 1e-12 * sc) return false;
        return (J.x * J.x + J.y * J.y) > (I.x * I.x + I.y * I.y);   // same direction: the farther one
    };
    mine = -1;
    for (int q = tid; q < c.n; q += OBB_THREADS) {
        if (q == p0) continue;
        const D3 Q = sub3(cpt(c, q), A);
        if (Q.x * Q.x + Q.y * Q.y <= c.tol2) continue;      // straight above / below p0
        if (mine < 0 || right_of(mine, q)) mine = q;
    }
    const int p1 = block_best(mine, right_of, s_red);
    if (p1 < 0) return 0;
    int n_faces = 0, sp = 0;
    if (tid == 0) {
        w.stack[0] = make_int4(p0, p1, -1, 0);     // -1: no face known yet; reference direction = out of the hull
        s_ctl[0] = 1;
    }
    __syncthreads();
    sp = 1;
    while (sp > 0) {
        // pop until an edge without a face turns up (thread 0 decides, everyone follows)
        if (tid == 0) {
            int a = -1, b = -1, op = -1, s = sp;
            while (s > 0) {
                const int4 e = w.stack[--s];
                if (!eset_has(w, e.x, e.y)) { a = e.x; b = e.y; op = e.z; break; }
            }
            s_ctl[0] = s; s_ctl[1] = a; s_ctl[2] = b; s_ctl[3] = op;
        }
        __syncthreads();
        sp = s_ctl[0];
        const int a = s_ctl[1], b = s_ctl[2], op = s_ctl[3];
        __syncthreads();
        if (a < 0) break;
        D3 O;
        if (op >= 0) O = sub3(cpt(c, op), cpt(c, a));
        else {
            // first edge: every point is to the left of a->b in the xy projection, so the right-hand side is outside
            const D3 e = sub3(cpt(c, b), cpt(c, a));
            O = D3{e.y, -e.x, 0.0};
        }
        const int d = wrap_edge(c, a, b, O, s_red);
        if (d < 0) {                       // everything collinear with this edge: degenerate
            if (n_faces == 0) return 0;
            if (tid == 0) eset_add(w, a, b, -1);
            __syncthreads();
            continue;
        }
        if (n_faces >= max_faces || sp + 3 >= OBB_STACK) return -1;
        if (tid == 0) {
            const int f = n_faces;
            w.faces[f * 3 + 0] = a; w.faces[f * 3 + 1] = b; w.faces[f * 3 + 2] = d;
            const D3 PA = cpt(c, a), PB = cpt(c, b), PD = cpt(c, d);
            D3 n = cross3(sub3(PB, PA), sub3(PD, PA));
            const double l = sqrt(dot3(n, n));
            if (l > 0.0) { n.x /= l; n.y /= l; n.z /= l; }
            w.planes[f * 4 + 0] = n.x; w.planes[f * 4 + 1] = n.y; w.planes[f * 4 + 2] = n.z;
            w.planes[f * 4 + 3] = dot3(n, PA);
            eset_add(w, a, b, f); eset_add(w, b, d, f); eset_add(w, d, a, f);
            int s = sp;
            if (!eset_has(w, b, a)) w.stack[s++] = make_int4(b, a, d, 0);
            if (!eset_has(w, d, b)) w.stack[s++] = make_int4(d, b, a, 0);
            if (!eset_has(w, a, d)) w.stack[s++] = make_int4(a, d, b, 0);
            s_ctl[0] = s;
        }
        __syncthreads();
        sp = s_ctl[0];
        ++n_faces;
        __syncthreads();
    }
    return n_faces;
}

struct ObbOut {   // mirrors pch_obb_result
    double extents[3];
    double center[3];
    double rot[9];      // columns = box axes in world coordinates (box -> world rotation, row-major)
    double volume;
    int32_t n_faces, n_verts, n_candidates, status;
};

__global__ void __launch_bounds__(OBB_THREADS)
k_obb(const float* __restrict__ points, const long long* __restrict__ ranges, int n_clusters, uint8_t* __restrict__ ws_base,
      size_t ws_stride, ObbOut* __restrict__ out, int max_cand, int spts, int pass) {
    __shared__ int s_red[OBB_WARPS];
    __shared__ int s_ctl[4];
    __shared__ float s_dirmax[OBB_DIRS];
    __shared__ int s_dirarg[OBB_DIRS];
    // one buffer (dynamic: spts * 12 bytes), two lives: the staged candidate coordinates during a wrap, the planes of
    // the seed hull during a cull
    extern __shared__ __align__(16) unsigned char s_buf[];
    float* s_pts = reinterpret_cast<float*>(s_buf);
    float4* s_plane = reinterpret_cast<float4*>(s_buf);
    __shared__ unsigned long long s_far[1024];
    __shared__ double s_best[OBB_WARPS];
    __shared__ int s_bestf[OBB_WARPS], s_beste[OBB_WARPS];
    __shared__ int s_count;
    const int k = blockIdx.x;
    if (k >= n_clusters) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = ranges[2 * k];
    const int n = (int)min((long long)INT_MAX / 4, ranges[2 * k + 1] - r0);
    const float* P = points + r0 * 3;
    ObbWs w = obb_ws(ws_base + (size_t)k * ws_stride);
    ObbOut* o = out + k;
    if (pass == 1 && o->status != 4) return;      // the second launch only takes the clusters the first one deferred
    auto fail = [&](int status) {
        if (tid == 0) { o->status = status; o->n_faces = 0; o->n_verts = 0; o->n_candidates = 0; o->volume = 0.0; }
    };
    if (n < 4) { fail(1); return; }

    // ---- extent -> tolerances
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += OBB_THREADS)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = P[i * 3 + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
    __shared__ uint32_t s_mm[6];
    if (tid < 6) s_mm[tid] = tid < 3 ? 0xffffffffu : 0u;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t lo = __reduce_min_sync(0xffffffffu, pch_f32_to_ordered(mn[a]));
        const uint32_t hi = __reduce_max_sync(0xffffffffu, pch_f32_to_ordered(mx[a]));
        if (lane == 0) { atomicMin(&s_mm[a], lo); atomicMax(&s_mm[3 + a], hi); }
    }
    __syncthreads();
    const double L = fmax(fmax((double)pch_ordered_to_f32(s_mm[3]) - (double)pch_ordered_to_f32(s_mm[0]),
                               (double)pch_ordered_to_f32(s_mm[4]) - (double)pch_ordered_to_f32(s_mm[1])),
                          (double)pch_ordered_to_f32(s_mm[5]) - (double)pch_ordered_to_f32(s_mm[2]));
    if (!(L > 0.0)) { fail(1); return; }
    WrapCtx c;
    c.P = P;
    c.S = nullptr;
    c.idx = nullptr;
    c.tol3 = 1e-11 * L * L * L;
    c.tol2 = 1e-18 * L * L;

    // ---- step 1: extreme points of n_dirs directions (Fibonacci sphere), hull of those, interior cull.  The hull of
    // the extreme points costs one serial wrap step per face, so small clusters use fewer directions (a coarser cull
    // is enough for them) and tiny ones skip the cull
    const int n_dirs = n <= 512 ? 0 : (n < 4096 ? 64 : (n < 16384 ? 128 : OBB_DIRS));
    for (int d = tid; d < OBB_DIRS; d += OBB_THREADS) { s_dirmax[d] = -INFINITY; s_dirarg[d] = -1; }
    __syncthreads();
    for (int d0 = 0; d0 < n_dirs; d0 += OBB_WARPS) {
        const int d = d0 + warp;
        const float zz = 1.0f - 2.0f * ((float)d + 0.5f) / (float)n_dirs;
        const float rr = sqrtf(fmaxf(0.f, 1.0f - zz * zz));
        const float ph = 2.399963229728653f * (float)d;
        const float dx = rr * cosf(ph), dy = rr * sinf(ph), dz = zz;
        float best = -INFINITY;
        int arg = -1;
        for (int i = lane; i < n; i += 32) {
            const float v = P[i * 3 + 0] * dx + P[i * 3 + 1] * dy + P[i * 3 + 2] * dz;
            if (v > best) { best = v; arg = i; }
        }
#pragma unroll
        for (int s = 16; s; s >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, s);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, s);
            if (ob > best || (ob == best && oa >= 0 && (arg < 0 || oa < arg))) { best = ob; arg = oa; }
        }
        if (lane == 0) { s_dirmax[d] = best; s_dirarg[d] = arg; }
    }
    __syncthreads();
    // S = distinct extreme points, in w.cand (direction order): every direction looks for an earlier direction
    // with the same extreme point in shared memory, the survivors are packed with ballots
    {
        __shared__ int s_wn[OBB_WARPS + 1];
        if (tid == 0) s_count = 0;
        __syncthreads();
        for (int d0 = 0; d0 < n_dirs; d0 += OBB_THREADS) {
            const int d = d0 + tid;
            bool keep = false;
            int a = -1;
            if (d < n_dirs) {
                a = s_dirarg[d];
                keep = a >= 0;
                for (int j = 0; j < d && keep; ++j) keep = s_dirarg[j] != a;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_wn[warp] = __popc(bal);
            __syncthreads();
            int base = s_count;
            for (int q = 0; q < warp; ++q) base += s_wn[q];
            if (keep) w.cand[base + __popc(bal & ((1u << lane) - 1u))] = a;
            __syncthreads();
            if (tid == 0) { int t = 0; for (int q = 0; q < OBB_WARPS; ++q) t += s_wn[q]; s_count += t; }
            __syncthreads();
        }
    }
    // The seed S (<= OBB_SMAX points) lives in w.cand[0..n_s); survivors ping-pong between listA / listB.
    // Rounds (quickhull in batches): H = hull(S); every point clearly inside H is interior and dropped; per face of H
    // the farthest point outside it joins S.  A thin sheet, which the first coarse H cannot thin at all, is down to its
    // hull neighbourhood after two or three rounds.
    int n_s = s_count;
    int32_t* listA = w.cand + OBB_SMAX;
    int32_t* listB = w.verts;                  // free until the hull vertices are collected
    const int list_cap = OBB_MAXC - OBB_SMAX;
    const int32_t* src = nullptr;              // nullptr = all rows
    int n_src = n, n_c = n;
    bool have_list = false;
    const float margin = (float)(1e-4 * L) + 1e-4f;
    for (int round = 0; round < OBB_ROUNDS && n_s >= 4; ++round) {
        c.idx = w.cand;
        c.n = n_s;
        stage_candidates(c, s_pts, spts);
        const int f1 = gift_wrap(c, w, 1024, s_red, s_ctl);
        __syncthreads();
        if (f1 <= 0) {
            // The hull of the extreme points did not close: the cluster is flat within the tolerances (f1 == 0), or so
            // nearly flat that the coplanar tie-breaks kept producing faces (f1 < 0).  Wrapping all points would only
            // repeat that at full cost: hand the cluster to the host, where Qhull decides (it raises for flat input).
            if (round == 0) { fail(f1 == 0 ? 2 : 3); return; }
            break;
        }
        for (int f = tid; f < f1; f += OBB_THREADS) {
            s_plane[f] = make_float4((float)w.planes[f * 4 + 0], (float)w.planes[f * 4 + 1], (float)w.planes[f * 4 + 2],
                                     (float)w.planes[f * 4 + 3]);
            s_far[f] = 0ull;
        }
        if (tid == 0) s_count = 0;
        __syncthreads();
        int32_t* dst = (src == listA) ? listB : listA;
        for (int k0 = 0; k0 < n_src; k0 += OBB_THREADS) {
            const int k = k0 + tid;
            bool keep = false;
            int i = -1;
            if (k < n_src) {
                i = src ? src[k] : k;
                const float x = P[i * 3 + 0], y = P[i * 3 + 1], z = P[i * 3 + 2];
                float dmax = -INFINITY;
                int fmax = 0;
                for (int f = 0; f < f1; ++f) {
                    const float4 pl = s_plane[f];
                    const float d = pl.x * x + pl.y * y + pl.z * z - pl.w;
                    if (d > dmax) { dmax = d; fmax = f; }
                }
                keep = dmax > -margin;                           // not clearly inside every face
                if (dmax > margin)                               // clearly outside: candidate for the next seed
                    atomicMax(&s_far[fmax], ((unsigned long long)__float_as_uint(dmax) << 32) | (uint32_t)i);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) {
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (pos < list_cap) dst[pos] = i;
            }
        }
        __syncthreads();
        const int kept = s_count;
        __syncthreads();
        if (kept <= list_cap) { src = dst; n_src = kept; n_c = kept; have_list = true; }
        // grow the seed by the farthest point of every face that still has something outside.  The same point can be the
        // farthest for several faces: every face looks for an earlier face with the same point in shared memory, the
        // distinct ones are appended with ballots (a serial version of this loop, reading the list from global memory,
        // was most of the kernel's time)
        if (tid == 0) s_ctl[0] = n_s;
        __syncthreads();
        for (int f0 = 0; f0 < f1; f0 += OBB_THREADS) {
            const int f = f0 + tid;
            bool add = false;
            int i = -1;
            if (f < f1 && s_far[f]) {
                i = (int)(uint32_t)s_far[f];
                add = true;
                for (int g = 0; g < f && add; ++g) add = (int)(uint32_t)s_far[g] != i || !s_far[g];
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, add);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_ctl[0], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (add) {
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                if (pos < OBB_SMAX) w.cand[pos] = i;
            }
        }
        __syncthreads();
        if (tid == 0 && s_ctl[0] > OBB_SMAX) s_ctl[0] = OBB_SMAX;
        __syncthreads();
        const int grown = s_ctl[0];
        __syncthreads();
        const bool done = grown == n_s || (have_list && n_c <= OBB_SPTS);     // small enough for the fast path of either launch
        n_s = grown;
        if (done) break;
    }
    if (have_list) {
        if (src == listB) {       // w.verts is needed for the hull vertices later: move the survivors to the other list
            for (int k = tid; k < n_c; k += OBB_THREADS) listA[k] = listB[k];
            __syncthreads();
            src = listA;
        }
        c.idx = src;
    } else {      // no cull: a small cluster (no seed rounds) is wrapped whole; a big one whose survivors never fitted the
                  // lists (tens of thousands of points all near the hull) costs Qhull far less than gift wrapping
        if (n > 4096) { fail(3); return; }
        c.idx = nullptr;
        n_c = n;
    }
    __syncthreads();

    // ---- step 2: hull of the candidates.  Gift wrapping costs (candidates x faces); a cluster the cull cannot thin (a
    // thin convex sheet: every point is a hull vertex) is cheaper in Qhull on the host than here
    if (n_c > max_cand) {
        if (tid == 0) { o->status = 3; o->n_faces = 0; o->n_verts = 0; o->n_candidates = n_c; o->volume = 0.0; }
        return;
    }
    if (pass == 0 && n_c > spts && n_c <= OBB_BIG_SPTS) {
        // more candidates than this launch can stage: wrapping them from global memory costs tens of milliseconds per
        // cluster; the second launch (large shared memory, one CTA per SM) takes it
        if (tid == 0) o->status = 4;
        return;
    }
    c.n = n_c;
    stage_candidates(c, s_pts, spts);
    const int F = gift_wrap(c, w, OBB_MAXF, s_red, s_ctl);
    __syncthreads();
    if (F <= 0) { fail(F == 0 ? 2 : 3); return; }
    // faces across every directed edge (the edge set speaks local ids), then the faces go back to row ids
    for (int i = tid; i < 3 * F; i += OBB_THREADS) {
        const int g = i / 3, j = i - 3 * g;
        w.twin[i] = eset_face(w, w.faces[g * 3 + (j + 1) % 3], w.faces[g * 3 + j]);
    }
    __syncthreads();
    if (c.idx)
        for (int i = tid; i < 3 * F; i += OBB_THREADS) w.faces[i] = c.idx[w.faces[i]];
    __syncthreads();
    // hull vertices
    if (tid == 0) s_count = 0;
    for (int i = tid; i < OBB_MAXC / 32 + 1; i += OBB_THREADS) w.mark[i] = 0u;
    __syncthreads();
    // vertex indices can be any row of the cluster (up to n): dedupe through the edge set instead of a bitmap:
    // a vertex v is recorded by the face that owns a directed edge starting at v with the smallest slot: simpler,
    // thread 0 walks the faces and keeps a small open-addressing set in w.stack (free now)
    if (tid == 0) {
        int m = 0;
        int* set = (int*)w.stack;                    // OBB_STACK*4 ints, all free after the wrap
        const int cap = 65536;                       // < OBB_STACK*4
        for (int i = 0; i < cap; ++i) set[i] = -1;
        for (int f = 0; f < F; ++f)
            for (int j = 0; j < 3; ++j) {
                const int v = w.faces[f * 3 + j];
                uint32_t s = ((uint32_t)v * 2654435761u) >> 16;
                bool found = false;
                while (set[s] >= 0) { if (set[s] == v) { found = true; break; } s = (s + 1) & (cap - 1); }
                if (!found) { set[s] = v; if (m < OBB_MAXC) w.verts[m] = v; ++m; }
            }
        s_count = m;
    }
    __syncthreads();
    const int V = s_count;
    if (V > OBB_MAXC) { fail(3); return; }

    // ---- step 3: box search.  One warp per face normal.
    double best_vol = INFINITY;
    int best_f = -1, best_e = -1;
    for (int f = warp; f < F; f += OBB_WARPS) {
        const D3 nrm{w.planes[f * 4 + 0], w.planes[f * 4 + 1], w.planes[f * 4 + 2]};
        if (!(dot3(nrm, nrm) > 0.5)) continue;                       // degenerate triangle
        double lo = INFINITY, hi = -INFINITY;
        for (int v = lane; v < V; v += 32) {
            const double t = dot3(nrm, ld3(P, w.verts[v]));
            lo = fmin(lo, t); hi = fmax(hi, t);
        }
#pragma unroll
        for (int s = 16; s; s >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, s)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, s)); }
        const double thick = hi - lo;
        // silhouette edges: a directed edge of a face that looks along n (n.n_g >= 0) whose twin's face looks away.
        // Lanes test 32 edges at a time; each hit is then evaluated by the whole warp.
        for (int i0 = 0; i0 < 3 * F; i0 += 32) {
            const int i = i0 + lane;
            bool sil = false;
            if (i < 3 * F) {
                const int g = i / 3, tf = w.twin[i];
                if (tf >= 0) {
                    const double sg = nrm.x * w.planes[g * 4 + 0] + nrm.y * w.planes[g * 4 + 1] + nrm.z * w.planes[g * 4 + 2];
                    const double st = nrm.x * w.planes[tf * 4 + 0] + nrm.y * w.planes[tf * 4 + 1] + nrm.z * w.planes[tf * 4 + 2];
                    sil = sg >= 0.0 && st < 0.0;
                }
            }
            uint32_t hits = __ballot_sync(0xffffffffu, sil);
            while (hits) {
                const int bit = __ffs(hits) - 1;
                hits &= hits - 1;
                const int ei = i0 + bit, g = ei / 3, j = ei - 3 * g;
                const int a = w.faces[g * 3 + j], b = w.faces[g * 3 + (j + 1) % 3];
                const D3 e = sub3(ld3(P, b), ld3(P, a));
                const double en = dot3(e, nrm);
                D3 u{e.x - en * nrm.x, e.y - en * nrm.y, e.z - en * nrm.z};
                const double ul = sqrt(dot3(u, u));
                if (!(ul > 1e-10)) continue;
                u.x /= ul; u.y /= ul; u.z /= ul;
                const D3 vv = cross3(nrm, u);
                double ulo = INFINITY, uhi = -INFINITY, vlo = INFINITY, vhi = -INFINITY;
                for (int v = lane; v < V; v += 32) {
                    const D3 p = ld3(P, w.verts[v]);
                    const double pu = dot3(u, p), pv = dot3(vv, p);
                    ulo = fmin(ulo, pu); uhi = fmax(uhi, pu); vlo = fmin(vlo, pv); vhi = fmax(vhi, pv);
                }
#pragma unroll
                for (int s = 16; s; s >>= 1) {
                    ulo = fmin(ulo, __shfl_xor_sync(0xffffffffu, ulo, s)); uhi = fmax(uhi, __shfl_xor_sync(0xffffffffu, uhi, s));
                    vlo = fmin(vlo, __shfl_xor_sync(0xffffffffu, vlo, s)); vhi = fmax(vhi, __shfl_xor_sync(0xffffffffu, vhi, s));
                }
                const double vol = (uhi - ulo) * (vhi - vlo) * thick;
                if (vol < best_vol || (vol == best_vol && (f < best_f || (f == best_f && ei < best_e)))) {
                    best_vol = vol; best_f = f; best_e = ei;
                }
            }
        }
    }
    if (lane == 0) { s_best[warp] = best_vol; s_bestf[warp] = best_f; s_beste[warp] = best_e; }
    __syncthreads();
    if (warp == 0) {
        double bv = INFINITY; int bf = -1, be = -1;
        for (int q = 0; q < OBB_WARPS; ++q)
            if (s_bestf[q] >= 0 && (s_best[q] < bv || (s_best[q] == bv && (s_bestf[q] < bf || (s_bestf[q] == bf && s_beste[q] < be))))) {
                bv = s_best[q]; bf = s_bestf[q]; be = s_beste[q];
            }
        // no candidate frame, or a box without volume (all points in one plane: Qhull / trimesh raise for such input)
        if (bf < 0 || !(bv > 1e-12 * L * L * L)) { if (lane == 0) { o->status = 2; o->n_faces = F; o->n_verts = V; o->n_candidates = n_c; o->volume = 0.0; } return; }
        // rebuild the winning frame and write the result
        const D3 nrm{w.planes[bf * 4 + 0], w.planes[bf * 4 + 1], w.planes[bf * 4 + 2]};
        const int g = be / 3, j = be % 3;
        const int a = w.faces[g * 3 + j], b = w.faces[g * 3 + (j + 1) % 3];
        const D3 e = sub3(ld3(P, b), ld3(P, a));
        const double en = dot3(e, nrm);
        D3 u{e.x - en * nrm.x, e.y - en * nrm.y, e.z - en * nrm.z};
        const double ul = sqrt(dot3(u, u));
        u.x /= ul; u.y /= ul; u.z /= ul;
        const D3 vv = cross3(nrm, u);
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int v = lane; v < V; v += 32) {
            const D3 p = ld3(P, w.verts[v]);
            const double t[3] = {dot3(u, p), dot3(vv, p), dot3(nrm, p)};
#pragma unroll
            for (int q = 0; q < 3; ++q) { lo[q] = fmin(lo[q], t[q]); hi[q] = fmax(hi[q], t[q]); }
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int s = 16; s; s >>= 1) { lo[q] = fmin(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], s)); hi[q] = fmax(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], s)); }
        if (lane == 0) {
            // trimesh convention: extents = (rect long, rect short, along-normal); box axes = (long, short, normal)
            double ex[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
            double ce[3] = {0.5 * (hi[0] + lo[0]), 0.5 * (hi[1] + lo[1]), 0.5 * (hi[2] + lo[2])};
            D3 ax0 = u, ax1 = vv;
            if (ex[0] < ex[1]) {      // rotate the rectangle by 90 degrees about the normal: (u, v) -> (v, -u)
                const double t = ex[0]; ex[0] = ex[1]; ex[1] = t;
                const double tc = ce[0]; ce[0] = ce[1]; ce[1] = -tc;
                ax0 = vv; ax1 = D3{-u.x, -u.y, -u.z};
            }
            o->extents[0] = ex[0]; o->extents[1] = ex[1]; o->extents[2] = ex[2];
            o->center[0] = ce[0] * ax0.x + ce[1] * ax1.x + ce[2] * nrm.x;
            o->center[1] = ce[0] * ax0.y + ce[1] * ax1.y + ce[2] * nrm.y;
            o->center[2] = ce[0] * ax0.z + ce[1] * ax1.z + ce[2] * nrm.z;
            o->rot[0] = ax0.x; o->rot[1] = ax1.x; o->rot[2] = nrm.x;
            o->rot[3] = ax0.y; o->rot[4] = ax1.y; o->rot[5] = nrm.y;
            o->rot[6] = ax0.z; o->rot[7] = ax1.z; o->rot[8] = nrm.z;
            o->volume = bv;
            o->n_faces = F; o->n_verts = V; o->n_candidates = n_c; o->status = 0;
        }
    }
}

extern "C" size_t pch_obb_workspace_bytes(int32_t n_clusters) { return (size_t)(n_clusters > 0 ? n_clusters : 1) * OBB_WS_BYTES; }

static int obb_max_candidates() {      // PCH_OBB_MAX_CAND: tuning knob (clusters with more hull candidates go to the host)
    const char* e = getenv("PCH_OBB_MAX_CAND");
    int v = e ? atoi(e) : OBB_MAXC;
    return v < 16 ? 16 : v;
}

extern "C" int pch_obb_batch(const float* points_dev, const int64_t* ranges_dev, int32_t n_clusters, pch_obb_result* out_dev,
                             void* workspace, size_t workspace_bytes, pch_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    static_assert(sizeof(ObbOut) == sizeof(pch_obb_result), "pch_obb_result layout");
    PCH_CHECK_ARG(n_clusters >= 0, "n_clusters must be >= 0");
    if (n_clusters == 0) return PCH_OK;
    PCH_CHECK_ARG(points_dev && ranges_dev && out_dev && workspace, "null pointer");
    if (workspace_bytes < pch_obb_workspace_bytes(n_clusters)) {
        pch_set_error("obb workspace too small: %zu < %zu", workspace_bytes, pch_obb_workspace_bytes(n_clusters));
        return PCH_ERR_WORKSPACE;
    }
    const int max_cand = obb_max_candidates();
    PCH_LAUNCH(st, "k_obb", k_obb<<<(unsigned)n_clusters, OBB_THREADS, (size_t)OBB_SPTS * 12, st>>>(
                                points_dev, (const long long*)ranges_dev, n_clusters, (uint8_t*)workspace, OBB_WS_BYTES,
                                (ObbOut*)out_dev, max_cand, OBB_SPTS, 0));
    PCH_LAUNCH_CHECK();
    // second launch for the clusters the first one deferred (status 4): same kernel, 192 KB of dynamic shared memory
    PCH_CUDA(cudaFuncSetAttribute(k_obb, cudaFuncAttributeMaxDynamicSharedMemorySize, OBB_BIG_SPTS * 12));
    PCH_LAUNCH(st, "k_obb_big", k_obb<<<(unsigned)n_clusters, OBB_THREADS, (size_t)OBB_BIG_SPTS * 12, st>>>(
                                    points_dev, (const long long*)ranges_dev, n_clusters, (uint8_t*)workspace, OBB_WS_BYTES,
                                    (ObbOut*)out_dev, max_cand, OBB_BIG_SPTS, 1));
    PCH_LAUNCH_CHECK();
    return PCH_OK;
}
