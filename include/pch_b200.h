/* pch_b200.h — C ABI of libpch_b200.so: the B200 (sm_100a) kernels behind pointcloudhookup's
 * per-point LAS hot path.
 *
 * The reference (Daniel-Starr/pointcloudhookup) is pure Python and has no FFI of its own; the
 * boundary it exposes is the Python module surface imported by pyGUI_towers_test.py:16-26.  Each
 * entry point below names the reference call (file:line, relative to the reference root) whose
 * per-point arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is DEVICE memory owned by the caller, 16-byte
 *     aligned; `const double* scales/offsets` are HOST pointers to 3 doubles (LAS header values).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing,
 *     never synchronises; scratch comes from the caller via *_workspace_bytes().
 *   - return 0 on success, <0 on error; pch_last_error() gives a thread-local message.
 *   - raw LAS point data (`rec_dev`) starts at record 0 (file offset offset_to_point_data), record
 *     i at byte i*rec_len, and must be readable up to the next multiple of 16 bytes past the end
 *     (bulk-TMA tiles are fetched in whole 16-byte units).
 */
#ifndef PCH_B200_H
#define PCH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCH_OK 0
#define PCH_ERR_INVALID (-1)   /* bad argument */
#define PCH_ERR_CUDA (-2)      /* CUDA runtime error (message has the call) */
#define PCH_ERR_RANGE (-3)     /* voxel index range does not fit the packed 64-bit key */
#define PCH_ERR_WORKSPACE (-4) /* workspace too small */

typedef void* pch_stream_t;

const char* pch_last_error(void);
int pch_version(void);
/* kernels launched by this library so far (process-wide; bench.py reports the per-step delta) */
long long pch_launch_count(void);
/* per-kernel device timing without a profiler: while enabled every launch is bracketed by CUDA
 * events on its stream; pch_profile_report synchronises the device and writes one
 * "kernel_name launches total_ms" line per kernel into buf, then clears the records. */
void pch_profile_enable(int on);
int pch_profile_report(char* buf, size_t cap);

/* ---------------------------------------------------------------- LAS attribute decode / encode */

/* laspy.read(...); per chunk of `chunk_size` consecutive records, the int32 lattice extrema
 * (ui/import_PC.py:45-48 slices las.points[start:end]; open3d takes the chunk's min bound,
 * ui/import_PC.py:12).  minmax_dev: [n_chunks][6] = minX,minY,minZ,maxX,maxY,maxZ. */
int pch_las_chunk_minmax(const uint8_t* rec_dev, int64_t n, int32_t rec_len, int64_t chunk_size,
                         int32_t* minmax_dev,
                         int32_t* xyz16_dev /* nullable: also emit the (n,4) int32 X,Y,Z,0 copy in this pass */,
                         pch_stream_t stream);

/* np.vstack((las.x, las.y, las.z)).T -> (n,3) float64, x = X*scale+offset
 * (ui/import_PC.py:47-48; ui/extract.py:114-115,361-362). */
int pch_las_decode_f64(const uint8_t* rec_dev, int64_t n, int32_t rec_len, const double* scales,
                       const double* offsets, double* xyz_dev, pch_stream_t stream);

/* np.stack([las.x, las.y, las.z], axis=1).astype(np.float32) (utils/tower_extraction.py:60-62). */
int pch_las_decode_f32(const uint8_t* rec_dev, int64_t n, int32_t rec_len, const double* scales,
                       const double* offsets, float* xyz_dev, pch_stream_t stream);

/* `las.x = arr` (ui/import_PC.py:61-63; utils/tower_extraction.py:253-255):
 * X = round_half_even((arr - offset)/scale) -> int32; (m,3) float64 -> (m,3) int32. */
int pch_las_quantise(const double* xyz_dev, int64_t m, const double* scales, const double* offsets,
                     int32_t* lattice_dev, pch_stream_t stream);

/* LasData.write (ui/import_PC.py:65): zero records of `rec_len` bytes with X,Y,Z filled in, plus the
 * lattice extrema for the header (minmax6_dev: minX,minY,minZ,maxX,maxY,maxZ). */
int pch_las_encode(const int32_t* lattice_dev, int64_t m, int32_t rec_len, uint8_t* rec_out_dev,
                   int32_t* minmax6_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- voxel-grid downsample */

typedef struct pch_voxel_plan {
    int32_t bits_x, bits_y, bits_z; /* bits per voxel-index axis (max over chunks)              */
    int32_t bits_idx;               /* bits of the in-chunk point index packed below the key     */
    int32_t key_bits;               /* bits_x + bits_y + bits_z                                   */
    int32_t n_passes;               /* 8-bit radix passes the sort needs = ceil(key_bits / 8)     */
    int32_t status;                 /* PCH_OK or PCH_ERR_RANGE                                    */
    int32_t reserved;
} pch_voxel_plan;

/* open3d voxel_down_sample set-up (ui/import_PC.py:12): per chunk origin = min_bound - 0.5*voxel
 * and the index range -> key layout.  origins_dev: [n_chunks][3] float64.  plan_dev: one struct. */
int pch_voxel_plan_build(const int32_t* minmax_dev, int64_t n_chunks, int64_t chunk_size,
                         const double* scales, const double* offsets, double voxel_size,
                         double* origins_dev, pch_voxel_plan* plan_dev, pch_stream_t stream);

/* idx = floor((p - origin)/voxel) per axis in float64 with a true divide (open3d, via
 * ui/import_PC.py:8-13); key = ix|iy|iz|index-in-chunk packed per `plan` (a HOST copy). */
int pch_voxel_keys(const uint8_t* rec_dev, int64_t n, int32_t rec_len, int64_t chunk_size,
                   const double* scales, const double* offsets, double voxel_size,
                   const double* origins_dev, const pch_voxel_plan* plan, uint64_t* keys_dev,
                   int32_t* xyz16_dev /* nullable: (n,4) int32 = X,Y,Z,0 copy for the reduce gathers */,
                   pch_stream_t stream);

/* pch_voxel_keys from the packed lattice copy written by pch_las_chunk_minmax (no second pass over
 * the raw records). */
int pch_voxel_keys_xyz16(const int32_t* xyz16_dev, int64_t n, int64_t chunk_size, const double* scales,
                         const double* offsets, double voxel_size, const double* origins_dev,
                         const pch_voxel_plan* plan, uint64_t* keys_dev, pch_stream_t stream);

/* The same two steps for an arbitrary (n,3) float64 point array — process_chunk(points_chunk,
 * voxel_size) (ui/import_PC.py:8-13, ui/Sampling.py:10-18) is not restricted to LAS-lattice input.
 * minmax_scratch_dev: [n_chunks*6] uint64.  Then sort, then pch_voxel_reduce with rec_dev = the
 * float64 array, rec_len = 0, scales = offsets = NULL (only mean_dev is available). */
int pch_voxel_plan_build_f64(const double* xyz_dev, int64_t n, int64_t chunk_size, double voxel_size,
                             uint64_t* minmax_scratch_dev, double* origins_dev, pch_voxel_plan* plan_dev,
                             pch_stream_t stream);
int pch_voxel_keys_f64(const double* xyz_dev, int64_t n, int64_t chunk_size, double voxel_size,
                       const double* origins_dev, const pch_voxel_plan* plan, uint64_t* keys_dev,
                       pch_stream_t stream);

/* Wide keys (plan.status == PCH_ERR_RANGE: the index range does not fit one 64-bit word; open3d itself only
 * refuses when voxel_size * INT_MAX < extent).  pch_voxel_index3_f64 stores the (ix,iy,iz) triples;
 * pch_voxel_wide_words builds  triple[axis] << bits_idx | index  in the order of the previous round
 * (prev_dev NULL = input order); three stable sort rounds (axis 2, 1, 0) order the chunk; pch_voxel_reduce
 * then takes vidx_dev and compares triples instead of word prefixes. */
int pch_voxel_index3_f64(const double* xyz_dev, int64_t n, int64_t chunk_size, double voxel_size,
                         const double* origins_dev, int32_t* vidx_dev, pch_stream_t stream);
int pch_voxel_wide_words(const uint64_t* prev_dev, const int32_t* vidx_dev, int64_t n, int64_t chunk_size,
                         int32_t bits_idx, int32_t axis, uint64_t* out_dev, pch_stream_t stream);

/* Stable LSD radix sort of 64-bit keys on bits [bit_lo, bit_hi), independently inside consecutive
 * segments of `seg_size` keys.  n_passes = ceil((bit_hi-bit_lo)/8); the result is left in
 * keys_dev when n_passes is even, in tmp_dev when odd. */
size_t pch_sort_workspace_bytes(int64_t n, int64_t seg_size, int32_t bit_lo, int32_t bit_hi);
int pch_sort_u64_segmented(uint64_t* keys_dev, uint64_t* tmp_dev, int64_t n, int64_t seg_size,
                           int32_t bit_lo, int32_t bit_hi, void* workspace_dev, size_t workspace_bytes,
                           pch_stream_t stream);

/* Per voxel (run of equal key>>bits_idx in a sorted chunk): running float64 sum in input order,
 * mean = sum/count (open3d AccumulatedPoint, via ui/import_PC.py:12-13).  Any of the three outputs
 * may be NULL:  mean_dev (m,3) f64 = process_chunk's return;  lattice_dev (m,3) int32 = the
 * re-quantised values `las.x = final_points[:,0]` stores (ui/import_PC.py:61-63);  f32_dev (m,3) =
 * astype(float32) of that file read back (utils/tower_extraction.py:60-62);  z32_dev (m) = the z
 * column of f32_dev as a dense array (the input of the percentile select, :82).
 * chunk_counts_dev[n_chunks] and total_dev[1] (int64) receive the voxel counts. */
size_t pch_voxel_reduce_workspace_bytes(int64_t n, int64_t chunk_size);
int pch_voxel_reduce(const uint64_t* sorted_keys_dev, int64_t n, int64_t chunk_size, int32_t bits_idx,
                     const uint8_t* rec_dev, int32_t rec_len,
                     const int32_t* xyz16_dev /* nullable: pch_voxel_keys' packed copy; gathers read it instead of rec_dev */,
                     const int32_t* vidx_dev /* nullable: (n,3) voxel indices for the wide-key path (see below) */,
                     const double* scales, const double* offsets,
                     double* mean_dev, int32_t* lattice_dev, float* f32_dev, float* z32_dev /* nullable */,
                     int64_t* chunk_counts_dev, int64_t* total_dev, void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);

/* The whole voxel stage of ui/import_PC.py:45-63 (every chunk of the file: open3d grid per chunk, concat,
 * LAS re-quantisation, float32 read-back) as ONE call without a host round trip inside: chunk extrema + the
 * 16-byte lattice copy -> plan (kept in plan_dev, DEVICE memory) -> keys + radix digit histograms -> scan and
 * radix passes (launched for the widest key the chunk size allows; the passes the plan does not need exit at
 * once) -> in-order reduce (reads the plan to know which of keys_dev / tmp_dev holds the sorted keys).
 * Scratch owned by the caller: xyz16_dev (n,4) int32, keys_dev / tmp_dev (n) uint64, minmax_dev [n_chunks][6]
 * int32, origins_dev [n_chunks][3] float64.  Outputs as for pch_voxel_reduce (each nullable, sized n rows).
 * After the call the first 64 bytes of workspace_dev hold, for ONE device->host read:
 *   int32 @0  error word of the look-backs (0 = fine)
 *   int64 @8  M, the number of voxels; -1 = the voxel index range does not fit one 64-bit key word
 *             (voxel_size far too small for the chunk extent): nothing was reduced, take the wide-key path
 *   pch_voxel_plan @32  the plan the device used. */
size_t pch_voxel_downsample_las_workspace_bytes(int64_t n, int64_t chunk_size);
int pch_voxel_downsample_las(const uint8_t* rec_dev, int64_t n, int32_t rec_len, int64_t chunk_size,
                             const double* scales, const double* offsets, double voxel_size,
                             int32_t* xyz16_dev, uint64_t* keys_dev, uint64_t* tmp_dev, int32_t* minmax_dev,
                             double* origins_dev, pch_voxel_plan* plan_dev,
                             double* mean_dev, int32_t* lattice_dev, float* f32_dev, float* z32_dev,
                             int64_t* chunk_counts_dev, void* workspace_dev, size_t workspace_bytes,
                             pch_stream_t stream);

/* Self-test: the voxel kernels evaluate `sum/count`, `(mean-offset)/scale` and `(p-origin)/voxel` as a
 * reciprocal product plus two FMA corrections (Markstein), which must equal the IEEE divide bit for bit.
 * mismatches_dev[0] (int64) = number of a_dev[i] for which it does not (expected: 0).  Returns
 * PCH_ERR_INVALID for divisors the kernels themselves would route to the true divide. */
int pch_selftest_fastdiv(const double* a_dev, int64_t n, double b, int64_t* mismatches_dev, pch_stream_t stream);
/* the float32 twin (the cell index of the grid min-z kernels, `(p - min) / cell`) */
int pch_selftest_fastdiv_f32(const float* a_dev, int64_t n, float b, int64_t* mismatches_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- tower extraction, stages A/B */

/* centroid = np.mean(raw_points_f32, axis=0) (utils/tower_extraction.py:63): numpy's SEQUENTIAL
 * float32 column sums (sums3_dev, float[3]) divided in float64 and cast to float32 (centroid3_dev).
 * With a workspace the sums are evaluated in parallel, bit-identically (binade-local integer maps
 * composed by scans; real float32 adds wherever a map would not be exact). */
size_t pch_f32_centroid_workspace_bytes(int64_t m);
int pch_f32_centroid(const float* xyz_dev, int64_t m, float* sums3_dev, float* centroid3_dev,
                     void* workspace_dev /* NULL = serial reference kernel */, size_t workspace_bytes,
                     pch_stream_t stream);

/* points = raw_points - centroid (utils/tower_extraction.py:64), float32 subtract.  zs_dev (m) gets
 * the shifted z column, shifted_dev (m,3) the whole shifted cloud; either may be NULL. */
int pch_f32_shift(const float* xyz_dev, int64_t m, const float* centroid3_dev, float* zs_dev,
                  float* shifted_dev, pch_stream_t stream);

/* One column of an (m,3) float32 cloud as a dense (m) array (z_values = points[:, 2],
 * utils/tower_extraction.py:81, before the shift: x -> float32(x - c) is monotone, so the order
 * statistics of the shifted column are the shifted order statistics of the raw column). */
int pch_f32_column(const float* xyz_dev, int64_t m, int32_t column, float* out_dev, pch_stream_t stream);

/* The two order statistics np.percentile(z_values, 25) interpolates between
 * (utils/tower_extraction.py:82): out2_dev = {sorted[rank0], sorted[rank1]} exactly. */
size_t pch_select_workspace_bytes(void);
int pch_select_f32(const float* v_dev, int64_t n, int64_t rank0, int64_t rank1, float* out2_dev,
                   void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);

/* filtered_points = points[z_values > thr] (utils/tower_extraction.py:83-89), order preserving.
 * keep[i] = zs_dev[i] > thr, or keep_mask_dev[i] != 0 when zs_dev is NULL, or — both NULL —
 * float32(xyz[i].z - centroid.z) > thr computed on the fly (no shifted-z array needed).  out_xyz_dev (kept,3)
 * receives xyz - centroid (float32; centroid3_dev may be NULL = no shift), out_src_dev the source
 * index of every kept point, out_mask_dev (m) the keep flags; each may be NULL.  count_dev: int64. */
size_t pch_compact_workspace_bytes(int64_t m);
int pch_compact_points(const float* xyz_dev, const float* zs_dev, const uint8_t* keep_mask_dev, int64_t m,
                       const float* centroid3_dev, float thr, float* out_xyz_dev, int32_t* out_src_dev,
                       uint8_t* out_mask_dev, int64_t* count_dev, void* workspace_dev,
                       size_t workspace_bytes, pch_stream_t stream);

/* Grid min-z ground model (north_star subsystem 2; the reference has no such code, cf. its percentile and
 * RANSAC variants test/main_ground.py:8-131): p = xyz - centroid (float32), cell (i,j) = floor((p.xy - min_xy) /
 * cell) in float32, ground = min p.z per cell, keep = p.z - ground > hag.  oracle/ground.py::grid_min_keep_mask.
 *   pch_grid_min            : cell_min_dev[nx*ny] (uint32, order-preserving encoding of the float32 minimum).
 *                             The centroid shift is applied on the fly (centroid3_dev nullable = no shift); tiles
 *                             are staged in shared memory, lanes of a warp that fall into the same cell reduce
 *                             with one warp-shuffle minimum, a per-CTA cell table in shared memory absorbs the
 *                             rest: one global atomicMin per (tile, cell).
 *   pch_compact_points_grid : pch_compact_points whose keep flag is the height-above-ground test, derived from
 *                             the cloud and the cell table inside the compaction (no mask, no shifted cloud).
 *   pch_grid_min_ground     : the two steps with explicit outputs (keep mask, ground z per point) on an
 *                             already shifted cloud. */
int pch_grid_min(const float* xyz_dev, int64_t m, const float* centroid3_dev, float min_x, float min_y, float cell,
                 int32_t nx, int32_t ny, uint32_t* cell_min_dev, pch_stream_t stream);
int pch_compact_points_grid(const float* xyz_dev, int64_t m, const float* centroid3_dev, float min_x, float min_y,
                            float cell, int32_t nx, int32_t ny, float hag, const uint32_t* cell_min_dev,
                            float* out_xyz_dev, int32_t* out_src_dev, uint8_t* out_mask_dev, int64_t* count_dev,
                            void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);
int pch_grid_min_ground(const float* xyz_dev, int64_t m, float min_x, float min_y, float cell, int32_t nx,
                        int32_t ny, float hag, uint32_t* cell_min_dev, uint8_t* keep_dev,
                        float* ground_z_dev, pch_stream_t stream);

/* componentwise float32 min/max of an (m,3) cloud: out6_dev = minx,miny,minz,maxx,maxy,maxz. */
int pch_f32_minmax(const float* xyz_dev, int64_t m, float* out6_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- tower extraction, stage C/D */

typedef struct pch_cluster_stats {
    int64_t count;   /* points carrying this label                                   */
    float min[3];    /* axis-aligned bounds of the cluster (float32, exact)          */
    float max[3];
    double sum[3];   /* float64 coordinate sums (centroid = sum/count)               */
} pch_cluster_stats;

/* sklearn DBSCAN(eps, min_samples, algorithm='ball_tree') run independently on consecutive
 * `chunk`-point slices of the (G,3) float32 candidates, labels offset per chunk
 * (utils/tower_extraction.py:96-122).  Two calls because the radix-pass count depends on the data:
 *   pch_dbscan_plan  -> bounds_dev [n_chunks*6] uint32 + plan_dev (copy it to the host)
 *   pch_dbscan_run   -> labels_dev[G] int32 (identical to the reference's all_labels, -1 = noise),
 *                       n_clusters_dev[1] int64, stats_dev[min(K,max_clusters)] per-label reduction
 *                       (the bounding-box / centroid inputs of utils/tower_extraction.py:131-151 and
 *                       of the AABB variant test/008.py:302-319). */
int pch_dbscan_plan(const float* xyz_dev, int64_t G, int64_t chunk, double eps, uint32_t* bounds_dev,
                    pch_voxel_plan* plan_dev, pch_stream_t stream);
size_t pch_dbscan_workspace_bytes(int64_t G, int64_t chunk, const pch_voxel_plan* plan, int64_t max_clusters);
int pch_dbscan_run(const float* xyz_dev, int64_t G, int64_t chunk, double eps, int32_t min_samples,
                   const uint32_t* bounds_dev, const pch_voxel_plan* plan, int32_t* labels_dev,
                   int64_t* n_clusters_dev, pch_cluster_stats* stats_dev, int64_t max_clusters,
                   void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);

/* The same clustering as ONE call with no host round trip inside (the plan stays in device memory and every
 * kernel reads the key layout from it): bounds -> plan -> pch_dbscan_run's kernels.  After the call the first
 * 256 bytes of workspace_dev are the scalar block, for ONE device->host read:
 *   int32 @0 error word (0 = fine) | int64 @128 occupied cells | int64 @136 clusters K | uint32 @192 points
 *   outside dense cells | pch_voxel_plan @208 (status != 0: the cell grid does not fit one key word and nothing
 *   was clustered).  K > max_clusters: stats_dev holds only the first max_clusters rows; call again with more. */
size_t pch_dbscan_fused_workspace_bytes(int64_t G, int64_t chunk, int64_t max_clusters);
int pch_dbscan(const float* xyz_dev, int64_t G, int64_t chunk, double eps, int32_t min_samples,
               int32_t* labels_dev, pch_cluster_stats* stats_dev, int64_t max_clusters,
               void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);

/* Two-phase clustering for spatial tiles with a halo (parity definition: the un-chunked variant
 * test/zzzzz.py:79-84 on the concatenated cloud).  xyz_dev = [halo from the left neighbour | own points | halo
 * from the right neighbour] in one common frame.
 *   pch_dbscan_cores : clusters the array up to the ids of the CORE points: labels_dev[i] = local cluster id
 *                      (ordered by smallest core index) for core points, -1 elsewhere; scalar block as for
 *                      pch_dbscan.  workspace = pch_dbscan_fused_workspace_bytes(G, chunk, max_clusters).
 *   -- the ranks exchange (point, local id) of the core points both sides know exactly, join local clusters that
 *      share a point, and number the global clusters by their smallest core index (dist.py) --
 *   pch_dbscan_finish: map_dev[local id] = global id or -1; rewrites the core labels, gives every border point the
 *                      smallest GLOBAL id among the clusters owning a core point within eps (scikit-learn's rule on
 *                      the concatenated cloud) and reduces count / AABB / sums over the original indices
 *                      [own_lo, own_hi) only into stats_dev[n_global].  Same workspace, untouched in between;
 *                      acc_dev = pch_dbscan_acc_bytes(n_global) bytes of scratch.
 *   pch_label_min_index: table_dev[k] = min(table_dev[k], base + i - lo) over i in [lo, hi) with labels_dev[i] == k. */
int pch_dbscan_cores(const float* xyz_dev, int64_t G, int64_t chunk, double eps, int32_t min_samples, int32_t* labels_dev,
                     int64_t max_clusters, void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);
size_t pch_dbscan_acc_bytes(int64_t n_clusters);
int pch_dbscan_finish(const float* xyz_dev, int64_t G, int64_t chunk, double eps, int32_t min_samples,
                      const int32_t* map_dev, int64_t n_global, int64_t own_lo, int64_t own_hi,
                      int32_t* labels_dev, pch_cluster_stats* stats_dev, void* acc_dev, int64_t max_clusters_phase1,
                      void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);
int pch_label_min_index(const int32_t* labels_dev, int64_t lo, int64_t hi, int64_t base, int64_t n_labels,
                        int64_t* table_dev, pch_stream_t stream);

/* Projection of (n,3) float32 points on a horizontal axis, s = x*ux + y*uy in float64: the coordinate along which
 * corridor tiles abut.  pch_axis_extent: minmax_dev[2] = min, max of s.  pch_axis_band_mask: mask_dev[i] =
 * (lo <= s_i <= hi) and (labels_dev == NULL or labels_dev[i] >= 0): the halo a tile sends to its neighbour, and
 * the zone next to a cut in which both ranks know a point's core status exactly. */
int pch_axis_extent(const float* xyz_dev, int64_t n, double ux, double uy, double* minmax_dev, pch_stream_t stream);
int pch_axis_band_mask(const float* xyz_dev, int64_t n, double ux, double uy, double lo, double hi,
                       const int32_t* labels_dev /* nullable */, uint8_t* mask_dev, pch_stream_t stream);

/* trimesh.PointCloud(cluster_points).bounding_box_oriented (utils/tower_extraction.py:137-139) for a batch of
 * clusters, on the device: convex hull by gift wrapping (after an exact interior cull), then for EVERY hull-face
 * normal the minimum-area rectangle aligned with an edge of the projected hull; the smallest volume wins.  (trimesh
 * thins the normals on a 0.1 rad grid in Qhull's facet order first; every box it can return is among the candidates
 * here, so this box is never larger.)  points_dev: (L,3) float32 rows; ranges_dev: int64 [n_clusters][2] = first row,
 * end row of each cluster; one CTA per cluster.  status: 0 ok, 1 fewer than 4 points / no extent, 2 degenerate (flat
 * or collinear), 3 capacity (hull with more than 16384 faces): the caller falls back to its host path for those. */
typedef struct pch_obb_result {
    double extents[3];   /* rectangle long side, rectangle short side, thickness along the face normal */
    double center[3];    /* box centre in the frame of the points */
    double rotation[9];  /* row-major 3x3, columns = box axes (long, short, normal) in the frame of the points */
    double volume;
    int32_t n_faces, n_vertices, n_candidates, status;
} pch_obb_result;
size_t pch_obb_workspace_bytes(int32_t n_clusters);
int pch_obb_batch(const float* points_dev, const int64_t* ranges_dev, int32_t n_clusters, pch_obb_result* out_dev,
                  void* workspace_dev, size_t workspace_bytes, pch_stream_t stream);

/* ---------------------------------------------------------------- tiled RANSAC ground
 *
 * remove_ground_tiled_ransac(points, tile_size, distance_threshold, max_iterations) of test/main_ground.py:77-115:
 * squares of tile_size metres between np.arange edges (:84-91), and in every square with >= 10 points (:101) a
 * scikit-learn RANSACRegressor plane z = f(x, y) (:8-32): inliers are ground, outliers non-ground, both stacked
 * tile by tile (:110-113).  Points beyond the last edge and squares with < 10 points appear in NEITHER output,
 * as in the reference.
 *
 * pch_xy_minmax_f64: out4_dev = min x, min y, max x, max y of (n,3) float64 rows (np.min / np.max, :84-85);
 *   scratch32_dev: 32 bytes of device scratch.
 * pch_ransac_tile_words: word = tile << 32 | index with tile = i * (n_y_edges-1) + j, the order of the reference's two
 *   loops; points outside every tile get tile = (n_x_edges-1)*(n_y_edges-1).  *_edges3 (HOST pointers) = edges[0],
 *   edges[1], edges[1]-edges[0] of np.arange(min, max, tile_size): np.arange fills edges[i] = edges[0] + i*delta,
 *   and the kernel evaluates exactly that, so `(x >= e[i]) & (x < e[i+1])` selects the same points.
 *   Sorting the words (pch_sort_u64_segmented) and pch_word_bounds give each tile's slice; pch_gather_rows_f64
 *   gathers rows[word & 0xffffffff] of the first m words = `points[tile_mask]` for every tile, concatenated.
 * pch_ransac_tiles: RANSACRegressor.fit per tile, one CTA each (scikit-learn 1.x _ransac.py: min_samples 3,
 *   residual = |z - prediction| <= distance_threshold, more inliers wins, equal inliers -> R^2 on the inliers decides,
 *   max_trials shrinks by _dynamic_max_trials(stop_probability, 0.99 in the reference)).  The three sample rows of
 *   trial k (1-based) of tile t come from triples_dev[(t*max_trials + k-1)*3 ..] when given (a test replays the
 *   draws of scikit-learn's own generator this way), else from splitmix64 streams keyed by (seed, t, k).  The plane
 *   through the three samples is solved in closed form (LinearRegression on 3 points is that plane; samples
 *   collinear in xy, where LinearRegression returns a minimum-norm fit, are skipped as a trial).
 *   flags_dev[row] = 1 ground (inlier), 0 non-ground, 2 in neither output.  status: 0 ok, 1 fewer than
 *   min_tile_points rows, 2 no valid consensus set (scikit-learn raises ValueError).
 * pch_ransac_split: ordered compaction of the gathered rows into the two outputs; *_off_dev[t] = first output row
 *   of tile t (exclusive sums of n_inliers / n_points - n_inliers over the tiles with status 0). */
typedef struct pch_ransac_tile {
    int32_t n_points, n_trials, n_inliers, status;
    double anchor[3];   /* first sample of the winning trial */
    double slope[2];    /* z - anchor_z = slope[0]*(x - anchor_x) + slope[1]*(y - anchor_y) */
    double score;       /* R^2 of the winning trial on its inliers */
} pch_ransac_tile;
int pch_xy_minmax_f64(const double* points_dev, int64_t n, double* out4_dev, void* scratch32_dev, pch_stream_t stream);
int pch_ransac_tile_words(const double* points_dev, int64_t n, const double* x_edges3, int32_t n_x_edges,
                          const double* y_edges3, int32_t n_y_edges, uint64_t* words_dev, pch_stream_t stream);
int pch_gather_rows_f64(const double* points_dev, const uint64_t* words_dev, int64_t m, double* out_dev,
                        pch_stream_t stream);
int pch_ransac_tiles(const double* tile_points_dev, const int64_t* bounds_dev, int32_t n_tiles,
                     double distance_threshold, int32_t max_trials, double stop_probability, uint64_t seed,
                     const int32_t* triples_dev /* nullable */, int32_t min_tile_points, uint8_t* flags_dev,
                     pch_ransac_tile* tiles_out_dev, pch_stream_t stream);
int pch_ransac_split(const double* tile_points_dev, const uint8_t* flags_dev, const int64_t* bounds_dev, int32_t n_tiles,
                     const int64_t* ground_off_dev, const int64_t* other_off_dev, double* ground_out_dev,
                     double* other_out_dev, pch_stream_t stream);

/* `cluster_points = filtered_points[all_labels == label]` for every label at once
 * (utils/tower_extraction.py:133-134): pch_label_words builds (label << 32 | index) words (noise sorts
 * last), pch_sort_u64_segmented orders them by label (stable), pch_gather_rows_f32 gathers the rows of
 * the first m words; cluster k then occupies rows [sum(count[:k]), sum(count[:k+1])). */
int pch_label_words(const int32_t* labels_dev, int64_t G, uint64_t* words_dev, pch_stream_t stream);
int pch_gather_rows_f32(const float* xyz_dev, const uint64_t* words_dev, int64_t m, float* out_dev,
                        int32_t* src_index_dev /* nullable */, pch_stream_t stream);

/* ---------------------------------------------------------------- geoid shift + CRS */

typedef struct pch_geoid_grid {
    double ll_lat, ll_lon;   /* lower-left node (degrees); rows run south -> north, cols west -> east */
    double dlat, dlon;       /* node spacing (degrees)                                               */
    int32_t rows, cols;
    int32_t pitch;           /* floats per row in memory (>= cols, multiple of 4)                     */
    int32_t is_global;       /* cols*dlon spans 360 degrees: east neighbour of the last column wraps  */
} pch_geoid_grid;

typedef struct pch_tm_params {
    double rect_radius;      /* A = a/(1+n)(1+n^2/4+n^4/64+n^6/256)                                   */
    double beta[6];          /* Krueger inverse series coefficients                                    */
    double ecc;              /* first eccentricity e                                                   */
    double lon0_deg, k0, fe, fn;
} pch_tm_params;

/* PROJ vgridshift forward: out_h = h + multiplier*N(lat,lon), N bilinear in float64 over float32
 * nodes (utils/elevation_converter.py:29-31,48 with multiplier=+1; crs.py:25-35 with -1).
 * lat/lon in degrees.  out_n_dev (nullable) receives N; NaN outside the grid / on nodata. */
int pch_geoid_shift(const double* lat_dev, const double* lon_dev, const double* h_dev, int64_t n,
                    const float* grid_dev, const pch_geoid_grid* grid, double multiplier,
                    double* out_h_dev, double* out_n_dev, pch_stream_t stream);

/* Transformer.from_crs("EPSG:4547","EPSG:4326",always_xy=True).transform(x, y)
 * (utils/table_match_gim.py:72-75,232; test/005test.py:37,55): inverse Gauss-Krueger -> degrees. */
int pch_gk_inverse(const double* x_dev, const double* y_dev, int64_t n, const pch_tm_params* tm,
                   double* lon_dev, double* lat_dev, pch_stream_t stream);

/* Fused per-point conversion of a whole LAS (test/005test.py:48-66 + utils/elevation_converter.py:48):
 * decode -> (tm != NULL ? inverse GK : x,y already lon,lat) -> geoid shift; out_dev (n,3) float64 =
 * lon, lat, h + multiplier*N.  win_* name the grid window staged in shared memory by bulk TMA
 * (rows/cols 0 = read the grid from global memory); nodes outside the window are still read exactly. */
int pch_las_geodetic(const uint8_t* rec_dev, int64_t n, int32_t rec_len, const double* scales,
                     const double* offsets, const pch_tm_params* tm, const float* grid_dev,
                     const pch_geoid_grid* grid, int32_t win_row0, int32_t win_col0, int32_t win_rows,
                     int32_t win_cols, double multiplier, double* out_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- tower-level match */

/* haversine(lat1, lon1, lat2, lon2) in metres with R = 6371 km (utils/table_match_gim.py:17-34) for
 * every pair of match_towers' double loop (:168-192): out_dev[i*n2 + j], degrees in. */
int pch_haversine_matrix(const double* lat1_dev, const double* lon1_dev, int64_t n1, const double* lat2_dev,
                         const double* lon2_dev, int64_t n2, double* out_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- per-tower crop / preview */

/* test/kuangxuan.py:69-79 for every tower at once: mask_b = (x>=xmin_b)&(x<=xmax_b)&(y...)&(z...) on the
 * float64 decoded coordinates.  One streaming pass emits a word  b << 32 | point index  per hit, in no
 * particular order, into words_dev (room for `capacity` words); total_dev[0] (int64) = number of hits, which
 * may exceed capacity (re-run with more room).  Sorting the words on bits [0, 32+bits(n_boxes)) with
 * pch_sort_u64_segmented restores `points[mask_b]` order per tower; pch_word_bounds then gives each tower's
 * slice [bounds[b], bounds[b+1]) and pch_las_gather_f64 the points.  boxes_dev: [n_boxes][6] float64 =
 * xmin,ymin,zmin,xmax,ymax,zmax. */
int pch_las_box_crop(const uint8_t* rec_dev, int64_t n, int32_t rec_len, const double* scales,
                     const double* offsets, const double* boxes_dev, int32_t n_boxes, uint64_t* words_dev,
                     int64_t capacity, int64_t* total_dev, pch_stream_t stream);
int pch_word_bounds(const uint64_t* sorted_words_dev, int64_t m, int32_t n_boxes, int64_t* bounds_dev /* [n_boxes+1] */,
                    pch_stream_t stream);
/* out_dev (m,3) float64 = decoded points[words_dev[j] & 0xffffffff]  (`points[mask]`, `xyz[indices]`). */
int pch_las_gather_f64(const uint8_t* rec_dev, int64_t n, int32_t rec_len, const double* scales,
                       const double* offsets, const uint64_t* words_dev, int64_t m, double* out_dev,
                       pch_stream_t stream);
/* Preview subsample indices (pyGUI_towers_test.py:174-177 `np.random.choice(len(xyz), 200000, replace=False)`;
 * ui/vtk_widget.py:115-118): k distinct indices of [0, n) as words.  seed == 0: evenly spaced floor(j*n/k);
 * seed != 0: a keyed bijection of [0, n) (4-round Feistel network, cycle-walked) evaluated at 0..k-1.  The
 * reference draws from numpy's unseeded global RNG, so only the distribution can be matched, not the draw. */
int pch_sample_indices(int64_t n, int64_t k, uint64_t seed, uint64_t* words_dev, pch_stream_t stream);

/* ---------------------------------------------------------------- host staging for the PCIe hop */

/* HOST function (no device work): gathers the X,Y,Z int32 triple at bytes 0..11 of each of the n
 * `rec_len`-byte LAS point records — the only fields laspy's las.x/.y/.z read (ui/import_PC.py:47-48;
 * utils/tower_extraction.py:60-62) — into a dense 12-byte record stream in `xyz12_host` (normally pinned
 * staging memory; 16-byte aligned), with `n_threads` host threads (<= 0: all).  The result is a valid
 * record stream with rec_len = 12 for every device entry point above, so a 34-byte PDRF-3 file crosses
 * PCIe as 12 bytes per point.  No arithmetic is performed on the host. */
int pch_host_pack_xyz(const void* records_host, int64_t n, int32_t rec_len, void* xyz12_host,
                      int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* PCH_B200_H */
