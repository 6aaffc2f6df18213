/*
 * vapc_b200.h — C ABI of libvapc_b200.so: the B200 (sm_100a) implementation of the
 * voxelisation + per-voxel statistics path of 3dgeo-heidelberg/vapc.
 *
 * The reference has no FFI of its own (it is pure Python on pandas/numpy); the drop-in
 * boundary is the method level of `vapc.Vapc` / `vapc.BiTemporalVapc`.  Each entry point
 * below names the reference lines it replaces (paths relative to the reference repo).
 * `vapc_b200/` (Python) mirrors the reference classes on top of exactly these calls;
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer owned by the caller unless the
 *    parameter name ends in `_host`.  The library never allocates caller-visible memory;
 *    scratch space is passed in (`ws`, `ws_bytes`) and sized by the `*_workspace_bytes`
 *    queries.
 *  - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing
 *    synchronises unless documented.
 *  - return value: 0 = VAPC_OK, negative = VAPC_E_*; `vapc_last_error()` returns a
 *    thread-local message.  Nothing throws.
 *  - points are referred to by uint32 indices (N < 2^32 per GPU); voxel rows by uint32
 *    too.  Voxel rows are always in lexicographic (voxel_x, voxel_y, voxel_z) order, the
 *    order the reference emits (tests/test_voxel.py:217-229); points inside a voxel keep
 *    their original order (stable sort), so "first point of a voxel" is the lowest index
 *    (`drop_duplicates`, vapc.py:1208).
 */
#ifndef VAPC_B200_H
#define VAPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAPC_OK 0
#define VAPC_E_INVALID (-1)   /* bad argument */
#define VAPC_E_CUDA (-2)      /* CUDA runtime error, see vapc_last_error() */
#define VAPC_E_KEYSPACE (-3)  /* voxel bounding box needs more than 64 key bits */
#define VAPC_E_WORKSPACE (-4) /* workspace too small */
#define VAPC_E_RANGE (-5)     /* a point fell outside the grid given to vapc_pack_keys */

#define VAPC_ABI_VERSION 1

#if defined(__GNUC__)
#define VAPC_API __attribute__((visibility("default")))
#else
#define VAPC_API
#endif

/* Voxel grid + packed-key layout.  key = ((vx-vmin[0]) << (bits[1]+bits[2])) |
 * ((vy-vmin[1]) << bits[2]) | (vz-vmin[2]); x is most significant so that ascending key
 * order is lexicographic (vx, vy, vz) order.  key_bytes is 4 when bits sum <= 32 else 8. */
typedef struct vapc_grid {
  double origin[3];   /* Vapc.origin        (vapc.py:60)  */
  double voxel_size;  /* Vapc.voxel_size    (vapc.py:59)  */
  int64_t vmin[3];    /* lowest voxel index per axis in the cloud */
  int32_t bits[3];    /* key bits per axis */
  int32_t key_bytes;  /* 4 or 8 */
} vapc_grid_t;

/* LAS header transform: coordinate = X * scale + offset per axis (datahandler.py:72-95 gets these doubles from laspy). */
typedef struct vapc_las_transform {
  double scale[3];
  double offset[3];
} vapc_las_transform_t;

VAPC_API int vapc_abi_version(void);
/* number of CUDA kernels this library has launched in the process so far (bench.py's gpu_launches) */
VAPC_API uint64_t vapc_launch_count(void);
/* cudaLimitMaxL2FetchGranularity for the current device (32, 64 or 128 bytes): the per-voxel gathers read
 * one random 32-byte record per point, so 32 avoids fetching unused neighbour sectors from HBM. */
VAPC_API int vapc_set_l2_fetch_granularity(int32_t bytes);
VAPC_API const char* vapc_last_error(void);
/* Per-kernel timing for bench.py: while enabled, every kernel launch of the library is bracketed by CUDA
 * events on its own stream.  vapc_profile_collect waits for them and writes one line
 * "kernel<TAB>launches<TAB>total_ms" per kernel since the previous collect into out (NUL-terminated,
 * truncated to cap); returns the full text length. */
VAPC_API int vapc_profile_enable(int32_t on);
VAPC_API int64_t vapc_profile_collect(char* out, int64_t cap);

/* ---- bounding box (feeds vapc_grid_t; also Vapc.compute_reduction_point, vapc.py:153-162,
 * and the min/max of compute_percentage_occupied, vapc.py:761-763) ----------------------
 * out6 = {min x, min y, min z, max x, max y, max z}; ws >= vapc_minmax_workspace_bytes(). */
VAPC_API size_t vapc_minmax_workspace_bytes(void);
VAPC_API int vapc_minmax_f64(const double* x, const double* y, const double* z, int64_t n, double* out6,
                    void* ws, size_t ws_bytes, void* stream);

/* ---- (f)1 LAS integer coordinates as the input of the path: 12 B/point instead of 24 --------------------------
 * The three entry points below are vapc_minmax_f64, vapc_pack_records_bucketed and a column materialiser for clouds that
 * arrive as the int32 X, Y, Z of a LAS file plus the header's scale / offset.  The coordinate of a point is
 * X * scale + offset evaluated as one IEEE multiply and one IEEE add — bit-identical to the float64 columns the reference's
 * DataHandler holds (laspy's scaled x / y / z, datahandler.py:72-95) — so keys, records and every statistic downstream are
 * the same bits as on the fp64 path; only the bytes read from the host and from HBM halve. */
VAPC_API int vapc_minmax_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n, const vapc_las_transform_t* t_host,
                    double* out6, void* ws, size_t ws_bytes, void* stream);
VAPC_API int vapc_las_coords_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n,
                        const vapc_las_transform_t* t_host, double* x, double* y, double* z, void* stream);
VAPC_API int vapc_pack_records_bucketed_i32(const int32_t* X, const int32_t* Y, const int32_t* Z, int64_t n,
                                   const vapc_las_transform_t* t_host, const vapc_grid_t* grid_host, void* keys_bucketed,
                                   double* xyzw_bucketed, void* keys_orig, uint32_t* bucket_start,
                                   int32_t* bucket_bits_host, int32_t* bucketed_key_bytes_host, int32_t* status, void* ws,
                                   size_t ws_bytes, void* stream);

/* ---- a2 Vapc.voxelize (vapc.py:188-207): v = floor((coord - origin) / voxel_size) as int64,
 * IEEE subtract + IEEE divide + floor, bit-exact with numpy. --------------------------- */
VAPC_API int vapc_voxel_coords_f64(const double* x, const double* y, const double* z, int64_t n,
                          const vapc_grid_t* grid_host, int64_t* vx, int64_t* vy, int64_t* vz,
                          void* stream);

/* Host helper: fill vmin/bits/key_bytes of *grid_host from the cloud's coordinate bounding box
 * minmax6_host (floor is monotone, so voxel(min) is the min voxel).  VAPC_E_KEYSPACE if the
 * box needs more than 64 bits. */
VAPC_API int vapc_grid_from_bounds(vapc_grid_t* grid_host, const double* minmax6_host);

/* ---- packed keys (internal form of the voxel triple; the groupby key of vapc.py:724, :843,
 * :877, :959, :1081).  keys: uint32 or uint64 per grid->key_bytes.  xyzw (nullable): 32-byte
 * records {x, y, z, bit pattern of the point index} written alongside so that later per-voxel
 * gathers cost one sector per point.  status (int32, device, caller-zeroed): bit 0 set if a point fell outside the grid. */
VAPC_API int vapc_pack_keys_f64(const double* x, const double* y, const double* z, int64_t n,
                       const vapc_grid_t* grid_host, void* keys, double* xyzw, int32_t* status,
                       void* stream);
/* Same keys and records, but written in KEY-PREFIX BUCKET order: bucket b = the leading
 * *bucket_bits_host = min(8, bits[0]) bits of the key (high bits of the voxel x index); inside a
 * bucket the points keep their input order (stable).  bucket_start (device, uint32[257]) receives
 * the first slot of every bucket and n at [2^bucket_bits].  The record of a point carries its
 * ORIGINAL index as a bit pattern in the fourth slot, so everything downstream that gathers records
 * by bucketed position (vapc_reduce_moments, vapc_closest_to_cog) still reports original indices.
 * Purpose: the per-voxel gathers then stay inside an L2-sized window of records instead of costing a
 * 128-byte HBM line per random 32-byte record, and the buckets are the segments of
 * vapc_sort_pairs_segmented, which only has to sort the remaining key bits.  keys_orig receives the
 * keys in input order (the first step; vapc_point2voxel_dense wants them too). */
VAPC_API size_t vapc_bucket_workspace_bytes(int64_t n);
VAPC_API int vapc_pack_records_bucketed(const double* x, const double* y, const double* z, int64_t n,
                               const vapc_grid_t* grid_host, void* keys_bucketed, double* xyzw_bucketed,
                               void* keys_orig, uint32_t* bucket_start, int32_t* bucket_bits_host,
                               int32_t* bucketed_key_bytes_host, int32_t* status, void* ws, size_t ws_bytes,
                               void* stream);
/* *bucketed_key_bytes_host = width of the keys written to keys_bucketed (allocate it for grid->key_bytes): 4 when the grid has
 * 64-bit keys but the key bits BELOW the bucket digit fit 32 bits (grids of up to 40 key bits) — keys_bucketed then holds
 * only those low bits as uint32, vapc_sort_pairs_segmented moves 4-byte keys, and vapc_expand_bucket_keys rebuilds the full
 * sorted 64-bit keys (bucket of a sorted position from bucket_start) for the kernels downstream. */
VAPC_API int vapc_expand_bucket_keys(const uint32_t* low_sorted, int64_t n, const uint32_t* bucket_start,
                            int32_t bucket_bits, int32_t low_bits, uint64_t* keys_out, void* stream);
/* voxel triple back from keys (n keys -> 3 x int64) */
VAPC_API int vapc_unpack_keys(const void* keys, int64_t n, const vapc_grid_t* grid_host, int64_t* vx,
                     int64_t* vy, int64_t* vz, void* stream);

/* ---- stable LSD radix sort of (key, point index) pairs, 8-bit digits, in-house ---------
 * Sorts bits [0, end_bit).  idx_is_iota != 0: the contents of idx are ignored and taken to be
 * 0..n-1 (saves reading them in the first pass).  Buffers ping-pong: keys -> keys_alt ->
 * (keys_alt2 if given, else keys) -> keys_alt -> ...; pass a third buffer keys_alt2 to keep the
 * caller's keys intact (NULL: they are overwritten when more than one pass runs).  On return
 * *result_location_host: bits 0-1 = which key buffer holds the sorted keys (0 keys, 1 keys_alt,
 * 2 keys_alt2), bit 2 = sorted indices are in idx_alt (else idx).  key_bytes 4 or 8.
 * One read of the keys builds the digit histograms of all passes; each pass is then a single
 * kernel (rank in the tile, decoupled look-back for the global offsets, coalesced scatter). */
VAPC_API size_t vapc_sort_workspace_bytes(int64_t n);
/* test hook: force the 64-bit look-back words that n >= 2^30 elements need (default: chosen by n) */
VAPC_API int vapc_sort_debug_wide_lookback(int32_t on);
VAPC_API int vapc_sort_pairs(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx,
                    uint32_t* idx_alt, int32_t idx_is_iota, int64_t n, int32_t key_bytes,
                    int32_t end_bit, void* ws, size_t ws_bytes, int32_t* result_location_host,
                    void* stream);
/* The same over the key bits [begin_bit, end_bit) only (e.g. one pass over the leading byte of 32-bit indices). */
VAPC_API int vapc_sort_pairs_bits(void* keys, void* keys_alt, void* keys_alt2, uint32_t* idx, uint32_t* idx_alt,
                         int32_t idx_is_iota, int64_t n, int32_t key_bytes, int32_t begin_bit, int32_t end_bit,
                         void* ws, size_t ws_bytes, int32_t* result_location_host, void* stream);
/* Segmented form: the array is nseg <= 256 consecutive segments [seg_start[s], seg_start[s+1])
 * (device uint32[nseg+1], seg_start[nseg] = n), each sorted independently by bits [0, end_bit) in the
 * same launches; idx_is_iota numbers the elements by their global position.  Two key buffers
 * (keys is overwritten).  Same workspace size and result_location as vapc_sort_pairs. */
VAPC_API int vapc_sort_pairs_segmented(void* keys, void* keys_alt, uint32_t* idx, uint32_t* idx_alt,
                              int32_t idx_is_iota, int64_t n, int32_t key_bytes, int32_t end_bit,
                              const uint32_t* seg_start, int32_t nseg, void* ws, size_t ws_bytes,
                              int32_t* result_location_host, void* stream);

/* ---- dense grids: the buckets sorted through a per-bucket cell table instead of radix passes -------------
 * Replaces vapc_sort_pairs_segmented + vapc_count_voxels when the key bits below the bucket digit number 7..16 (grids of up
 * to 24 key bits: 2^low_bits cells of one bucket fit 16-bit counters in one SM's shared memory) and the keys are 32-bit:
 * per bucket count -> scan -> place -> put every cell's positions in ascending order.  keys_bucketed / bucket_start: as
 * written by vapc_pack_records_bucketed (not modified).  keys_sorted, pos_sorted: the same contents
 * vapc_sort_pairs_segmented would produce (stable: input order inside a voxel), bit for bit.  *nvoxels (device int64)
 * and ws: as vapc_count_voxels leaves them (vapc_reduce_moments runs next on the same ws).  status (device int32, caller-
 * zeroed): bit 2 is set when the method does not apply to this cloud — a cell holding more than 128 points — and the
 * outputs are then undefined: fall back to vapc_sort_pairs_segmented + vapc_count_voxels.  (The pandas groupby this
 * stands for: vapc.py:724, :843, :959.) */
VAPC_API int vapc_sort_cells_segmented(const void* keys_bucketed, int64_t n, int32_t key_bytes, int32_t low_bits,
                              const uint32_t* bucket_start, int32_t nseg, void* keys_sorted, uint32_t* pos_sorted,
                              int64_t* nvoxels, int32_t* status, void* ws, size_t ws_bytes, void* stream);

/* ---- run-length segmentation of sorted keys into the voxel table ----------------------
 * Step 1: count voxels.  Writes per-tile head counts + their exclusive scan into ws and the
 * number of voxels into *nvoxels (device int64).  Step 2 (vapc_reduce_moments) needs the same
 * ws untouched. */
VAPC_API size_t vapc_segment_workspace_bytes(int64_t n);
VAPC_API int vapc_count_voxels(const void* keys_sorted, int64_t n, int32_t key_bytes, int64_t* nvoxels,
                      void* ws, size_t ws_bytes, void* stream);

/* ---- a4, a7, a8, a12 single-pass segmented reduction (vapc.py:715-726, 829-887, 929-1040)
 * For every voxel v (row order = key order): count[v], first_idx[v] (lowest original index),
 * start[v] (first sorted position; start[V] = n), voxel_keys[v], and the FP64 moments
 *   mean3[a][v] = mean of (p_a - c_a(v)),   c(v) = origin + (v + 0.5) * voxel_size,
 *   m2[k][v]    = SUM (p_a - mean_a)(p_b - mean_b)   (k over xx, xy, xz, yy, yz, zz).
 * c(v) is an exact function of the key, so (count, mean3, m2) triples of the same voxel from
 * different tiles or different GPUs combine with Chan's pairwise update without loss
 * (vapc_merge_partials).  Inside a voxel, points are visited in original point order.
 * Layout: mean3 is [3][V], m2 is
 * [6][V] (column stride V = nvoxels_host).  xyzw = the records vapc_pack_keys_f64 (input order) or
 * vapc_pack_records_bucketed (bucket order) wrote, pos_sorted = the sorted payload = position of each
 * point's record in xyzw; original point indices come from the records.  idx_sorted (nullable)
 * receives the original point index of every sorted position, row_sorted (nullable) its voxel row:
 * together they are the point -> voxel map in sorted order.  To get it in ORIGINAL point order
 * (point2voxel[idx_sorted[p]] = row_sorted[p]) without 1e8 random 4-byte writes to HBM, group the pairs
 * by the leading byte of the index (vapc_sort_pairs_bits, one pass) and vapc_scatter_u32 them: the writes
 * of a group then stay inside an L2-sized window.  Small grids: vapc_point2voxel_dense is cheaper. */
VAPC_API int vapc_reduce_moments(const void* keys_sorted, const uint32_t* pos_sorted, int64_t n,
                        const vapc_grid_t* grid_host, const double* xyzw, int64_t nvoxels_host,
                        void* voxel_keys, uint32_t* start, uint32_t* count, uint32_t* first_idx,
                        double* mean3, double* m2, uint32_t* row_sorted, uint32_t* idx_sorted,
                        void* ws, size_t ws_bytes, void* stream);
/* How a tile's voxels are summed follows the mean number of points per voxel (n / nvoxels_host): below 8 one thread per voxel with
 * groups of 8 lanes for voxels that are long for their tile, from 32 on four lanes per voxel (and a warp for voxels of 64 points or
 * more inside a tile, always) — in a fixed order, so results are reproducible run to run.  Test hook: force 1, 2 or 4 lanes per voxel (0: back to the rule). */
VAPC_API int vapc_reduce_debug_lanes(int32_t lanes);
/* idx_sorted[p] = original index of the point at sorted position p, read back from its record (vapc_reduce_moments writes the same
 * array when given idx_sorted; callers that only need it for the per-attribute statistics, vapc.py:393-406, form it on demand). */
VAPC_API int vapc_record_indices(const double* xyzw, const uint32_t* pos_sorted, int64_t n, uint32_t* idx_sorted, void* stream);
/* out[index[i]] = values[i] */
VAPC_API int vapc_scatter_u32(const uint32_t* index, const uint32_t* values, int64_t n, uint32_t* out, void* stream);

/* point -> voxel map in ORIGINAL point order through a rank bitmap over the packed key space: per 32 cells
 * {occupancy bits, row of the first occupied cell}, row(key) = first + popc(bits below the key's bit) — the rows
 * are in key order, so a voxel's row is its rank among the occupied keys.  2^total_bits / 4 bytes (4 MB for 2^24
 * cells): the lookups are cache hits and the pass costs 8 B/point of HBM traffic instead of the 64 B/point of the
 * scattered writes that vapc_reduce_moments does when given point2voxel.  keys_unsorted: the keys vapc_pack_keys_f64
 * wrote (keep them intact by giving vapc_sort_pairs a keys_alt2 buffer).  table: scratch of
 * vapc_point2voxel_table_bytes() (zeroed by the call); 8-byte aligned; total_bits <= 32. */
VAPC_API size_t vapc_point2voxel_table_bytes(int32_t total_bits);
VAPC_API int vapc_point2voxel_dense(const void* keys_unsorted, int64_t n, const void* voxel_keys,
                           int64_t nvoxels, int32_t key_bytes, int32_t total_bits, void* table,
                           size_t table_bytes, uint32_t* point2voxel, void* stream);
/* CTAs per SM of the lookup kernel of vapc_point2voxel_dense (1..32, default 16).  The lookup is bound by the rate of uncoalesced
 * 8-byte gathers, not by issue slots or DRAM: a host that runs it on a side stream beside vapc_voxel_stats (both only read the voxel
 * table of vapc_reduce_moments) lowers the residency so that both kernels share every SM.  Process-wide. */
VAPC_API int vapc_point2voxel_set_ctas_per_sm(int32_t ctas);

/* ---- a5, a7, a8, a12, a13, a14: per-voxel outputs from the moments ---------------------
 * (vapc.py:730-741 density, 829-852 centroid, 856-887 std ddof=1, 929-1040 covariance,
 *  1044-1090 eigenvalues — symmetric 3x3 cyclic Jacobi instead of LAPACK dgeev —, 1094-1146
 *  the 8 features + the 3 normalised eigenvalues the reference leaves in the frame.)
 * All outputs nullable; layouts [k][V].  cov6 order xx, xy, xz, yy, yz, zz; feat8 order
 * Sum_of_Eigenvalues, Omnivariance, Eigenentropy, Anisotropy, Planarity, Linearity,
 * Surface_Variation, Sphericity.  n == 1 voxels get NaN std/cov/eig/features as in the
 * reference (vapc.py:977-1006, 1057-1058). */
VAPC_API int vapc_voxel_stats(const void* voxel_keys, const uint32_t* count, const double* mean3,
                     const double* m2, int64_t nvoxels, const vapc_grid_t* grid_host,
                     double* density, double* cog3, double* std3, double* cov6, double* eig3,
                     double* eign3, double* feat8, void* stream);
/* The same for a table whose keys and moments live in a PERMUTED axis order (multi-GPU tables: the axis the ranks' chunks are
 * separated along leads the key, so that contiguous key ranges follow the chunks): axis_order_host[i] = the coordinate axis (0 x, 1 y,
 * 2 z) that internal axis i stands for.  cog3 / std3 / cov6 are written in x, y, z order; NULL = identity = vapc_voxel_stats. */
VAPC_API int vapc_voxel_stats_axes(const void* voxel_keys, const uint32_t* count, const double* mean3, const double* m2,
                          int64_t nvoxels, const vapc_grid_t* grid_host, const int32_t* axis_order_host, double* density,
                          double* cog3, double* std3, double* cov6, double* eig3, double* eign3, double* feat8, void* stream);

/* ---- a11 centre / corner of voxel (vapc.py:891-925): ((v*vs) + vs/2) + origin and
 * (v*vs) + origin, in the reference's operation order (bit-exact). out: [3][n]. */
VAPC_API int vapc_voxel_center_corner(const void* voxel_keys, int64_t nvoxels, const vapc_grid_t* grid_host,
                             double* center3, double* corner3, void* stream);

/* ---- a9, a10, a15 closest-to-centroid (vapc.py:776-825, 1188-1195) ---------------------
 * Per voxel: min over its points of sqrt((X-cog_x)^2 + (Y-cog_y)^2 + (Z-cog_z)^2) (plain
 * products, left-to-right sum, no FMA — numpy's evaluation order) and the LOWEST original
 * index attaining it.  pos_sorted / xyzw as in vapc_reduce_moments. */
VAPC_API int vapc_closest_to_cog(const void* keys_sorted, const uint32_t* pos_sorted, int64_t n,
                        int32_t key_bytes, const double* xyzw, const double* cog3,
                        int64_t nvoxels_host, double* min_dist, uint32_t* argmin_idx, void* ws,
                        size_t ws_bytes, void* stream);
/* Per point (original order): distance to its voxel's centroid (vapc.py:790-794). */
VAPC_API int vapc_point_distance(const double* x, const double* y, const double* z, int64_t n,
                        const uint32_t* point2voxel, const double* cog3, int64_t nvoxels,
                        double* distance, void* stream);

/* ---- a12, the reference's own numbers: cov_** exactly as vapc.py:947-1006 produces them -----------------------------------
 * vapc_voxel_stats returns the covariance M2 / (n - 1) from centred moments (accurate to ~1e-13 at any coordinate offset).  The
 * reference forms (S_ab - S_a S_b / n) / (n - 1) from raw sums, which loses every digit once |coordinate|^2 / variance nears
 * 1e16 (UTM northings: O(1) errors).  For callers who must DIFF against the reference's output this entry point reproduces its
 * columns bit for bit, noise included: per-point products rounded to fp64, per-voxel Kahan sums in row order (pandas' groupby
 * sum), one rounding per operation of the closing formula.  xyzw / pos_sorted / start as vapc_reduce_moments reads and writes
 * them; cov6 = [6][nvoxels] (xx xy xz yy yz zz), NaN for single-point voxels.  Pinned to the unmodified reference by
 * tests/golden (assert_array_equal). */
VAPC_API int vapc_cov_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start,
                               int64_t nvoxels, double* cov6, void* stream);
/* ---- a7 / a9 / a10, the reference's own numbers: cog_* exactly as vapc.py:843-850 produces them ---------------------------------
 * pandas' groupby mean: the Kahan-compensated sum of the voxel's coordinates in row order, divided by the count.  It differs from
 * the exact mean of vapc_voxel_stats in the last bit at most — but `distance`, `min_distance` and the rows the reference keeps in
 * its closest-to-centroid reduction (`distance == min_distance`, vapc.py:1195) are functions of THIS double: with it as the cog3
 * argument of vapc_point_distance / vapc_closest_to_cog the three columns and the selected rows (exact ties included) equal the
 * reference's bit for bit.  cog3 = [3][nvoxels].  Pinned to the unmodified reference by tests/golden (assert_array_equal). */
VAPC_API int vapc_cog_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start,
                               int64_t nvoxels, double* cog3, void* stream);
/* ---- a8, the reference's own numbers: std_* exactly as vapc.py:873-880 produces them ---------------------------------------------
 * pandas' groupby std: Welford's recurrence over the voxel's coordinates in row order, M2 / (n - 1), square root — bit for bit,
 * including the ~1e-7 relative error that recurrence has at UTM-sized coordinates (vapc_voxel_stats' std, from moments about a
 * shift inside the voxel, does not have it).  For callers who DIFF against the reference's columns.  std3 = [3][nvoxels], NaN for
 * single-point voxels.  Pinned to the unmodified reference by tests/golden (assert_array_equal). */
VAPC_API int vapc_std_reference_formula(const double* xyzw, const uint32_t* pos_sorted, const uint32_t* start,
                               int64_t nvoxels, double* std3, void* stream);
/* ---- broadcast / gather helpers: out[i] = table[index[i]]  (the per-point broadcast of
 * groupby.transform / merge(how="left"), vapc.py:724-726, 843-850, 1087-1089) ------------ */
VAPC_API int vapc_gather_f64(const double* table, const uint32_t* index, int64_t n, double* out, void* stream);
VAPC_API int vapc_gather_u32(const uint32_t* table, const uint32_t* index, int64_t n, uint32_t* out, void* stream);

/* ---- a16 mode / mode_count of an integer attribute (vapc.py:407-433) --------------------
 * vapc_mode_keys builds composite keys (voxel row << attr_bits) | int(attr) from the per-point
 * attribute column (fp64 holding integers, truncated toward zero like astype(int),
 * vapc.py:409, :423); status bit 1 = negative / NaN value (np.bincount would raise), bit 2 =
 * value needs more than attr_bits.  After vapc_sort_pairs on those keys (8-byte keys,
 * end_bit = voxel-row bits + attr_bits) the points are ordered by (voxel, value) and the voxel
 * slices are the same [start[v], start[v+1]) as before.  vapc_mode_count then gives per voxel:
 * mode = smallest most frequent value (np.argmax(np.bincount)), mode_count = number of distinct
 * values whose share count / n >= threshold (fp64 division, as the reference).  Either output
 * nullable.  key_bytes: 4 when voxel-row bits + attr_bits <= 32 (half the sort traffic), else 8. */
VAPC_API int vapc_mode_keys(const uint32_t* point2voxel, const double* attr, int64_t n, int32_t attr_bits,
                   int32_t key_bytes, void* keys, int32_t* status, void* stream);
VAPC_API int vapc_mode_count(const void* keys_sorted, int64_t n, int32_t attr_bits, int32_t key_bytes,
                    const uint32_t* start, int64_t nvoxels, double threshold, int64_t* mode,
                    int64_t* mode_count, void* stream);

/* mode / mode_count over N GPUs ("(key, value, count) runs instead of moments").  vapc_run_lengths: run-length
 * encoding of sorted 64-bit keys with optional weights (NULL = 1 each): run_keys[r] = r-th distinct key, run_weight[r] = sum of
 * the weights of its elements; call vapc_count_voxels(keys_sorted, n, 8, ...) first for the number of runs and pass the same
 * ws.  vapc_mode_count_runs: vapc_mode_count on such runs, keys = voxel row << attr_bits | value, sorted and distinct. */
VAPC_API int vapc_run_lengths(const uint64_t* keys_sorted, const uint32_t* weights, int64_t n, int64_t nruns_host, uint64_t* run_keys,
                     int64_t* run_weight, void* ws, size_t ws_bytes, void* stream);
VAPC_API int vapc_mode_count_runs(const uint64_t* run_keys, const int64_t* run_weight, int64_t nruns, int32_t attr_bits,
                         int64_t nvoxels, double threshold, int64_t* mode, int64_t* mode_count, void* stream);

/* ---- a16 (rest) mean / sum / min / max / var (ddof = 0) / median of an fp64 attribute
 * (vapc.py:393-406).  order = point indices grouped by voxel ([start[v], start[v+1]) slices):
 * idx_sorted for vapc_attr_stats (original order inside a voxel), and the (voxel, value) order
 * produced by sorting vapc_f64_sort_keys by value, then stably by voxel row, for the median.
 * std = sqrt(var).  All outputs nullable.  sum / mean / var are added the way numpy adds the reference's per-voxel slice (fewer
 * than 8 values left to right, up to 128 in eight interleaved accumulators, beyond that numpy's pairwise halves; var as the
 * pairwise sum of the rounded squared deviations from the rounded mean, no FMA): the columns equal the reference's bit for
 * bit (tests/golden: assert_array_equal). */
VAPC_API int vapc_attr_stats(const double* attr, const uint32_t* order, const uint32_t* start,
                             int64_t nvoxels, double* sum, double* mean, double* vmin, double* vmax,
                             double* var, void* stream);
VAPC_API int vapc_f64_sort_keys(const double* attr, int64_t n, uint64_t* keys, void* stream);
VAPC_API int vapc_attr_median(const double* attr, const uint32_t* order, const uint32_t* start,
                              int64_t nvoxels, double* median, void* stream);

/* ---- a18 mask membership by hashed key (MultiIndex.isin, vapc.py:594-598) ---------------
 * The mask's voxel triples are packed to uint64 keys in the mask's own grid (needs <= 63 bits;
 * triples outside the grid pack to the reserved key ~0 and status bit 0 is set) and inserted
 * into an open-addressing table (uint64[capacity], capacity a power of two >= 2 * nmask,
 * linear probing, multiplicative hash).  vapc_mask_lookup_points fuses voxelisation of the
 * queried points (vapc.py:199-201, same origin / voxel_size as the mask grid) with the probe;
 * vapc_mask_lookup_keys probes already packed keys.  member[i] = 1 / 0. */
VAPC_API int vapc_pack_triples_i64(const int64_t* vx, const int64_t* vy, const int64_t* vz, int64_t n,
                          const vapc_grid_t* grid_host, uint64_t* keys, int32_t* status, void* stream);
VAPC_API size_t vapc_mask_table_bytes(int64_t nmask, int64_t* capacity_host);
VAPC_API int vapc_mask_build(const uint64_t* mask_keys, int64_t nmask, void* table, int64_t capacity,
                    void* stream);
VAPC_API int vapc_mask_lookup_points(const double* x, const double* y, const double* z, int64_t n,
                            const vapc_grid_t* grid_host, const void* table, int64_t capacity,
                            uint8_t* member, void* stream);
VAPC_API int vapc_mask_lookup_keys(const uint64_t* keys, int64_t n, const void* table, int64_t capacity,
                          uint8_t* member, void* stream);

/* ---- a17 voxel buffer (vapc.py:495-541): every triple + all (2b+1)^3 offsets, in the
 * reference's expansion order (voxel-major, offsets in meshgrid 'ij' order); de-duplication is
 * done by the caller with sort + segment.  out: [3][n * (2b+1)^3]. */
VAPC_API int vapc_buffer_expand(const int64_t* vx, const int64_t* vy, const int64_t* vz, int64_t n,
                       int32_t buffer_size, int64_t* out3, void* stream);

/* ---- a22, a25 two-epoch voxel join (bi_vapc.py:78-84, 25-36) ----------------------------
 * keys1 / keys2: sorted, unique uint64 keys packed with the SAME grid.  For each row of epoch
 * 1: match1[i] = row in epoch 2 or 0xFFFFFFFF; likewise match2. */
VAPC_API int vapc_join_sorted_keys(const uint64_t* keys1, int64_t n1, const uint64_t* keys2, int64_t n2,
                          uint32_t* match1, uint32_t* match2, void* stream);

/* ---- a23, a24 centroid distance + Mahalanobis change test (bi_vapc.py:86-138, 198-207) --
 * For each pair r: rows i1[r], i2[r] of the two voxel tables (cog3 [3][V], cov6 [6][V]).
 * distance = |cog1 - cog2|; C += 1e-10 I; d1 = diff' inv(C2) diff; d2 = diff' inv(C1) diff;
 * p = 1 - chi2.cdf(d, 3) in closed form (1 - cdf, NOT a survival function: p == 0 must happen
 * where the reference's does); significance = p1 < alpha or p2 < alpha; p_value = p1 if
 * p1 < alpha else p2; change_type 1, 0 where not significant, 2 where p_value == 0.
 * All outputs nullable. */
VAPC_API int vapc_mahalanobis(const double* cog3_1, const double* cov6_1, int64_t v1, const double* cog3_2,
                     const double* cov6_2, int64_t v2, const uint32_t* i1, const uint32_t* i2,
                     int64_t npairs, double alpha, double* distance, double* d1, double* d2,
                     double* p_value, int32_t* significance, int32_t* change_type, void* stream);

/* ---- (f)4 neighbour queries on a voxelised point set used as a cell list ------------------
 * The searched set is described by what a voxelisation of it with cell size grid->voxel_size produced: cell_keys
 * (sorted packed keys, ncells), start[ncells + 1], pos_sorted[n] and the 32-byte records xyzw[n] (vapc_pack_keys_f64 /
 * vapc_pack_records_bucketed + sort + vapc_reduce_moments).
 *
 * vapc_cell_nn: exact nearest neighbour of every query point (bi_vapc.py:173-186 mutual_nearest_neighbors and
 * :51-74 merge_vapcs_with_closest_point: the KDTree.query(k=1) calls).  nn_idx = original index of the nearest point
 * of the set (lowest index among exactly equidistant ones), nn_dist = sqrt(dx*dx + dy*dy + dz*dz) formed like scipy's
 * KD-tree forms it (plain products, left-to-right sum).  Queries may lie outside the set's bounding box.  bounded != 0: a
 * query whose neighbour is farther than ~8 cells is given up (nn_idx = 0xFFFFFFFF, nn_dist = inf) instead of scanning ever
 * larger cubes — a grid degenerates to a scan of the whole set per query there; vapc_nn_all_pairs (tiled all-pairs search
 * over the listed queries `which`, same distances and tie rule) finishes those with a bounded cost.
 *
 * vapc_cluster_components: connected components of the graph "distance <= cluster_distance" between the n units of
 * the set (vapc.py:1274-1357 compute_clusters: KD-tree region growing, labels merged to the lowest).  Needs
 * grid->voxel_size >= distance (a 3x3x3 block of cells then holds every neighbour).  first (nullable = unit
 * index): the lowest original point index a unit stands for, weight (nullable = 1): its number of points — units are
 * single points, or occupied voxels standing for all their points.  Outputs per unit: root = lowest unit index of its
 * component; hop2_first = lowest `first` within two hops (a unit opens a new label in the reference's sequential sweep
 * iff hop2_first == first, which yields the reference's cluster ids: see vapc_b200/engine.py:cluster_components);
 * and at index root: comp_first = lowest `first` of the component, comp_size = its number of points. */
VAPC_API int vapc_cell_nn(const double* qx, const double* qy, const double* qz, int64_t nq, const vapc_grid_t* grid_host,
                 const void* cell_keys, int64_t ncells, const uint32_t* start, const uint32_t* pos_sorted,
                 const double* xyzw, int64_t n, int32_t bounded, uint32_t* nn_idx, double* nn_dist, void* stream);
VAPC_API int vapc_nn_all_pairs(const double* qx, const double* qy, const double* qz, const int64_t* which, int64_t nwhich,
                      const double* xyzw, int64_t n, uint32_t* nn_idx, double* nn_dist, void* stream);
VAPC_API size_t vapc_cluster_workspace_bytes(int64_t n);
VAPC_API int vapc_cluster_components(const vapc_grid_t* grid_host, const void* cell_keys, int64_t ncells, const uint32_t* start,
                            const uint32_t* pos_sorted, const double* xyzw, int64_t n, double distance,
                            const uint32_t* first, const uint32_t* weight, uint32_t* root, uint32_t* hop2_first,
                            uint32_t* comp_first, int64_t* comp_size, void* ws, size_t ws_bytes, void* stream);
/* the same code run serially on HOST arrays (CPU tests of the traversal logic) */
VAPC_API int vapc_selftest_cell_nn_host(const double* qx, const double* qy, const double* qz, int64_t nq,
                               const vapc_grid_t* grid_host, const void* cell_keys, int64_t ncells, const uint32_t* start,
                               const uint32_t* pos_sorted, const double* xyzw, int64_t n, int32_t bounded, uint32_t* nn_idx,
                               double* nn_dist);
VAPC_API int vapc_selftest_cluster_host(const vapc_grid_t* grid_host, const void* cell_keys, int64_t ncells, const uint32_t* start,
                               const uint32_t* pos_sorted, const double* xyzw, int64_t n, double distance,
                               const uint32_t* first, const uint32_t* weight, uint32_t* root, uint32_t* hop2_first,
                               uint32_t* comp_first, int64_t* comp_size);

/* ---- multi-GPU (north_star: partial accumulators exchanged by key range) ---------------
 * Merge partial voxel rows that share a key: inputs are R partial rows (keys sorted, equal
 * keys adjacent, in source-rank order), outputs the merged rows.  Same ws protocol as
 * vapc_count_voxels (call it first on the received keys to get the merged row count). */
VAPC_API int vapc_merge_partials(const void* keys_sorted, const uint32_t* order, int64_t nrows,
                        int32_t key_bytes, const uint32_t* count_in, const double* mean3_in,
                        const double* m2_in, int64_t in_stride, int64_t nvoxels_host,
                        void* voxel_keys, uint32_t* count, double* mean3, double* m2, void* ws,
                        size_t ws_bytes, void* stream);
/* Multi-GPU exchange of partial voxel rows.  Every rank builds its local table with its OWN grid (keys relative
 * to its own lowest voxel: fewer key bits, fewer radix passes, and no dependence on the global bounds);
 * vapc_global_keys re-expresses the (sorted) local voxel keys in the global layout (same lexicographic
 * order) and counts the leading hist_bits bits of them into hist (zeroed here; the splitters come from the
 * all-reduced histogram).  vapc_pack_partial_rows writes the rows [begin, end) that leave this rank as
 * row-major records of 11 x 8 bytes {key (uint64), count (int64), mean[3], M2[6]}.  vapc_merge_partial_rows
 * merges like vapc_merge_partials over two sources: order values < n_local address this rank's own rows
 * local_first + value of the SOA columns (never copied), the others the received `rows`[value - n_local]. */
/* Exchange planning on the device.  hist_all = the all-gathered per-rank histograms [world][2^hist_bits] of
 * vapc_global_keys.  out (int64 [world * world + world - 1]) = every rank's send counts to every destination (row r
 * = rank r) followed by the world - 1 key thresholds: destination g owns keys in [t_g, t_{g+1}).  Thresholds lie on
 * histogram-bin boundaries and balance the global number of partial rows per range.  vapc_merge_keys gathers the sort keys
 * of the rows a rank then merges: its own rows [own_first, own_first + n_own) followed by the received rows' keys. */
VAPC_API int vapc_plan_exchange(const uint32_t* hist_all, int32_t world, int32_t hist_bits, int32_t global_bits,
                       int64_t* out, void* stream);
VAPC_API int vapc_merge_keys(const uint64_t* global_keys, int64_t own_first, int64_t n_own, const double* rows,
                    int64_t n_in, int32_t key_bytes, void* keys_out, void* stream);
VAPC_API int vapc_selftest_plan_exchange_host(const uint32_t* hist_all_host, int32_t world, int32_t nbins,
                                     int32_t shift, int64_t* out_host);
VAPC_API int vapc_global_keys(const void* voxel_keys, int64_t nvoxels, const vapc_grid_t* local_grid_host,
                     const vapc_grid_t* global_grid_host, int32_t hist_bits, uint64_t* keys_out,
                     uint32_t* hist, void* stream);
VAPC_API int vapc_pack_partial_rows(const uint64_t* global_keys, const uint32_t* count, const double* mean3,
                           const double* m2, int64_t nvoxels, int64_t begin, int64_t end, double* rows_out,
                           void* stream);
VAPC_API int vapc_merge_partial_rows(const void* keys_sorted, const uint32_t* order, int64_t nrows,
                            int32_t key_bytes, const double* rows, const uint32_t* local_count,
                            const double* local_mean3, const double* local_m2, int64_t local_stride,
                            int64_t local_first, int64_t n_local, int64_t nvoxels_host, void* voxel_keys,
                            uint32_t* count, double* mean3, double* m2, void* ws, size_t ws_bytes,
                            void* stream);
/* The owner's merge WITHOUT re-sorting its own rows: own = rows [own_first, own_first + n_own) of the local table (SoA, column stride
 * own_stride) with their global keys own_keys[n_own]; u = the received rows, already sorted and combined per key (SoA, stride nu,
 * 64-bit keys; e.g. by vapc_sort_pairs + vapc_merge_partial_rows with n_local = 0).  Both key-sorted, keys unique.  u_new_before
 * (device int64 [nu + 1]): exclusive prefix of "the own rows do not hold this key".  Output: the merged table of
 * nvoxels_host = n_own + u_new_before[nu] rows in key order; rows with equal keys are combined with Chan's update (own first).
 * One pass over the own rows instead of a radix sort of all keys (the rows a rank keeps are the bulk: SURVEY 8(e)). */
VAPC_API int vapc_merge_sorted_tables(const uint64_t* own_keys, const uint32_t* own_count, const double* own_mean3, const double* own_m2,
                             int64_t own_stride, int64_t own_first, int64_t n_own, const uint64_t* u_keys,
                             const uint32_t* u_count, const double* u_mean3, const double* u_m2, int64_t nu,
                             const int64_t* u_new_before, int32_t key_bytes, int64_t nvoxels_host, void* out_keys,
                             uint32_t* out_count, double* out_mean3, double* out_m2, void* stream);
/* The same merge with the per-voxel statistics of vapc_voxel_stats_axes formed in the same pass from the merged moments while they
 * are in registers: out_mean3 / out_m2 may then be NULL (the merged moments are neither written nor read back: 72 B per voxel each
 * way).  Statistic outputs nullable as in vapc_voxel_stats, column stride nvoxels_host. */
VAPC_API int vapc_merge_sorted_tables_stats(const uint64_t* own_keys, const uint32_t* own_count, const double* own_mean3,
                                   const double* own_m2, int64_t own_stride, int64_t own_first, int64_t n_own,
                                   const uint64_t* u_keys, const uint32_t* u_count, const double* u_mean3, const double* u_m2,
                                   int64_t nu, const int64_t* u_new_before, int32_t key_bytes, int64_t nvoxels_host,
                                   void* out_keys, uint32_t* out_count, double* out_mean3, double* out_m2,
                                   const vapc_grid_t* grid_host, const int32_t* axis_order_host, double* density, double* cog3,
                                   double* std3, double* cov6, double* eig3, double* eign3, double* feat8, void* stream);
/* Multi-GPU closest-to-centroid: every rank proposes, per partial voxel row, its nearest point to the FINAL
 * centroid; the owner reduces the n candidates (rows[i] = merged voxel row, dist[i] >= 0, idx[i] = global point index >= 0)
 * to the lexicographic minimum (distance, index) per row — the single-GPU tie-break "lowest original index".  Rows without a
 * candidate keep best_dist = NaN (all-ones pattern) and best_idx = -1. */
VAPC_API int vapc_reduce_candidates(const int64_t* rows, const double* dist, const int64_t* idx, int64_t n, int64_t nrows,
                           double* best_dist, int64_t* best_idx, void* stream);
/* histogram of the top `hist_bits` key bits (splitter selection) */
VAPC_API int vapc_key_histogram(const void* keys, int64_t n, int32_t key_bytes, int32_t total_bits,
                       int32_t hist_bits, uint32_t* hist, void* stream);

/* ---- host-side self-tests of the fp64 helpers the kernels share (run on the CPU, no GPU
 * needed): Jacobi eigenvalues of n symmetric 3x3 (cov6 rows xx xy xz yy yz zz -> 3 descending
 * eigenvalues), Chan combination of n (count, mean3, m2) rows, and the Mahalanobis quadratic
 * form + chi-square p of n (cov6, diff3) rows.  All pointers are HOST pointers. */
VAPC_API int vapc_selftest_eig3_host(const double* cov6_host, int64_t n, double* eig3_host);
VAPC_API int vapc_selftest_chan_host(const double* rows_host, int64_t n, double* out10_host);
VAPC_API int vapc_selftest_mahalanobis_host(const double* cov6_host, const double* diff3_host, int64_t n,
                                   double* d_host, double* p_host);

/* ---- (f)1 LASzip decoding on the host (file I/O in front of the hot path; datahandler.py:72-73 gets it from
 * laspy[lazrs]).  Decodes npoints records of point_size bytes from a LAZ file image in host memory into
 * records_out_host.  compressor / chunk_size / items (nitems triples type, size, version) come from the
 * "laszip encoded" VLR (record id 22204).  Supported: compressor 2 (pointwise chunked) with POINT10 v2, GPSTIME11 v2, RGB12 v2 and BYTE v2
 * items, i.e. LAS point formats 0 - 3 (+ extra bytes) — the layout of the reference's vapc_in.laz / small_mask.laz — and
 * compressor 3 (layered chunked) with POINT14 v3 (+ RGB14 v3) (+ BYTE14 v3), i.e. LAS 1.4 point formats 6 / 7 (+ extra bytes) — the layout of its
 * tree_wind_condition.laz / als_multichannel_leg000_points.laz; anything else returns VAPC_E_INVALID with a message naming
 * the item.  A layered chunk whose layer decoders do not end exactly on their streams' last bytes is reported as corrupt. */
VAPC_API int vapc_laz_decode(const uint8_t* file_host, int64_t file_bytes, int64_t offset_to_points, int64_t npoints,
                    int32_t point_size, int32_t compressor, uint32_t chunk_size, const uint16_t* items_host,
                    int32_t nitems, uint8_t* records_out_host);

/* The mirror: npoints records -> LASzip stream with a fixed chunk_size; items POINT10 v2 / GPSTIME11 v2 / RGB12 v2 / BYTE v2 (compressor 2,
 * pointwise chunked) or POINT14 v3 (+ RGB14 v3) (+ BYTE14 v3) (compressor 3, layered chunked: layers that never change inside a chunk are
 * omitted).  The stream starts with the 8-byte position of its chunk table, counted from the start of the file:
 * stream_offset_host = the header's offset to point data.  *out_bytes_host receives the stream length (also when out_cap is too
 * small: VAPC_E_WORKSPACE).  Pinned: re-encoding the decoded records of each of the reference's four LAZ fixtures reproduces the
 * file's point data byte for byte, layer sizes and chunk table included (tests/test_laz.py). */
VAPC_API int vapc_laz_encode(const uint8_t* records_host, int64_t npoints, int32_t point_size, uint32_t chunk_size,
                    const uint16_t* items_host, int32_t nitems, int64_t stream_offset_host, uint8_t* out_host, int64_t out_cap,
                    int64_t* out_bytes_host);

#ifdef __cplusplus
}
#endif
#endif /* VAPC_B200_H */
