/*
 * yawb -- C ABI of the B200 pair-counting engine for yet_another_wizz.
 *
 * The reference (jlvdb/yet_another_wizz v3.1.1, pure Python) has no FFI; the
 * drop-in boundary is the set of Python seams listed in SURVEY.md section 8b.
 * Each entry point below names the reference interface it replaces (paths are
 * relative to the reference checkout, `/root/reference/`).  The Python host
 * (`yet_another_wizz_b200/_lib.py`) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the
 *     library owns every device buffer;
 *   - every function returns 0 on success, non-zero on failure with a message
 *     available from yawb_last_error() (thread local);
 *   - one context per process/GPU, not re-entrant per context; all work of a
 *     context is issued on one CUDA stream owned by the context;
 *   - there is no CPU fallback: without a CUDA device yawb_create() fails.
 */
#ifndef YAWB_H
#define YAWB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct yawb_ctx yawb_ctx;
typedef struct yawb_cat yawb_cat;

/* flags for yawb_count() */
#define YAWB_FLAG_EXACT_BRUTEFORCE 1u /* FP64 all-pairs kernel, no pruning (validation / cross-check) */
#define YAWB_FLAG_OUT_DEVICE 2u       /* out_f64 / out_i64 are device pointers (e.g. torch tensors)  */

/* roles for yawb_build_index() */
#define YAWB_ROLE_FIRST 1  /* catalog used as first argument of yawb_count(): sky-cell index        */
#define YAWB_ROLE_SECOND 2 /* catalog used as second argument: Hilbert-ordered register tiles        */

typedef struct {
    double kernel_ms;          /* device time of planner + pair-count kernel, CUDA events on the ctx stream */
    double index_ms;           /* device time of index builds triggered by this call (0 if prebuilt)      */
    uint64_t pair_tests;       /* pair tests executed (after sky-cell / tile pruning)                     */
    uint64_t pair_tests_naive; /* sum over requested patch pairs and z-bins of n1 * n2                    */
    uint64_t rechecks;         /* tests re-evaluated in exact FP64                                        */
    uint64_t work_items;       /* work items written by the planner (tile x linked patch, non-empty)      */
    uint64_t launches;         /* kernels launched by this call                                           */
    double plan_ms;            /* part of kernel_ms spent in the planner (k_plan); the rest is the count kernel */
} yawb_stats;

/* Thread-local description of the last failure. */
const char *yawb_last_error(void);

/* Create / destroy an engine context on CUDA device `device`.
 * Replaces the reference's worker farm set-up, src/yaw/utils/parallel.py:318-343
 * (a fresh multiprocessing.Pool per iter_unordered call). */
int yawb_create(int device, yawb_ctx **out);
int yawb_destroy(yawb_ctx *ctx);

/* Upload one patch-partitioned catalog into HBM (once) and compute per-patch
 * frames, bounding spheres and sums of weights.
 *
 * Replaces Catalog.build_trees / BinnedTrees.build / build_trees /
 * AngularTree.__init__ (src/yaw/catalog/catalog.py:1406-1460,
 * src/yaw/catalog/trees.py:365-429, 215-246): instead of one pickled cKDTree
 * per patch and z-bin the points are kept resident, sorted by
 * (patch, z-bin, sky cell).
 *
 *   xyz        n x 3 float64, exactly AngularCoordinates.to_3d() of the rows
 *              (src/yaw/coordinates.py:134-147), rows grouped by patch
 *   w          n float64 weights or NULL (unweighted: sum_weights = count,
 *              trees.py:225-227; counts are then integers)
 *   zbin       n int32 z-bin index np.digitize(z, edges, right=closed=="right") - 1
 *              (trees.py:408-414); rows with zbin outside [0, n_bins) are
 *              dropped.  NULL = unbinned catalog (one "tree" per patch that is
 *              paired with every z-bin of the first catalog, trees.py:600-601)
 *   patch_off  n_patch + 1 row offsets (patch ids are 0..n_patch-1,
 *              src/yaw/correlation/measurements.py:358-364)
 *   n_bins     number of z-bins (ignored when zbin == NULL)
 *
 * The call is asynchronous: it only allocates and enqueues the host-to-device copies on a dedicated copy
 * stream (which never carries a kernel, so copies of later catalogs proceed while pair counts of earlier
 * ones occupy the GPU); the per-patch reductions run when the catalog is first used.  The host buffers
 * must stay valid until the first call that uses the catalog (or yawb_sync) has returned.  Enqueue every
 * catalog first and count in arrival order to overlap PCIe with the counts.
 */
int yawb_upload_catalog(yawb_ctx *ctx, const double *xyz, const double *w, const int32_t *zbin,
                        const int64_t *patch_off, int n_patch, int n_bins, yawb_cat **out);
/* Same with byte-sized z-bin ids (n_bins <= 254; any id >= n_bins, e.g. 255, marks a dropped row): the ids
 * are 5 % of the bytes of a z-binned catalog as int32, 1.3 % as bytes. */
int yawb_upload_catalog_u8(yawb_ctx *ctx, const double *xyz, const double *w, const uint8_t *zbin,
                           const int64_t *patch_off, int n_patch, int n_bins, yawb_cat **out);
/* Same with the raw redshifts: the rows are assigned to z-bins ON THE DEVICE, with the arithmetic of
 * np.digitize(z, edges, right = closed_right) - 1 (src/yaw/catalog/trees.py:408-414, Binning edges of
 * src/yaw/binning.py): bin b holds edges[b] < z <= edges[b + 1] (closed_right) or edges[b] <= z < edges[b + 1];
 * rows outside the binning (and NaN) are dropped.  Comparisons only, so the ids are identical to numpy's.
 *   z       n redshifts;   edges   n_bins + 1 strictly increasing doubles (copied by the call) */
int yawb_upload_catalog_z(yawb_ctx *ctx, const double *xyz, const double *w, const double *z, const double *edges,
                          int closed_right, const int64_t *patch_off, int n_patch, int n_bins, yawb_cat **out);
int yawb_free_catalog(yawb_cat *cat);

/* Build (or rebuild) the device-side index for a role ahead of time; otherwise
 * yawb_count() builds it on first use.  Device time is returned in *ms. */
int yawb_build_index(yawb_cat *cat, int role, double *ms);
int yawb_drop_index(yawb_cat *cat);

/* Number of rows kept (zbin in range) and device bytes held by the catalog. */
int yawb_catalog_info(const yawb_cat *cat, int64_t *n_rows, int64_t *device_bytes);

/* Sum of weights per z-bin and patch, out[n_bins][n_patch] (n_bins = 1 for an
 * unbinned catalog): AngularTree.sum_weights collected by process_patch_pair,
 * src/yaw/correlation/measurements.py:123-124, trees.py:225-234. */
int yawb_sum_weights(const yawb_cat *cat, double *out);

/* Jackknife by subtraction on the device: the sum over all patch pairs and its leave-one-patch-out samples,
 *     total[b] = sum_k v[k][b],   samples[p][b] = total[b] - sum_{k : i_k == p or j_k == p} v[k][b]
 * -- SampledPatchSum of src/yaw/correlation/paircounts.py:113-141 (total - row_p - column_p + diagonal_p) written
 * on the list of linked patch pairs instead of the dense (n_patch x n_patch) array.
 *   values[n_pairs][n_bins]  per-pair, per-z-bin counts or products of sums of weights (host)
 *   total[n_bins], samples[n_patch][n_bins]   (host)
 * Integer-valued inputs give exact results (sums of integers below 2^53 do not depend on the order); others
 * agree with numpy's einsum to rounding. */
int yawb_jackknife(yawb_ctx *ctx, const double *values, const int32_t *pair_i, const int32_t *pair_j, int n_pairs,
                   int n_patch, int n_bins, double *total, double *samples);

/* Patch meta data computed on the device from the uploaded rows (all rows of the patch, whatever their z-bin):
 * the quantities of Metadata.compute, src/yaw/catalog/patch.py:104-147.
 *   center_xyz[n_patch][3]  normalised mean direction (the reference converts it to RA / Dec)
 *   radius_chord[n_patch]   largest chord distance of a row from the centre, rounded up by 1e-12 relative
 *                           (the reference's radius is the angle 2 asin(chord / 2))
 *   num_records[n_patch]    rows of the patch
 * Any pointer may be NULL.  The sums run in a different order than numpy's, so centres agree to ~1e-15, not
 * bit for bit; the reference-exact meta data of this package's Catalog stays on the host. */
int yawb_patch_metadata(const yawb_cat *cat, double *center_xyz, double *radius_chord, int64_t *num_records);

/* Count pairs for a list of linked patch pairs.
 *
 * Replaces the body of PatchLinkage.count_pairs -> process_patch_pair ->
 * AngularTree.count -> scipy cKDTree.count_neighbors
 * (src/yaw/correlation/measurements.py:307-367, 88-128;
 *  src/yaw/catalog/trees.py:303-362, scipy call :348-353).
 *
 * For pair k = (pair_i[k], pair_j[k]) and z-bin b of cat1, every point a of
 * cat1 (patch pair_i[k], bin b) is tested against every point c of cat2
 * (patch pair_j[k]; bin b if cat2 is binned, all rows otherwise):
 *
 *     d2 = (dx*dx + dy*dy) + dz*dz          IEEE double, no FMA (scipy's order)
 *     sub-bin s (0-based) iff r2[b][s] < d2 <= r2[b][s+1]
 *
 *   r2_edges  [n_bins][n_edges] squared chord thresholds, already
 *             pow(2 sin(theta/2), 2.0) of get_ang_bins() (trees.py:84-117, :350)
 *   out_i64   [n_pairs][n_bins][n_edges-1] number of pairs (may be NULL)
 *   out_f64   [n_pairs][n_bins][n_edges-1] sum of w1*w2 (may be NULL; equals the
 *             integer count when both catalogs are unweighted)
 *
 * The host applies the r-weights / scale sums (trees.py:358-362), the 0.5 on
 * auto diagonals (measurements.py:362-363) and the scatter into PatchedCounts.
 * Results are bit-exact for integer counts; weighted sums are accumulated in
 * FP64 (order differs from the tree walk: ~1e-15 relative).
 */
int yawb_count(yawb_ctx *ctx, yawb_cat *cat1, yawb_cat *cat2, const int32_t *pair_i,
               const int32_t *pair_j, int n_pairs, const double *r2_edges, int n_edges,
               uint32_t flags, double *out_f64, int64_t *out_i64, yawb_stats *stats);

/* Two first catalogs against the same second catalog in ONE pass: results identical to
 *     yawb_count(ctx, cat1a, cat2, ...) -> out_*_a   and   yawb_count(ctx, cat1b, cat2, ...) -> out_*_b.
 *
 * This is how crosscorrelate's DD + RD (reference sample and its randoms against the unknown sample) and
 * DR + RR (the same two against the unknown sample's randoms) are counted,
 * src/yaw/correlation/measurements.py:623-626: both first catalogs share the patches, the z-binning and the
 * thresholds, so their rows are indexed together (one fused sky-cell index, cached until either catalog is
 * dropped or freed) and every register tile of the second catalog is loaded, rotated and matched against the
 * sky cells once instead of twice.  cat1a and cat1b need the same n_patch / n_bins; YAWB_FLAG_EXACT_BRUTEFORCE
 * is not supported here.  stats cover both counts. */
int yawb_count2(yawb_ctx *ctx, yawb_cat *cat1a, yawb_cat *cat1b, yawb_cat *cat2, const int32_t *pair_i,
                const int32_t *pair_j, int n_pairs, const double *r2_edges, int n_edges, uint32_t flags,
                double *out_f64_a, int64_t *out_i64_a, double *out_f64_b, int64_t *out_i64_b, yawb_stats *stats);

/* The four counts of a cross-correlation in ONE launch: two first catalogs (fused index, as in yawb_count2)
 * against TWO second catalogs.  Results identical to
 *     yawb_count(ctx, cat1a, cat2a) -> out[0]    yawb_count(ctx, cat1b, cat2a) -> out[1]
 *     yawb_count(ctx, cat1a, cat2b) -> out[2]    yawb_count(ctx, cat1b, cat2b) -> out[3]
 * i.e. DD, RD, DR, RR of src/yaw/correlation/measurements.py:623-626 for (cat1a, cat1b, cat2a, cat2b) =
 * (reference, its randoms, unknown, its randoms).  The planner lists the work items of both second catalogs for
 * one persistent kernel, so a small job (a GPU's share of a multi-GPU run) pays the ramp-up and the tail of a
 * launch, the host synchronisation and the result transfer once instead of twice.  cat2a and cat2b need the same
 * n_patch / n_bins.  out_f64 / out_i64: tables of four pointers each, any of which may be NULL.
 * YAWB_FLAG_EXACT_BRUTEFORCE is not supported.  stats cover all four counts. */
int yawb_count4(yawb_ctx *ctx, yawb_cat *cat1a, yawb_cat *cat1b, yawb_cat *cat2a, yawb_cat *cat2b, const int32_t *pair_i,
                const int32_t *pair_j, int n_pairs, const double *r2_edges, int n_edges, uint32_t flags,
                double *const out_f64[4], int64_t *const out_i64[4], yawb_stats *stats);

/* Pinned host memory helpers so callers can stage inputs for async copies. */
int yawb_host_alloc(void **ptr, uint64_t bytes);
int yawb_host_free(void *ptr);

/* Block until all work of the context has finished. */
int yawb_sync(yawb_ctx *ctx);

/* Device-side stopwatch on the context's stream (CUDA events), for benchmarks: everything the
 * context enqueues between start and stop is covered.  stop waits for the stream. */
int yawb_timer_start(yawb_ctx *ctx);
int yawb_timer_stop(yawb_ctx *ctx, double *ms);

/* Nearest patch centre of every row, the catalog-creation step that precedes the hot path (SURVEY.md
 * section 8f, rank 2).  Replaces `assign_patch_centers` (src/yaw/catalog/catalog.py:229-249), i.e.
 * scipy.cluster.vq.vq(xyz, centers): squared Euclidean distance summed x -> y -> z in double without FMA,
 * first minimum wins.  xyz: n x 3 and centers_xyz: n_centers x 3 float64 on the host, out_ids: n int32 on
 * the host.  Synchronous. */
int yawb_assign_patches(yawb_ctx *ctx, const double *xyz, int64_t n, const double *centers_xyz, int n_centers,
                        int32_t *out_ids);

/* Library version and number of SMs of the context's device. */
int yawb_version(void);
int yawb_device_sms(const yawb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* YAWB_H */
