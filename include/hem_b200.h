/* hem_b200.h — C ABI of libhem_b200.so, the B200-native (sm_100a) replacement
 * for the hot path of erikszasz/GaussianSplattingRegistration: the hierarchical
 * EM (HEM) Gaussian-mixture downsampler of src/cpp_ext.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference repository root).  Plain pointers and sizes only; no
 * exceptions cross the boundary (0 = success, negative = error, message via
 * hem_last_error).  Data pointers may be HOST or DEVICE pointers (detected with
 * cudaPointerGetAttributes); the caller owns every buffer it passes; the
 * context owns all device memory it allocates.
 *
 * Reference call surface replaced (pybind11 module `mixture_bind`,
 * src/cpp_ext/mixture_bind.cpp:11-60):
 *   MixtureLevel.CreateMixtureLevel(xyz, colors, opacities, covariance, features)
 *       src/cpp_ext/src/mixturelevel.cpp:14-28       -> the five arrays of hem_build_pyramid
 *   MixtureCreator.CreateMixture(clusterLevel, hemReduction, distanceDelta, colorDelta, decayRate, level)
 *       src/cpp_ext/mixture_wrapper.cpp:10-18         -> hem_build_pyramid
 *   MixtureLevel.CreatePythonLists(level)
 *       src/cpp_ext/src/mixturelevel.cpp:30-70        -> hem_level_size + hem_level_copy
 */
#ifndef HEM_B200_H
#define HEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hem_ctx hem_ctx;

/* error codes */
#define HEM_OK               0
#define HEM_ERR_INVALID     -1   /* bad argument */
#define HEM_ERR_CUDA        -2   /* CUDA runtime failure (message holds the CUDA error string) */
#define HEM_ERR_NO_DEVICE   -3   /* no usable sm_100 device */
#define HEM_ERR_STATE       -4   /* call sequence error (e.g. level not built) */
#define HEM_ERR_NCCL        -5   /* NCCL failure / NCCL not loadable */

/* Create a context on CUDA device `device` (one context per GPU, one process
 * per GPU in multi-GPU runs).  `seed` keys the counter-based parent-selection
 * RNG that stands in for the reference's unseeded libc generator
 * (src/cpp_ext/include/base.hpp:44-56).  Returns NULL on failure; the reason is
 * then available from hem_last_error(NULL). */
hem_ctx* hem_create(int device, uint64_t seed);

/* Destroy the context and free every device allocation it owns.  (The
 * reference leaks its whole Mixture per call, mixture_wrapper.cpp:11.) */
void hem_destroy(hem_ctx* ctx);

/* Last error message of `ctx` (or of the failed hem_create when ctx == NULL). */
const char* hem_last_error(const hem_ctx* ctx);

/* Run on `stream` (a cudaStream_t) instead of the context's own stream; pass
 * NULL to return to the context's stream. */
int hem_set_stream(hem_ctx* ctx, void* stream);

/* Options (all optional):
 *   "timing"      1 -> record CUDA events around every stage (hem_step_stats ms fields);
 *                 2 -> only around each level and its two hot kernels (ms_total, ms_k_gate,
 *                 ms_k_moments; the other ms fields stay 0).  The event times are read on the
 *                 first hem_step_stats of a step after its build, not by the build itself.
 *   "keep_pairs"  1 -> retain each step's (parent, child, wL) lists for hem_step_pairs
 *   "cell_scale"  grid cell = cell_scale * geometric-mean parent radius (default 0.55)
 *   "smd_prefilter" 1 (default) -> certified Mahalanobis pre-test inside the KLD gate (same
 *                 accepted sets, fewer full KLD evaluations); 0 evaluates every candidate in full
 *   "tight_box"   1 (default) -> while the pre-test is on and no component of a level is uncertified, a certified parent
 *                 searches the axis-aligned box of its Mahalanobis ellipsoid (intersected with its sphere's box): every
 *                 candidate outside is one the pre-test rejects, so the accepted sets are unchanged; 0 -> sphere boxes
 *   "sumlw_mode"  0 (default) auto, 1 stable radix sort by child, 2 claim table: two bit-identical
 *                 ways of adding a child's claims in ascending parent order
 *   "overlap"     1 (default) -> stages of a step that do not depend on each other run side by side on a second stream
 *                 (control-block reset + parent-rank scan beside the grid build, orphan scan + copy beside the moment
 *                 kernel); 0 -> one chain.  Same results either way.
 *   "geo_first"   1 (default) -> a build whose SH rows come from HOST memory produces, level after level, the geometry
 *                 rows only for as long as the SH rows are still crossing PCIe (gates, claims and geometry moments need
 *                 the 64-byte geometry rows alone), sends bound geometry outputs back meanwhile, and runs the SH passes
 *                 of those levels once the rows have landed; 0 -> every step waits for the SH rows it reads; 2 -> defer
 *                 on every level of such a build (tests).  Results are bit-identical in all three.
 *   "poll_sync"   1 (default) -> the one host wait per level polls a mapped word instead of synchronising the stream
 *   "heavy_rows", "heavy_rows_rel"  a parent whose search box has more rows of cells than max(heavy_rows,
 *                 heavy_rows_rel x the box of a parent of geometric-mean radius) is scanned by the whole grid in
 *                 segments instead of by one warp (defaults 256 and 8; heavy_rows = 0: never; single GPU)
 *   "shard_min_n" multi-GPU: levels with fewer components are built whole on every rank instead of
 *                 being sharded (default 2000000; the per-level exchanges cost more than they save below it)
 *   "shard_upload" multi-GPU, host inputs: 1 (default) -> every rank uploads 1/world of each array over its own
 *                 PCIe link and the slices are all-gathered over NVLink; 0 -> every rank uploads everything
 *   "heavy_pairs" parents with more accepted pairs than this are accumulated by the grid-wide split kernel
 *                 instead of one warp (default 16384; 0 = never; single GPU)
 *   "cell_occupancy" lower bound of the grid cell edge in units of the mean component spacing (default 1.3; 0 = off)
 *   "box_quantile" the grid spans the [q, 1-q] percentile range of the level-0 means per axis (default 5e-4;
 *                 0 = full bounding box); outliers are clamped into border cells, results do not depend on it
 *   "rng_base"    counter of the first OUTPUT flag of the first step; default = n
 *                 (the reference consumes one draw per level-0 component first,
 *                 src/cpp_ext/src/mixture.cpp:330, then one per pre-filter
 *                 output, :257-259) */
int hem_set_option(hem_ctx* ctx, const char* name, double value);

/* Build `cluster_level` HEM levels from the n input splats — replaces
 * MixtureCreator::CreateMixture (src/cpp_ext/mixture_wrapper.cpp:10-18 ->
 * Mixture::initMixture src/cpp_ext/src/mixture.cpp:287-333 and
 * Mixture::createClusterLevel :66-285, once per level).
 *   xyz      [n*3]  means
 *   color    [n*3]  raw SH-DC colour
 *   opacity  [n]    raw opacity
 *   cov6     [n*6]  packed covariance 00,01,02,11,12,22 (src/utils/general_utils.py:20-29)
 *   sh       [n*sh_floats] SH rest coefficients (may be NULL iff sh_floats == 0); 0 <= sh_floats <= 256
 *            (45 at SH degree 3; the reference's FeatureVector takes any length, featurevector.cpp:20-63)
 *   weight   [n] or NULL (=> 1.0f, mixture.cpp:315); non-NULL for teacher-forced per-level parity
 *   level0_parent_mask [n] or NULL (=> Philox flags with counters 0..n-1)
 * Arguments 9-13 are CreateMixture's clusterLevel, hemReduction, distanceDelta,
 * colorDelta, decayRate with unchanged meaning.  Levels 1..cluster_level are
 * retained in the context (level 0 = the input, also retained) until the next
 * build or hem_destroy.  Synchronous: results are complete on return. */
int hem_build_pyramid(hem_ctx* ctx,
                      const float* xyz, const float* color, const float* opacity,
                      const float* cov6, const float* sh, const float* weight,
                      int64_t n, int32_t sh_floats,
                      uint32_t cluster_level, float hem_reduction, float distance_delta,
                      float color_delta, float decay_rate,
                      const uint8_t* level0_parent_mask);

/* Number of levels held (cluster_level + 1, level 0 = input) and size of one —
 * replaces len(result) / len(level.pointSet) on the reference's return value. */
int     hem_num_levels(const hem_ctx* ctx);
int64_t hem_level_size(const hem_ctx* ctx, int level);

/* Copy one level out in the reference's array layout — replaces
 * MixtureLevel::CreatePythonLists (src/cpp_ext/src/mixturelevel.cpp:30-70) and
 * Mixture::GetResult (src/cpp_ext/src/mixture.cpp:335-370).  Any pointer may be
 * NULL to skip that attribute.  `weight` and `is_parent` are not exported by
 * the reference API; they exist for parity checks. */
int hem_level_copy(hem_ctx* ctx, int level,
                   float* xyz, float* color, float* opacity, float* cov6, float* sh,
                   float* weight, uint8_t* is_parent);

/* "Next" row: the registration inputs the reference derives from a level on
 * the host right after the HEM step — replaces the array part of convert_gs_to_open3d_pc
 * (src/utils/point_cloud_converter.py:31-49): points [n*3] FP64, colors [n*3] FP64 = sh2rgb(DC)
 * (src/utils/graphics_utils.py:72-73), cov33 [n*9] FP64 row-major (src/utils/general_utils.py:32-40),
 * normals [n*3] FP64 = unit eigenvector of the smallest covariance eigenvalue (the reference's own
 * TODO, point_cloud_converter.py:44-47; largest component positive).  Any pointer may be NULL;
 * host or device pointers. */
int hem_level_registration_inputs(hem_ctx* ctx, int level, double* points, double* colors, double* cov33,
                                  double* normals);

/* What GaussianModel.from_mixture derives from a level's covariances on the GPU
 * (src/models/gaussian_model.py:141-153, decompose_covariance_matrix :242-265, matrices_to_quaternions
 * src/utils/general_utils.py:94-100): `scaling` [n,3] = eigenvalues after the reference's axis-correspondence
 * reorder, `rotation` [n,4] = (w,x,y,z) of the reordered eigenvector matrix, exactly as the reference writes them
 * (rows scattered, variances not standard deviations).  Optional `eigenvalues` [n,3] (ascending) and
 * `eigenvectors` [n,3,3] (columns, largest-magnitude component positive) expose the decomposition itself; the
 * reference's torch.linalg.eigh leaves the eigenvector signs to the backend, so `rotation` agrees with it up to
 * those signs.  Any pointer may be NULL; host or device memory. */
int hem_level_from_mixture(hem_ctx* ctx, int level, float* scaling, float* rotation, float* eigenvalues, float* eigenvectors);

/* "Next" row: the consumer of the pyramid.  Point-to-point ICP with the definitions of
 * the Open3D call the reference makes per level — replaces
 *   o3d.pipelines.registration.registration_icp(source, target, max_correspondence, init,
 *       TransformationEstimationPointToPoint(), ICPConvergenceCriteria(relative_fitness, relative_rmse, max_iteration))
 * as issued by do_icp_registration (src/utils/local_registration_util.py:76-100) with the default
 * LocalRegistrationParams (src/params/registration_parameters.py:8-15: point-to-point, no robust loss).
 * Open3D is a third-party dependency (pinned 0.16.0, requirements.txt:3) and is not part of the reference tree;
 * the algorithm restated is RegistrationICP of its Registration.cpp: correspondences = nearest target point of each
 * transformed source point with distance < max_correspondence; fitness = #correspondences / n_source;
 * inlier_rmse = sqrt(sum d^2 / #correspondences); update = rigid least squares (umeyama, no scale); stop when both
 * |d fitness| < relative_fitness and |d inlier_rmse| < relative_rmse, or after max_iteration updates.
 *   source [n_source*3], target [n_target*3]: FP64 points, host or device (hem_level_registration_inputs layout)
 *   init: 4x4 row-major FP64, NULL = identity (host pointer)
 *   result (host): 0..15 transformation (row-major 4x4)  16 fitness  17 inlier_rmse  18 iterations (updates applied)
 *                  19 n_correspondences  20 stop reason (1 converged, 2 max_iteration, 3 fewer than 3 correspondences)
 *                  21 ms on the device  22 grid cells  23 grid cell size
 *   correspondence [n_source] or NULL: index of the matched target point under the final transformation, -1 = none
 *                  (Open3D's correspondence_set_); host or device. */
#define HEM_ICP_RESULT_LEN 24
int hem_icp_point_to_point(hem_ctx* ctx, const double* source, int64_t n_source, const double* target, int64_t n_target,
                           const double* init, double max_correspondence, double relative_fitness, double relative_rmse,
                           int max_iteration, double* result, int32_t* correspondence);

/* Coarse-to-fine driver over two pyramids held in two contexts on the same device — replaces
 * MultiScaleRegistratorMixture._register_main_point_clouds (src/gui/workers/registration/qt_multiscale_registrator.py:203-230):
 * scale i registers level (num_levels-1-i) of the source pyramid to the same level of the target pyramid
 * (pc_list[-(i+1)]: coarsest first, level 0 = the full cloud last) with max_correspondence[i] (voxel_values) and
 * max_iteration[i] (iter_values), starting from the previous scale's transformation.  n_scales must equal
 * hem_num_levels of both contexts (the reference's own validity checks, :170-194, same messages).  The level means
 * go from the canonical rows to the FP64 search structures on the device; nothing crosses PCIe but the results:
 * results [n_scales * HEM_ICP_RESULT_LEN], one block per scale in the layout above. */
int hem_icp_multiscale(hem_ctx* source_ctx, hem_ctx* target_ctx, int n_scales, const double* max_correspondence,
                       const int32_t* max_iteration, double relative_fitness, double relative_rmse, const double* init,
                       double* results);

/* Optional copy/compute overlap for host callers: bind destination buffers (ideally pinned)
 * for level `level` (1..cluster_level) BEFORE hem_build_pyramid; the level is then copied out
 * on a second stream as soon as it is complete, while the next level is being built, and the
 * build returns with the data in place.  capacity_rows must be >= the level's size (n is always
 * enough); any pointer may be NULL.  Bindings persist until hem_clear_outputs.  (A host `sh`
 * INPUT is likewise uploaded on the second stream, under the level-0 grid and gate kernels.) */
int hem_bind_output(hem_ctx* ctx, int level, float* xyz, float* color, float* opacity, float* cov6, float* sh,
                    int64_t capacity_rows);
int hem_clear_outputs(hem_ctx* ctx);

/* Per-step statistics of clustering step `step` (input = level `step`).
 * out[0..15]:
 *   0 n_in  1 n_parents  2 n_pairs  3 n_orphans  4 n_dropped  5 n_out
 *   6 grid_cells  7 cell_size  8 max_radius
 *   9 ms_total  10 ms_prep  11 ms_grid  12 ms_gate  13 ms_sumlw  14 ms_moments  15 ms_compact
 *   16 ms_k_gate (the gate kernel alone)  17 ms_k_moments (the moment kernel alone)
 *   18 n_certified_parents (parents whose covariance is certified for the smd pre-test)
 *   19 ms_host_enqueue (host wall time spent enqueueing the step, up to its closing sync)
 *   20 n_parents_local  21 n_pairs_local (this rank's share of the gate / moment work in a multi-GPU build;
 *      equal to 1 and 2 on one GPU)  22 shards (ranks the step was sharded over; 1 = built whole on every rank)
 *   23 tight_box (1: certified parents searched the box of their Mahalanobis ellipsoid)
 * ms_* are CUDA-event times on the context's stream and are 0 unless the
 * "timing" option is set. */
#define HEM_STEP_STATS_LEN 24
int hem_step_stats(const hem_ctx* ctx, int step, double* out);

/* Cumulative counters of the context: "launches" (kernels of this library
 * launched so far), "bytes_h2d" / "bytes_d2h" (bytes staged across PCIe for
 * host-pointer arguments), "sm_count", "rank", "world". */
int hem_get_info(const hem_ctx* ctx, const char* name, double* out);

/* White-box access for parity tests (requires option "keep_pairs"): the
 * accepted (parent, child) pairs of clustering step `step`, parent-major in
 * ascending parent index; `off` has n_parents+1 entries.  Children inside a
 * parent are in this implementation's accumulation order (not sorted).  Any
 * pointer may be NULL.  The reference keeps these as locals of
 * createClusterLevel (src/cpp_ext/src/mixture.cpp:103,140) and never exports
 * them. */
int hem_step_pairs(hem_ctx* ctx, int step, uint32_t* parent_idx, int64_t* off,
                   uint32_t* child, float* wl, float* sumlw, float* radius);

/* ---- multi-GPU (one process per GPU; spatial-slab ownership of parents) ---- */

/* Size of an NCCL unique id and a getter for rank 0 to create one (the caller
 * broadcasts the bytes to the other ranks by any means, e.g. torch.distributed). */
#define HEM_NCCL_ID_BYTES 128
int hem_nccl_unique_id(uint8_t* id_out);

/* Join a communicator of `world` ranks.  After this, hem_build_pyramid must be
 * called collectively with identical arguments (identical input arrays, and the
 * same kind of pointer — host or device — for each argument) on every rank;
 * each rank runs the gate and moment kernels for the parents of its slab, the
 * accepted pairs and the merged rows are all-gathered per level, so every rank
 * ends with the complete pyramid, bit-identical to a 1-GPU run.  Host inputs are
 * uploaded in 1/world slices per rank and all-gathered (option "shard_upload"). */
int hem_comm_init(hem_ctx* ctx, const uint8_t* id, int rank, int world);

/* Pure host arithmetic, no GPU needed: the slab of parents (positions in grid-cell order)
 * [*begin, *end) that rank `rank` of `world` owns when a level has n_parents parents. */
int hem_slab_range(int64_t n_parents, int rank, int world, int64_t* begin, int64_t* end);

/* Library/ABI version (major*1000 + minor). */
int hem_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HEM_B200_H */
