/*
 * nlo_cuda.h -- C ABI of libnlo_cuda.so: the B200 (sm_100a) Gauss-Newton / damped-LM
 * normal-equation assembly + device-resident iteration loop for the Mahalanobis/NDT (6-DoF and
 * 3-DoF planar) and reprojection-error pose minimizers.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its interface for this
 * path is two C++ abstract classes (all citations relative to /root/reference/nonlinear_optimizer/)
 *   mahalanobis_distance_minimizer/mahalanobis_distance_minimizer.h:31-33   Solve(Options, vector<Correspondence>, Pose*)
 *   reprojection_error_minimizer/reprojection_error_minimizer.h:30-32       Solve(Options, vector<Correspondence>, CameraIntrinsics, Pose*)
 *   .../mahalanobis_distance_minimizer.h:29, reprojection_error_minimizer.h:20-22   SetLossFunction
 * The C++ classes under nonlinear_optimizer_for_slam_b200/cxx/ keep those signatures and call
 * the functions below; nothing here exposes C++, Eigen, torch or CUDA types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative NLO_E* code on failure, never throws;
 *     nlo_last_error(ctx) gives the message of the last failure on that context.
 *   - all floating point data is IEEE double.  pose[16] is a 4x4 homogeneous transform in
 *     COLUMN-major order (the memory of Eigen::Isometry3d, types.h:31).
 *   - host arrays use the reference's record order: point[3n] xyz interleaved, mean[3n],
 *     sqrt_info[9n] ROW-major 3x3 per correspondence (S(i,j) at 9k+3i+j), local_point[3n],
 *     pixel[2n].  On the device they are repacked into a tile-interleaved SoA layout (DESIGN.md
 *     section 2): 256 correspondences per tile, the planes of a tile back to back; an NDT record
 *     becomes 12 doubles (point 3, mean 3, the 6 unique entries of S^T S formed once at ingest).
 *   - H is the packed upper triangle in row-major order: 6-DoF 21 values
 *     (0,0),(0,1)..(0,5),(1,1)..(5,5); 3-DoF 6 values.  g is J^T W r (6 or 3).
 *   - a context owns one CUDA device + one stream (nlo_context_create) or several devices of this
 *     process (nlo_context_create_multi); calls on one context are serialised by the caller (same
 *     contract as the reference minimizers: not re-entrant per instance).
 *   - there is no CPU fallback: without a usable CUDA device nlo_context_create fails.
 */
#ifndef NLO_CUDA_H_
#define NLO_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLO_ABI_VERSION 2

#if defined(__GNUC__)
#define NLO_API __attribute__((visibility("default")))
#else
#define NLO_API
#endif

enum {
  NLO_OK = 0,
  NLO_EINVAL = -1,   /* bad argument */
  NLO_ECUDA = -2,    /* CUDA runtime error (message in nlo_last_error) */
  NLO_ENOMEM = -3,   /* allocation failed */
  NLO_ECOMM = -4,    /* NCCL / peer-memory communicator error */
  NLO_ENUMERIC = -5, /* non-finite H, g or step met during a solve */
  NLO_ETIMEOUT = -6  /* a grid-wide wait of the persistent iteration kernel expired (GPU shared / preempted) */
};

/* loss_function.h:20-77.  params: EXPONENTIAL {c1, c2}; HUBER {threshold}; CAUCHY {c}
 * (Cauchy is an addition, Ceres convention rho = c^2 log(1+s/c^2)); NONE = plain least squares
 * (the reference's loss_function_ == nullptr branch, ..._analytic.cc:44-48). */
enum { NLO_LOSS_NONE = 0, NLO_LOSS_EXPONENTIAL = 1, NLO_LOSS_HUBER = 2, NLO_LOSS_CAUCHY = 3 };

/* options.h:15-28 -- only the fields the reference's Solve() bodies read. */
typedef struct nlo_solve_options {
  int32_t max_iterations;     /* Options::max_iterations, default 40 */
  int32_t reserved;
  double parameter_tolerance; /* convergence_handle.parameter_tolerance, default 1e-6 */
  double gradient_tolerance;  /* convergence_handle.gradient_tolerance, default 1e-6 */
} nlo_solve_options;

typedef struct nlo_solve_result {
  int32_t iterations; /* the reference's `iteration` when the loop ends (printed as "iter:") */
  int32_t status;     /* 0 ok, NLO_ENUMERIC if a non-finite value stopped the loop */
  double final_cost;  /* the reference's `previous_cost` (printed as "COST:") */
  double device_ms;   /* CUDA-event time of the device-resident loop on the context stream */
} nlo_solve_result;

/* Row width of the optional per-iteration trace (tests only):
 *   6-DoF / reprojection: H21 | g6 | cost | t3 | q4(x,y,z,w) | lambda          = 36
 *   3-DoF               : H6  | g3 | cost | t2 | R2 (row-major 4) | lambda     = 17
 * pose and lambda are the values AFTER that iteration's update. */
#define NLO_TRACE6 36
#define NLO_TRACE3 17

typedef struct nlo_context nlo_context;
typedef struct nlo_problem nlo_problem;

/* ---- context ---- */
NLO_API int nlo_abi_version(void);
/* CUDA devices visible to this process (0 without a driver / GPU). */
NLO_API int nlo_visible_device_count(void);
NLO_API int nlo_context_create(int device, nlo_context** ctx);
/* One context over `num_devices` (<= 8) devices of THIS process -- the reference splits the
 * correspondence vector over its thread pool inside Solve (..._analytic.cc:59-73,104-119,
 * ..._analytic_simd.cc:55-76); here the split is over B200s, behind the same calls.  A single problem is
 * sharded by contiguous point range at upload / generate, every device runs the iteration kernel on
 * its shard, the 28 (10) doubles are summed each iteration by the one-shot all-reduce over directly
 * mapped peer memory (NVLink), and every device applies the identical step; the 3-DoF solve drops
 * the last n mod 4 correspondences of the WHOLE list, as the reference.  A batched problem is
 * partitioned by registration id, no exchange.  Every call of this header except nlo_comm_* works on
 * such a context (maps, scans and nlo_ndt_register run on its first device). */
NLO_API int nlo_context_create_multi(const int32_t* devices, int32_t num_devices, nlo_context** ctx);
NLO_API int nlo_context_device_count(const nlo_context* ctx);
NLO_API int nlo_context_destroy(nlo_context* ctx);
NLO_API const char* nlo_last_error(const nlo_context* ctx);
/* sm_count, and the grid (CTAs) a single-problem assembly launch uses */
NLO_API int nlo_context_info(const nlo_context* ctx, int* sm_count, int* assemble_grid);
NLO_API int nlo_synchronize(nlo_context* ctx);
/* SetLossFunction equivalent (…minimizer.h:29).  Applies to later assemble/solve calls. */
NLO_API int nlo_set_loss(nlo_context* ctx, int kind, const double params[2]);

/* pinned host memory for callers that want full-rate uploads */
NLO_API int nlo_host_alloc(void** ptr, size_t bytes);
NLO_API int nlo_host_free(void* ptr);

/* Memory-safety check of the library's own device buffers, for boxes where compute-sanitizer is not
 * available.  With NLO_GUARD=1 in the environment when the library is first used, every device
 * allocation sits between two 64 KB guard bands of 0xFF bytes and starts out filled with 0xFF itself
 * (NaN as doubles): a write past either end of a buffer shows up here as corrupted guard bytes (bands
 * are compared when a buffer is freed and, for the live ones, by this call), a read of memory nothing
 * wrote poisons the result.  NLO_GUARD=selftest additionally overwrites one guard byte of the first
 * allocation on purpose (the report must then say 1).  Without NLO_GUARD: *enabled = 0, zero counts. */
NLO_API int nlo_debug_guard_report(int32_t* enabled, int64_t* allocations_checked, int64_t* allocations_live,
                                   int64_t* corrupted_bytes);

/* ---- NDT / Mahalanobis correspondences (types.h:11-26) ---- */
/* One problem of up to `capacity` correspondences. */
NLO_API int nlo_ndt_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem);
/* Same, with the correspondences STORED as float on the device (48 instead of 96 bytes each: point,
 * mean and the 6 unique entries of S^T S, formed in fp64 and rounded once) and all arithmetic still
 * in fp64 -- float storage as in the reference's SIMD minimizers (..._analytic_simd.cc:19-28,
 * SOAData).  Results equal the fp64 path run on the float-rounded records (nlo_ndt_download
 * returns them); against the unrounded double inputs they differ by that quantisation (~1e-7
 * relative), so this is an opt-in throughput mode, not the parity mode.  Supported: upload, generate, download,
 * ndt6/ndt3 assemble and solve (single problem, also sharded over GPUs). */
NLO_API int nlo_ndt_create_f32(nlo_context* ctx, int64_t capacity, nlo_problem** problem);
/* Upload into an fp32-storage problem from FLOAT host arrays (same record order as nlo_ndt_upload):
 * half the PCIe bytes of the double upload. */
NLO_API int nlo_ndt_upload_f32(nlo_context* ctx, nlo_problem* problem, int64_t n, const float* point,
                       const float* mean, const float* sqrt_info);
/* `num_problems` independent registrations; counts[k] correspondences each (BASELINE cfg5). */
NLO_API int nlo_ndt_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts,
                           nlo_problem** problem);
/* Host -> device.  For a batched problem the arrays are the concatenation in problem order and
 * n must equal the sum of counts. */
NLO_API int nlo_ndt_upload(nlo_context* ctx, nlo_problem* problem, int64_t n, const double* point,
                   const double* mean, const double* sqrt_info);
/* Ingest the reference's AoS records in place (std::vector<Correspondence>::data()):
 * `stride` bytes between records, byte offsets of point (3 doubles), mean (3 doubles) and
 * sqrt_information (9 doubles, column-major if sqrt_info_col_major != 0 as Eigen stores it). */
NLO_API int nlo_ndt_upload_aos(nlo_context* ctx, nlo_problem* problem, int64_t n, const void* records,
                       size_t stride, size_t offset_point, size_t offset_mean,
                       size_t offset_sqrt_info, int sqrt_info_col_major);
/* Wall time of the last nlo_*_upload* call on the context and the part of it the calling thread spent
 * in the host-side gather (ms).  Uploads are chunked and pipelined: host threads gather the 15 (5) hot
 * doubles of each record into a ring of pinned chunks while the previous chunk crosses PCIe and is
 * repacked on the device; the caller's memory may be pageable; device staging is O(chunk). */
NLO_API int nlo_ingest_stats(const nlo_context* ctx, double* total_ms, double* host_gather_ms);
/* Device-side synthetic correspondences (bench / large-scale tests; no host copy):
 * point i = counter-based PRNG(seed, global_index_offset + i) on the surfaces of the 7x5x2.5 m
 * room of tests/simple_optimization_test.cc:170-204 expressed in the sensor frame of
 * true_pose[16], associated with the cell of the dense voxel grid that contains the point under
 * init_pose[16].  grid: origin[3], dims[3], voxel size, cell_mean[3*cells], cell_sqrt_info[9*cells]
 * (row-major), cell_valid[cells] -- host arrays.  Points whose cell is invalid are resampled. */
NLO_API int nlo_ndt_generate(nlo_context* ctx, nlo_problem* problem, int64_t n, uint64_t seed,
                     int64_t global_index_offset, double noise_sigma, const double true_pose[16],
                     const double init_pose[16], const double grid_origin[3],
                     const int32_t grid_dims[3], double voxel_size, const double* cell_mean,
                     const double* cell_sqrt_info, const uint8_t* cell_valid);
/* Batched twin (BASELINE cfg5): registration k gets counts[k] points from stream seed + k in the
 * sensor frame of true_poses[16*k .. 16*k+15]; all are associated under the same init_pose. */
NLO_API int nlo_ndt_generate_batched(nlo_context* ctx, nlo_problem* problem, uint64_t seed,
                             double noise_sigma, const double* true_poses,
                             const double init_pose[16], const double grid_origin[3],
                             const int32_t grid_dims[3], double voxel_size,
                             const double* cell_mean, const double* cell_sqrt_info,
                             const uint8_t* cell_valid);
/* Device -> host copy of correspondences [begin, end) (tests): point[3n], mean[3n] and the 6
 * unique entries of the information matrix S^T S per correspondence (00 01 02 11 12 22) -- the
 * device keeps S only through S^T S, which is all both NDT minimizers use. */
NLO_API int nlo_ndt_download(nlo_context* ctx, const nlo_problem* problem, int64_t begin, int64_t end,
                     double* point, double* mean, double* information);
/* Same for registration `problem_index` of a batched problem (0 for a single one). */
NLO_API int nlo_ndt_download_problem(nlo_context* ctx, const nlo_problem* problem, int32_t problem_index,
                             int64_t begin, int64_t end, double* point, double* mean, double* information);

/* ---- reprojection correspondences (reprojection_error_minimizer/types.h:14-28) ---- */
NLO_API int nlo_reproj_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem);
/* `num_problems` independent PnP problems sharing one camera (the realistic multi-GPU use of the
 * reprojection minimizer: sets are small, so registrations are batched and sharded by problem). */
NLO_API int nlo_reproj_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts,
                              nlo_problem** problem);
/* intrinsics = {fx, fy, cx, cy, inv_fx, inv_fy}.  For a batched problem the arrays are the
 * concatenation in problem order and n must equal the sum of counts. */
NLO_API int nlo_reproj_upload(nlo_context* ctx, nlo_problem* problem, int64_t n,
                      const double* local_point, const double* pixel,
                      const double intrinsics[6]);

/* Ingest the reference's AoS records in place (std::vector<Correspondence>::data(), types.h:14-17):
 * byte offsets of local_point (3 doubles) and matched_pixel_point (2 doubles). */
NLO_API int nlo_reproj_upload_aos(nlo_context* ctx, nlo_problem* problem, int64_t n, const void* records,
                          size_t stride, size_t offset_local_point, size_t offset_pixel,
                          const double intrinsics[6]);

NLO_API int nlo_problem_destroy(nlo_context* ctx, nlo_problem* problem);
NLO_API int64_t nlo_problem_size(const nlo_problem* problem);

/* ---- one assembly pass (parity tests, benchmarks) ----
 * H, g, cost of correspondences [begin, end) at `pose` under the context's loss; with a
 * communicator attached the sums are all-reduced over ranks.
 *   ndt6   : ..._analytic.cc:12-52 + :159-185        H21, g6
 *   ndt3   : ..._analytic_3dof.cc:33-68 + :110-139   H6,  g3 (caller chooses end; Solve uses floor(n/4)*4)
 *   reproj : reprojection_error_minimizer_analytic.cc:31-63 + :107-162   H21, g6
 * For a batched problem, `problem_index` selects the registration; otherwise pass 0. */
NLO_API int nlo_ndt6_assemble(nlo_context* ctx, nlo_problem* problem, int32_t problem_index,
                      const double pose[16], int64_t begin, int64_t end, double H21[21],
                      double g[6], double* cost);
NLO_API int nlo_ndt3_assemble(nlo_context* ctx, nlo_problem* problem, int32_t problem_index,
                      const double pose[16], int64_t begin, int64_t end, double H6[6],
                      double g[3], double* cost);
NLO_API int nlo_reproj_assemble(nlo_context* ctx, nlo_problem* problem, int32_t problem_index,
                        const double pose[16], int64_t begin, int64_t end, double H21[21],
                        double g[6], double* cost);

/* ---- device-resident solves ----
 * pose: in = initial guess, out = optimized pose (3-DoF writes only x,y and the 2x2 block, as
 * ..._analytic_3dof.cc:104-105).  trace may be NULL, else max_iterations * NLO_TRACE{6,3}
 * doubles.  Returns NLO_OK also when the iteration cap is hit (the reference always returns
 * true); NLO_ENUMERIC if a non-finite value was met.
 *   ndt6   : ..._analytic.cc:54-157
 *   ndt3   : ..._analytic_3dof.cc:14-108
 *   reproj : reprojection_error_minimizer_analytic.cc:12-105 */
NLO_API int nlo_ndt6_solve(nlo_context* ctx, nlo_problem* problem, const nlo_solve_options* options,
                   double pose[16], nlo_solve_result* result, double* trace);
NLO_API int nlo_ndt3_solve(nlo_context* ctx, nlo_problem* problem, const nlo_solve_options* options,
                   double pose[16], nlo_solve_result* result, double* trace);
NLO_API int nlo_reproj_solve(nlo_context* ctx, nlo_problem* problem, const nlo_solve_options* options,
                     double pose[16], nlo_solve_result* result, double* trace);
/* All registrations of a batched problem in one launch: poses[16*num_problems] in/out,
 * results[num_problems].  No collective: problems are independent. */
NLO_API int nlo_ndt6_solve_batched(nlo_context* ctx, nlo_problem* problem,
                           const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results);

/* Planar and reprojection twins of nlo_ndt6_solve_batched (the 3-DoF one applies the reference's
 * floor(n/4)*4 truncation to every registration). */
NLO_API int nlo_ndt3_solve_batched(nlo_context* ctx, nlo_problem* problem,
                           const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results);
NLO_API int nlo_reproj_solve_batched(nlo_context* ctx, nlo_problem* problem,
                             const nlo_solve_options* options, double* poses,
                             nlo_solve_result* results);

/* ---- next rows of the scope table: device NDT map, matcher and the outer registration loop ----
 * (the reference keeps these in its test mains, mahalanobis_distance_minimizer/tests/
 *  simple_optimization_test.cc; they are inside every timing it publishes) */
typedef struct nlo_ndt_map nlo_ndt_map;
typedef struct nlo_scan nlo_scan;

/* A dense-voxel-grid NDT map from host tables: cell (x,y,z) at (z*dims[1]+y)*dims[0]+x covers
 * [origin + idx*voxel, +voxel); cell_mean[3*cells], cell_sqrt_info[9*cells] row-major,
 * cell_valid[cells]. */
NLO_API int nlo_ndt_map_create(nlo_context* ctx, const double grid_origin[3], const int32_t grid_dims[3],
                       double voxel_size, const double* cell_mean, const double* cell_sqrt_info,
                       const uint8_t* cell_valid, nlo_ndt_map** map);
/* UpdateNdtMap (:236-280) on the device from n host points (xyz interleaved): per-voxel
 * count/sum/moment, count >= 5, cov = (I + sum p p^T)/n - mean mean^T, symmetric 3x3
 * eigen-decomposition, reject lambda_max < 0.01, clamp the two small eigenvalues to
 * 0.01 lambda_max, sqrt_information = diag(lambda^-1/2) V^T -- or diag * V, exactly as the
 * reference writes it (:275-276), when v_not_transposed != 0 (that form depends on the arbitrary
 * eigenvector signs; here each eigenvector's largest component is made positive). */
NLO_API int nlo_ndt_map_build(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size,
                      int v_not_transposed, nlo_ndt_map** map);
/* nlo_ndt_map_build keeps a dense grid over the points' bounding box and switches to the voxel
 * hash by itself when that box has more than 2^28 voxels; nlo_ndt_map_build_hashed always builds
 * the sparse form.  The reference holds its map in a std::unordered_map keyed by the voxel indices
 * (:282-294); the device form is an open-addressing table (linear probing, at most half full) over
 * the occupied voxels only, so memory follows the occupancy rather than the bounding box.  Matching
 * against either form returns the same correspondences. */
NLO_API int nlo_ndt_map_build_hashed(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size,
                             int v_not_transposed, nlo_ndt_map** map);
/* *hashed = 1 for the sparse form; *cells = rows of the tables nlo_ndt_map_download fills (grid
 * cells when dense, hash slots when sparse). */
NLO_API int nlo_ndt_map_layout(nlo_context* ctx, const nlo_ndt_map* map, int32_t* hashed, int64_t* cells);
/* Sparse maps only: the key of every slot, x | y << 21 | z << 42 with voxel indices relative to
 * grid_origin (nlo_ndt_map_info), or UINT64_MAX for a free slot. */
NLO_API int nlo_ndt_map_download_keys(nlo_context* ctx, const nlo_ndt_map* map, uint64_t* slot_keys);
/* dims/origin/voxel/cells of a map; then the tables (arrays sized from the first call). */
NLO_API int nlo_ndt_map_info(nlo_context* ctx, const nlo_ndt_map* map, double grid_origin[3],
                     int32_t grid_dims[3], double* voxel_size, int64_t* valid_cells);
NLO_API int nlo_ndt_map_download(nlo_context* ctx, const nlo_ndt_map* map, double* cell_mean,
                         double* cell_sqrt_info, uint8_t* cell_valid);
NLO_API int nlo_ndt_map_destroy(nlo_context* ctx, nlo_ndt_map* map);

/* A scan (points in the sensor frame) resident on the device. */
NLO_API int nlo_scan_create(nlo_context* ctx, int64_t n, const double* points_xyz, nlo_scan** scan);
NLO_API int nlo_scan_destroy(nlo_context* ctx, nlo_scan* scan);

/* MatchPointCloud (:296-342): for every scan point warped by `pose`, the <= max_neighbors (1 or 2)
 * nearest valid cell means with squared distance < radius^2; fills `problem` (an nlo_ndt_create'd
 * problem of capacity >= max_neighbors * n) with max_neighbors * n correspondences -- neighbour
 * slot j of point i at index j*n + i, missing neighbours as zero-information records (they add
 * exactly nothing).  *matched (nullable) = number of real correspondences. */
NLO_API int nlo_ndt_match(nlo_context* ctx, const nlo_scan* scan, const nlo_ndt_map* map, const double pose[16],
                  double radius, int32_t max_neighbors, nlo_problem* problem, int64_t* matched);

typedef struct nlo_register_result {
  int32_t outer_iterations;  /* match + solve rounds executed */
  int32_t inner_iterations;  /* sum of the solves' `iterations` */
  int32_t status;
  int32_t reserved;
  double final_cost;         /* of the last solve that iterated */
  double device_ms;          /* CUDA-event time of all match + solve work */
  int64_t matched;           /* correspondences of the last match */
} nlo_register_result;

/* The outer loop of OptimizePoseAnalytic (:473-505): up to max_outer x { match at the current
 * pose, Solve }, stopping when |dt| < 1e-5 and |dq.vec| < 1e-5 between rounds (:495-499).
 * three_dof != 0 uses the planar solver (3dof_6dof_comparison_test.cc).  Everything runs on the
 * device; the host sees one pose per round. */
NLO_API int nlo_ndt_register(nlo_context* ctx, const nlo_scan* scan, const nlo_ndt_map* map,
                     const nlo_solve_options* options, double radius, int32_t max_neighbors,
                     int32_t max_outer, int32_t three_dof, double pose[16],
                     nlo_register_result* result);

/* ---- multi-GPU, one process per GPU (a large scan sharded by point range; for several GPUs inside
 * one process see nlo_context_create_multi) ----
 * With a communicator attached, every assemble/solve on the context sums its 28 (10 for 3-DoF)
 * partial doubles over all ranks each iteration and every rank applies the identical update.
 * NCCL flavour: rank 0 calls nlo_comm_unique_id and ships the 128 bytes to the others. */
NLO_API int nlo_comm_unique_id(nlo_context* ctx, uint8_t id[128]);
NLO_API int nlo_comm_init_nccl(nlo_context* ctx, const uint8_t id[128], int32_t rank, int32_t nranks);
/* Peer-memory flavour (NVLink/NVSwitch one-shot all-reduce fused into the iteration kernel):
 * every rank exports a 64-byte handle, gathers all of them (rank order) and opens the peers. */
NLO_API int nlo_comm_peer_export(nlo_context* ctx, uint8_t handle[64]);
NLO_API int nlo_comm_peer_init(nlo_context* ctx, const uint8_t* handles, int32_t rank, int32_t nranks);
/* This problem holds correspondences [global_begin, global_begin + n) of a scan of global_total that is
 * sharded over ranks.  Only the planar solve needs to know: the reference drops the last
 * global_total mod 4 correspondences of the WHOLE list (..._analytic_3dof.cc:33-36), so nlo_ndt3_solve
 * cuts this shard at floor(global_total / 4) * 4 - global_begin instead of at floor(n / 4) * 4.
 * global_total < 0 clears the setting.  (A multi-device context does this by itself.) */
NLO_API int nlo_problem_set_global_range(nlo_context* ctx, nlo_problem* problem, int64_t global_begin,
                                 int64_t global_total);
/* suspended != 0: later assemble / solve calls behave as if no communicator were attached (this
 * rank's own, un-reduced sums); 0 restores the collective.  Parity checks use it to compare the
 * all-reduced sums with the per-rank ones. */
NLO_API int nlo_comm_suspend(nlo_context* ctx, int32_t suspended);
NLO_API int nlo_comm_destroy(nlo_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* NLO_CUDA_H_ */
