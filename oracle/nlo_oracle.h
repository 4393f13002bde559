/*
 * nlo_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A dependency-free (no Eigen, no simd_helper) double-precision restatement of the
 * reference's scalar Gauss-Newton / damped-LM pose minimizers.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (libnlo_cuda.so) never links or calls it.
 *
 * Parity pin: the PnP known-answer test of the reference
 *   results/reproj_amd64.txt:5   "COST: 2.33228e-11, iter: 6"
 *   results/reproj_amd64.txt:10  pose^-1 = (-0.1, 0.123, -0.5 | 0 0 0.0499792 0.99875)
 * is reproduced by nlo_oracle_reproj_solve on the fixture of
 *   reprojection_error_minimizer/tests/simple_optimization_test.cc:42-71,115-158
 * (tests/test_oracle_golden.py).  The NDT logs (results/maha_*.txt) depend on Eigen's
 * eigenvector sign/order and unordered_map iteration order, so they pin the oracle only
 * to a +-0.5 % cost band and the outer-iteration count.
 *
 * The reference itself cannot be compiled in this image (needs Eigen3, Ceres, flann and
 * the external simd_helper; none installed, no network) -> oracle/_ref is not built.
 *
 * Array conventions (all double):
 *   point[n*3], mean[n*3]        interleaved xyz per correspondence
 *   sqrt_info[n*9]               ROW-major 3x3 per correspondence (S(i,j) at 3*i+j)
 *   local_point[n*3], pixel[n*2] PnP correspondence
 *   pose[16]                     4x4 homogeneous, COLUMN-major (Eigen::Isometry3d memory)
 *   H                            upper triangle, row-major order:
 *                                6-DoF: 21 values (0,0),(0,1)..(0,5),(1,1)..(5,5)
 *                                3-DoF:  6 values (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
 *   trace row (per iteration)    6-DoF/PnP: H21 | g6 | cost | t3 | q4(x,y,z,w) | lambda  = 36
 *                                3-DoF    : H6  | g3 | cost | t2 | R2 (row-major 4) | lambda = 17
 *                                (pose and lambda are the values AFTER the iteration's update)
 */
#ifndef NLO_ORACLE_H_
#define NLO_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NLO_ORACLE_LOSS_NONE = 0, NLO_ORACLE_LOSS_EXPONENTIAL = 1, NLO_ORACLE_LOSS_HUBER = 2,
       NLO_ORACLE_LOSS_CAUCHY = 3 };

#define NLO_ORACLE_TRACE6 36
#define NLO_ORACLE_TRACE3 17

/* loss_function.h:28-33 (Exponential), :57-66 (Huber); Cauchy is an addition (Ceres convention). */
void nlo_oracle_loss(int kind, const double params[2], double squared_residual, double out[3]);

/* Rotation helpers restating the Eigen operations the reference relies on. */
void nlo_oracle_rotmat_to_quat(const double R_rowmajor[9], double q_xyzw[4]);
void nlo_oracle_quat_to_rotmat(const double q_xyzw[4], double R_rowmajor[9]);
/* mahalanobis_distance_minimizer.cc:20-33 / reprojection_error_minimizer.h:35-52 */
void nlo_oracle_compute_quaternion(const double w[3], double q_xyzw[4]);

/* ---- NDT / Mahalanobis, 6-DoF: mahalanobis_distance_minimizer_analytic.cc ---- */
/* :159-185 */
void nlo_oracle_ndt6_jacobian_residual(const double R_rowmajor[9], const double t[3],
                                       const double point[3], const double mean[3],
                                       const double sqrt_info[9], double J_rowmajor_3x6[18],
                                       double r[3]);
/* :12-52 over [begin,end).  long_double_accum != 0 sums in long double (a "truth" for tests). */
void nlo_oracle_ndt6_assemble(int64_t begin, int64_t end, const double* point, const double* mean,
                              const double* sqrt_info, const double R_rowmajor[9],
                              const double t[3], int loss_kind, const double loss_params[2],
                              int long_double_accum, double H21[21], double g[6], double* cost);
/* :54-157.  num_threads == 0: single range; > 0: the executor split of :59-73,104-119
 * (chunks of max(1, N/T), tail dropped, partials summed in thread order; real std::threads). */
int nlo_oracle_ndt6_solve(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, int loss_kind, const double loss_params[2],
                          int max_iterations, double parameter_tolerance,
                          double gradient_tolerance, int num_threads, double pose[16],
                          int* iterations, double* final_cost, double* trace);

/* ---- NDT / Mahalanobis, 3-DoF planar: mahalanobis_distance_minimizer_analytic_3dof.cc ---- */
/* :110-139 */
void nlo_oracle_ndt3_jacobian_residual(const double R2_rowmajor[4], const double t2[2],
                                       const double point[3], const double mean[3],
                                       const double sqrt_info[9], double J_rowmajor_3x3[9],
                                       double r[3]);
/* :33-68 (the caller applies the floor(N/4)*4 truncation of :33-36) */
void nlo_oracle_ndt3_assemble(int64_t begin, int64_t end, const double* point, const double* mean,
                              const double* sqrt_info, const double R2_rowmajor[4],
                              const double t2[2], int loss_kind, const double loss_params[2],
                              int long_double_accum, double H6[6], double g[3], double* cost);
/* :14-108 */
int nlo_oracle_ndt3_solve(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, int loss_kind, const double loss_params[2],
                          int max_iterations, double parameter_tolerance,
                          double gradient_tolerance, double pose[16], int* iterations,
                          double* final_cost, double* trace);

/* ---- Reprojection error (PnP): reprojection_error_minimizer_analytic.cc ---- */
/* intrinsics = {fx, fy, cx, cy, inv_fx, inv_fy}   (types.h:17-26) */
/* :107-162 */
void nlo_oracle_reproj_jacobian_residual(const double R_rowmajor[9], const double t[3],
                                         const double local_point[3], const double pixel[2],
                                         const double intrinsics[6], double J_rowmajor_2x6[12],
                                         double r[2]);
/* :31-63 */
void nlo_oracle_reproj_assemble(int64_t begin, int64_t end, const double* local_point,
                                const double* pixel, const double intrinsics[6],
                                const double R_rowmajor[9], const double t[3], int loss_kind,
                                const double loss_params[2], int long_double_accum,
                                double H21[21], double g[6], double* cost);
/* :12-105 */
int nlo_oracle_reproj_solve(int64_t n, const double* local_point, const double* pixel,
                            const double intrinsics[6], int loss_kind,
                            const double loss_params[2], int max_iterations,
                            double parameter_tolerance, double gradient_tolerance,
                            double pose[16], int* iterations, double* final_cost, double* trace);

/* One damped step on already-reduced sums (shared by the solves above and by the multi-rank
 * gloo tests): reflect, H(k,k)*=1+lambda, delta=-H^-1 g, pose update, convergence tests,
 * lambda schedule -- mahalanobis_distance_minimizer_analytic.cc:122-148.
 * state6 = t3 | q4(x,y,z,w) | lambda | previous_cost  (9 doubles).  Returns 1 if converged
 * (break taken; lambda/previous_cost then untouched, as in the reference). */
int nlo_oracle_gn6_step(const double H21[21], const double g[6], double cost,
                        double parameter_tolerance, double gradient_tolerance, double state6[9]);
/* 3-DoF twin (mahalanobis_distance_minimizer_analytic_3dof.cc:69-99);
 * state3 = t2 | R2 row-major 4 | lambda | previous_cost (8 doubles). */
int nlo_oracle_gn3_step(const double H6[6], const double g[3], double cost,
                        double parameter_tolerance, double gradient_tolerance, double state3[8]);


/* ---- CPU timing baselines (nlo_oracle_simd.cc) ---- */
/* AoS double -> 15 float planes (x y z | mx my mz | s00..s22), plane k at planes + k*n;
 * the per-Solve conversion of ..._analytic_simd_various.cc:1252-1266. */
void nlo_oracle_simd_pack(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, float* planes);
/* float 8-lane AVX2+FMA assembly (= SolveFloatIntrinsicAligned, ..._simd_various.cc:1300-1447)
 * over floor(n/8)*8 correspondences, thread split of ..._analytic_simd.cc:55-76. */
void nlo_oracle_simd_ndt6_assemble(int64_t n, const float* planes, const double R_rowmajor[9],
                                   const double t[3], int loss_kind, const double loss_params[2],
                                   int num_threads, double H21[21], double g[6], double* cost);
/* float 8-lane twin of the planar minimizer (..._analytic_3dof_simd.cc:85-157) over floor(n/8)*8
 * correspondences of the same 15 float planes.  The reference runs it on ONE thread; num_threads > 1
 * applies the stride split of ..._analytic_simd.cc:55-76. */
void nlo_oracle_simd_ndt3_assemble(int64_t n, const float* planes, const double R2_rowmajor[4],
                                   const double t2[2], int loss_kind, const double loss_params[2],
                                   int num_threads, double H6[6], double g[3], double* cost);
/* float 8-lane twin of the reprojection minimizer
 * (reprojection_error_minimizer/reprojection_error_minimizer_analytic_simd.cc:19-27 pack, :55-137 loop),
 * with its quirks: ||r|| (not squared) goes to the loss, the gate is z > 0.  One thread in the
 * reference; num_threads > 1 as above. */
void nlo_oracle_simd_reproj_pack(int64_t n, const double* local_point, const double* pixel, float* planes);
void nlo_oracle_simd_reproj_assemble(int64_t n, const float* planes, const double intrinsics[6],
                                     const double R_rowmajor[9], const double t[3], int loss_kind,
                                     const double loss_params[2], int num_threads, double H21[21],
                                     double g[6], double* cost);
/* scalar-double assembly on `num_threads` threads (..._analytic.cc:59-73,104-119). */
void nlo_oracle_ndt6_assemble_threads(int64_t n, const double* point, const double* mean,
                                      const double* sqrt_info, const double R_rowmajor[9],
                                      const double t[3], int loss_kind,
                                      const double loss_params[2], int num_threads,
                                      double H21[21], double g[6], double* cost);

/* ---- Fixtures restated from the reference's test mains ---- */
/* reprojection_error_minimizer/tests/simple_optimization_test.cc:115-135 (630 points);
 * returns the count, fills at most `capacity` points. */
int64_t nlo_oracle_pnp_reference_points(double* xyz, int64_t capacity);
/* mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:170-204 (954 605 points) */
int64_t nlo_oracle_room_points(double* xyz, int64_t capacity);
/* :282-294  zig-zag fold + Cantor pairing twice */
uint64_t nlo_oracle_voxel_key(const double point[3], double inverse_voxel_resolution);

#ifdef __cplusplus
}
#endif
#endif /* NLO_ORACLE_H_ */
