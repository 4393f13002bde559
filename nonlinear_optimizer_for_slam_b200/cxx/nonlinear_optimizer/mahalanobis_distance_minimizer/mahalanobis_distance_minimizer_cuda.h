// mahalanobis_distance_minimizer_cuda.h -- B200 implementations of MahalanobisDistanceMinimizer.
//
//   MahalanobisDistanceMinimizerCuda      replaces MahalanobisDistanceMinimizerAnalytic / ...SIMD
//       (reference: mahalanobis_distance_minimizer_analytic.cc:54-157, ..._analytic_simd.cc:16-111)
//   MahalanobisDistanceMinimizerCuda3DOF  replaces MahalanobisDistanceMinimizerAnalytic3DOF / ...SIMD
//       (reference: mahalanobis_distance_minimizer_analytic_3dof.cc:14-108)
//
// Same Solve() contract: *pose is the initial guess on entry and the optimum on return (the 3-DoF
// class rewrites only x, y and the top-left 2x2), one "COST: <c>, iter: <k>" line goes to stderr
// (..._analytic.cc:150), the call blocks until the result is on the host.  Differences:
//   * returns false instead of true when the device path fails (no GPU, CUDA error, non-finite H)
//     or when the loss cannot be expressed on the device; the reference always returns true.
//   * the executor is ignored, so the `N mod num_threads` tail the reference's threaded path drops
//     is processed (result = the reference's single-thread result).
//   * constructed with a device LIST, Solve splits the correspondences by point range over those
//     B200s (one process, peer-mapped exchange buffers, one all-reduce of the 28 / 10 sums per
//     iteration inside the iteration kernel) -- the place where the reference splits them over its
//     thread pool (..._analytic.cc:59-73,104-119).  The pose is bit-identical to the 1-GPU one only up
//     to the summation order of the shards (1e-12 relative on H, g).
#ifndef NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_CUDA_H_
#define NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_CUDA_H_

#include <cstddef>
#include <iostream>
#include <vector>

#include "nonlinear_optimizer/cuda_backend.h"
#include "nonlinear_optimizer/mahalanobis_distance_minimizer/mahalanobis_distance_minimizer.h"

namespace nonlinear_optimizer {
namespace mahalanobis_distance_minimizer {

namespace internal {

// Uploads the reference's AoS records in place: stride sizeof(Correspondence), Eigen (column-major)
// sqrt_information.  This replaces the per-Solve AoS->SoA conversion of ..._analytic_simd.cc:19-28.
inline bool UploadCorrespondences(cuda_backend::Session* session,
                                  const std::vector<Correspondence>& correspondences) {
  const int64_t n = static_cast<int64_t>(correspondences.size());
  if (!session->EnsureProblem(n, /*reproj=*/false)) return false;
  const size_t off_point = offsetof(Correspondence, point);
  const size_t off_mean = offsetof(Correspondence, ndt) + offsetof(NDT, mean);
  const size_t off_sqrt = offsetof(Correspondence, ndt) + offsetof(NDT, sqrt_information);
  const int rc = nlo_ndt_upload_aos(session->ctx(), session->problem(), n, correspondences.data(),
                                    sizeof(Correspondence), off_point, off_mean, off_sqrt,
                                    /*sqrt_info_col_major=*/1);
  return rc == NLO_OK ? true : session->Report("nlo_ndt_upload_aos", rc);
}

}  // namespace internal

class MahalanobisDistanceMinimizerCuda : public MahalanobisDistanceMinimizer {
 public:
  explicit MahalanobisDistanceMinimizerCuda(int device = 0) : session_(device) {}
  explicit MahalanobisDistanceMinimizerCuda(const std::vector<int>& devices) : session_(devices) {}

  bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
             Pose* pose) final {
    if (pose == nullptr || !session_.EnsureContext()) return false;
    if (!session_.ApplyLoss(loss_function_)) return false;
    if (!internal::UploadCorrespondences(&session_, correspondences)) return false;
    const nlo_solve_options o = cuda_backend::Session::ToC(options);
    nlo_solve_result result{};
    const int rc = nlo_ndt6_solve(session_.ctx(), session_.problem(), &o, PoseData(*pose), &result,
                                  nullptr);
    last_result_ = result;
    if (rc != NLO_OK) return session_.Report("nlo_ndt6_solve", rc);
    std::cerr << "COST: " << result.final_cost << ", iter: " << result.iterations << std::endl;
    return true;
  }

  const nlo_solve_result& last_result() const { return last_result_; }
  // wall / host-gather milliseconds of the last Solve's ingest (nlo_ingest_stats)
  void last_ingest_ms(double* total_ms, double* host_gather_ms) { nlo_ingest_stats(session_.ctx(), total_ms, host_gather_ms); }

 private:
  cuda_backend::Session session_;
  nlo_solve_result last_result_{};
};

class MahalanobisDistanceMinimizerCuda3DOF : public MahalanobisDistanceMinimizer {
 public:
  explicit MahalanobisDistanceMinimizerCuda3DOF(int device = 0) : session_(device) {}
  explicit MahalanobisDistanceMinimizerCuda3DOF(const std::vector<int>& devices) : session_(devices) {}

  bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
             Pose* pose) final {
    if (pose == nullptr || !session_.EnsureContext()) return false;
    if (!session_.ApplyLoss(loss_function_)) return false;
    if (!internal::UploadCorrespondences(&session_, correspondences)) return false;
    const nlo_solve_options o = cuda_backend::Session::ToC(options);
    nlo_solve_result result{};
    const int rc = nlo_ndt3_solve(session_.ctx(), session_.problem(), &o, PoseData(*pose), &result,
                                  nullptr);
    last_result_ = result;
    if (rc != NLO_OK) return session_.Report("nlo_ndt3_solve", rc);
    std::cerr << "COST: " << result.final_cost << ", iter: " << result.iterations << std::endl;
    return true;
  }

  const nlo_solve_result& last_result() const { return last_result_; }
  // wall / host-gather milliseconds of the last Solve's ingest (nlo_ingest_stats)
  void last_ingest_ms(double* total_ms, double* host_gather_ms) { nlo_ingest_stats(session_.ctx(), total_ms, host_gather_ms); }

 private:
  cuda_backend::Session session_;
  nlo_solve_result last_result_{};
};

}  // namespace mahalanobis_distance_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_CUDA_H_
