// reprojection_error_minimizer_cuda.h -- B200 implementation of ReprojectionErrorMinimizer.
// Replaces ReprojectionErrorMinimizerAnalytic / ...SIMD
// (reference: reprojection_error_minimizer_analytic.cc:12-105, ..._analytic_simd.cc:14-198).
// Solve() contract and differences: see mahalanobis_distance_minimizer_cuda.h.
#ifndef NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_
#define NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_

#include <iostream>
#include <vector>

#include "nonlinear_optimizer/cuda_backend.h"
#include "nonlinear_optimizer/reprojection_error_minimizer/reprojection_error_minimizer.h"

namespace nonlinear_optimizer {
namespace reprojection_error_minimizer {

class ReprojectionErrorMinimizerCuda : public ReprojectionErrorMinimizer {
 public:
  explicit ReprojectionErrorMinimizerCuda(int device = 0) : session_(device) {}

  bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
             const CameraIntrinsics& camera_intrinsics, Pose* pose) final {
    if (pose == nullptr || !session_.EnsureContext()) return false;
    if (!session_.ApplyLoss(loss_function_)) return false;
    const int64_t n = static_cast<int64_t>(correspondences.size());
    if (!session_.EnsureProblem(n, /*reproj=*/true)) return false;
    // 40-byte records: split into the two host arrays of the C ABI (the only host-side repack;
    // 5 doubles per correspondence)
    points_.resize(3 * correspondences.size());
    pixels_.resize(2 * correspondences.size());
    FlattenCorrespondences(correspondences.data(), correspondences.size(), points_.data(),
                           pixels_.data());
    const double K[6] = {camera_intrinsics.fx, camera_intrinsics.fy, camera_intrinsics.cx,
                         camera_intrinsics.cy, camera_intrinsics.inv_fx, camera_intrinsics.inv_fy};
    int rc = nlo_reproj_upload(session_.ctx(), session_.problem(), n, points_.data(), pixels_.data(), K);
    if (rc != NLO_OK) return session_.Report("nlo_reproj_upload", rc);
    const nlo_solve_options o = cuda_backend::Session::ToC(options);
    nlo_solve_result result;
    rc = nlo_reproj_solve(session_.ctx(), session_.problem(), &o, PoseData(*pose), &result, nullptr);
    std::cerr << "COST: " << result.final_cost << ", iter: " << result.iterations << std::endl;
    last_result_ = result;
    return rc == NLO_OK ? true : session_.Report("nlo_reproj_solve", rc);
  }

  const nlo_solve_result& last_result() const { return last_result_; }

 private:
  cuda_backend::Session session_;
  std::vector<double> points_, pixels_;
  nlo_solve_result last_result_{};
};

}  // namespace reprojection_error_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_
