// reprojection_error_minimizer_cuda.h -- B200 implementation of ReprojectionErrorMinimizer.
// Replaces ReprojectionErrorMinimizerAnalytic / ...SIMD
// (reference: reprojection_error_minimizer_analytic.cc:12-105, ..._analytic_simd.cc:14-198).
// Solve() contract and differences: see mahalanobis_distance_minimizer_cuda.h.
#ifndef NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_
#define NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_

#include <cstddef>
#include <iostream>
#include <vector>

#include "nonlinear_optimizer/cuda_backend.h"
#include "nonlinear_optimizer/reprojection_error_minimizer/reprojection_error_minimizer.h"

namespace nonlinear_optimizer {
namespace reprojection_error_minimizer {

class ReprojectionErrorMinimizerCuda : public ReprojectionErrorMinimizer {
 public:
  explicit ReprojectionErrorMinimizerCuda(int device = 0) : session_(device) {}
  explicit ReprojectionErrorMinimizerCuda(const std::vector<int>& devices) : session_(devices) {}

  bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
             const CameraIntrinsics& camera_intrinsics, Pose* pose) final {
    if (pose == nullptr || !session_.EnsureContext()) return false;
    if (!session_.ApplyLoss(loss_function_)) return false;
    const int64_t n = static_cast<int64_t>(correspondences.size());
    if (!session_.EnsureProblem(n, /*reproj=*/true)) return false;
    // the reference's 40-byte records are ingested in place (no host-side repack)
    const double K[6] = {camera_intrinsics.fx, camera_intrinsics.fy, camera_intrinsics.cx,
                         camera_intrinsics.cy, camera_intrinsics.inv_fx, camera_intrinsics.inv_fy};
    int rc = nlo_reproj_upload_aos(session_.ctx(), session_.problem(), n, correspondences.data(),
                                   sizeof(Correspondence), offsetof(Correspondence, local_point),
                                   offsetof(Correspondence, matched_pixel), K);
    if (rc != NLO_OK) return session_.Report("nlo_reproj_upload_aos", rc);
    const nlo_solve_options o = cuda_backend::Session::ToC(options);
    nlo_solve_result result{};
    rc = nlo_reproj_solve(session_.ctx(), session_.problem(), &o, PoseData(*pose), &result, nullptr);
    last_result_ = result;
    if (rc != NLO_OK) return session_.Report("nlo_reproj_solve", rc);
    std::cerr << "COST: " << result.final_cost << ", iter: " << result.iterations << std::endl;
    return true;
  }

  const nlo_solve_result& last_result() const { return last_result_; }

 private:
  cuda_backend::Session session_;
  nlo_solve_result last_result_{};
};

}  // namespace reprojection_error_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_CUDA_H_
