// ndt_registration_cuda.h -- the whole scan-to-map registration on the device.
//
// The reference has no class for this: its test mains (mahalanobis_distance_minimizer/tests/
// simple_optimization_test.cc) build the NDT map (UpdateNdtMap :236-280), match with a flann
// KD-tree (MatchPointCloud :296-342) and loop match + Solve up to 10 times
// (OptimizePoseAnalytic :473-505) on the host.  Those steps are inside every timing the reference
// publishes, so they are offered here as one call that keeps scan, map, correspondences and the
// Gauss-Newton loop resident on the GPU:
//
//   NdtRegistrationCuda reg;                          // device 0
//   reg.SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
//   reg.BuildMap(global_points, /*voxel=*/1.0);       // or SetMap(...) with host-built cells
//   reg.SetScan(local_points);
//   reg.Register(options, &pose);                     // pose: in = initial guess, out = result
#ifndef NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_NDT_REGISTRATION_CUDA_H_
#define NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_NDT_REGISTRATION_CUDA_H_

#include <iostream>
#include <vector>

#include "nonlinear_optimizer/cuda_backend.h"
#include "nonlinear_optimizer/mahalanobis_distance_minimizer/types.h"

namespace nonlinear_optimizer {
namespace mahalanobis_distance_minimizer {

class NdtRegistrationCuda {
 public:
  explicit NdtRegistrationCuda(int device = 0) : session_(device) {}
  ~NdtRegistrationCuda() {
    if (scan_ != nullptr) nlo_scan_destroy(session_.ctx(), scan_);
    if (map_ != nullptr) nlo_ndt_map_destroy(session_.ctx(), map_);
  }
  NdtRegistrationCuda(const NdtRegistrationCuda&) = delete;
  NdtRegistrationCuda& operator=(const NdtRegistrationCuda&) = delete;

  void SetLossFunction(const std::shared_ptr<LossFunction>& loss_function) { loss_function_ = loss_function; }

  // UpdateNdtMap on the device.  reference_literal_sqrt_information selects the reference's
  // `diag * V` form (:275-276) instead of `diag * V^T`.  voxel_hash keeps only the occupied voxels
  // in a device hash table (the reference's unordered_map, :282-294) instead of a dense grid over
  // the bounding box; maps whose box exceeds 2^28 voxels take that form by themselves.
  bool BuildMap(const std::vector<Vec3>& points, double voxel_resolution,
                bool reference_literal_sqrt_information = false, bool voxel_hash = false) {
    if (!session_.EnsureContext()) return false;
    Flatten(points);
    if (map_ != nullptr) nlo_ndt_map_destroy(session_.ctx(), map_);
    map_ = nullptr;
    const auto build = voxel_hash ? nlo_ndt_map_build_hashed : nlo_ndt_map_build;
    const int rc = build(session_.ctx(), static_cast<int64_t>(points.size()), flat_.data(), voxel_resolution,
                         reference_literal_sqrt_information ? 1 : 0, &map_);
    return rc == NLO_OK ? true : session_.Report("nlo_ndt_map_build", rc);
  }

  bool SetScan(const std::vector<Vec3>& local_points) {
    if (!session_.EnsureContext()) return false;
    Flatten(local_points);
    if (scan_ != nullptr) nlo_scan_destroy(session_.ctx(), scan_);
    scan_ = nullptr;
    const int rc = nlo_scan_create(session_.ctx(), static_cast<int64_t>(local_points.size()), flat_.data(), &scan_);
    return rc == NLO_OK ? true : session_.Report("nlo_scan_create", rc);
  }

  // <= max_outer x { match (<= max_neighbors nearest means within radius), Solve }.
  bool Register(const Options& options, Pose* pose, bool planar_3dof = false, double radius = 1.0,
                int max_neighbors = 2, int max_outer = 10) {
    if (pose == nullptr || map_ == nullptr || scan_ == nullptr) return false;
    if (!session_.ApplyLoss(loss_function_)) return false;
    const nlo_solve_options o = cuda_backend::Session::ToC(options);
    const int rc = nlo_ndt_register(session_.ctx(), scan_, map_, &o, radius, max_neighbors, max_outer,
                                    planar_3dof ? 1 : 0, PoseData(*pose), &last_result_);
    std::cerr << "COST: " << last_result_.final_cost << ", iter: " << last_result_.inner_iterations
              << " (outer " << last_result_.outer_iterations << ")" << std::endl;
    return rc == NLO_OK ? true : session_.Report("nlo_ndt_register", rc);
  }

  const nlo_register_result& last_result() const { return last_result_; }

 private:
  void Flatten(const std::vector<Vec3>& pts) {
    flat_.resize(3 * pts.size());
    for (size_t i = 0; i < pts.size(); ++i) {
      flat_[3 * i] = pts[i](0);
      flat_[3 * i + 1] = pts[i](1);
      flat_[3 * i + 2] = pts[i](2);
    }
  }

  cuda_backend::Session session_;
  std::shared_ptr<LossFunction> loss_function_{nullptr};
  nlo_ndt_map* map_{nullptr};
  nlo_scan* scan_{nullptr};
  std::vector<double> flat_;
  nlo_register_result last_result_{};
};

}  // namespace mahalanobis_distance_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_NDT_REGISTRATION_CUDA_H_
