// mahalanobis_distance_minimizer.h -- abstract base of the NDT / Mahalanobis pose minimizers.
//
// Same public interface as
// /root/reference/nonlinear_optimizer/mahalanobis_distance_minimizer/mahalanobis_distance_minimizer.h:20-42
// (Solve, SetLossFunction, SetMultiThreadExecutor).  The protected Eigen helpers of the reference
// (ComputeQuaternion, PartialResult) belong to its CPU implementations and are not part of the
// boundary; their device twins live in csrc/nlo_device.cuh.
#ifndef NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_H_
#define NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_H_

#include <memory>
#include <vector>

#include "nonlinear_optimizer/loss_function.h"
#include "nonlinear_optimizer/mahalanobis_distance_minimizer/types.h"
#include "nonlinear_optimizer/multi_thread_executor.h"
#include "nonlinear_optimizer/options.h"

namespace nonlinear_optimizer {
namespace mahalanobis_distance_minimizer {

class MahalanobisDistanceMinimizer {
 public:
  MahalanobisDistanceMinimizer() = default;
  virtual ~MahalanobisDistanceMinimizer() = default;

  // Accepted for source compatibility; the device path has no use for a CPU thread pool.
  void SetMultiThreadExecutor(const std::shared_ptr<MultiThreadExecutor>& multi_thread_executor) {
    multi_thread_executor_ = multi_thread_executor;
  }

  void SetLossFunction(const std::shared_ptr<LossFunction>& loss_function) {
    loss_function_ = loss_function;
  }

  // pose: in = initial guess, out = optimized pose.  Returns true on success.
  virtual bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
                     Pose* pose) = 0;

 protected:
  std::shared_ptr<LossFunction> loss_function_{nullptr};
  std::shared_ptr<MultiThreadExecutor> multi_thread_executor_{nullptr};
};

}  // namespace mahalanobis_distance_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_H_
