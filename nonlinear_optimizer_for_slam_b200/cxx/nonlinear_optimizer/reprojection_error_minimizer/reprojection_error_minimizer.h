// reprojection_error_minimizer.h -- abstract base of the PnP pose minimizers.
// Same public interface as
// /root/reference/nonlinear_optimizer/reprojection_error_minimizer/reprojection_error_minimizer.h:14-32.
#ifndef NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_H_
#define NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_H_

#include <memory>
#include <vector>

#include "nonlinear_optimizer/loss_function.h"
#include "nonlinear_optimizer/options.h"
#include "nonlinear_optimizer/reprojection_error_minimizer/types.h"

namespace nonlinear_optimizer {
namespace reprojection_error_minimizer {

class ReprojectionErrorMinimizer {
 public:
  ReprojectionErrorMinimizer() = default;
  virtual ~ReprojectionErrorMinimizer() = default;

  void SetLossFunction(const std::shared_ptr<LossFunction>& loss_function) {
    loss_function_ = loss_function;
  }

  // pose: transform from the reference frame to the query frame (T_qr); in = initial guess.
  virtual bool Solve(const Options& options, const std::vector<Correspondence>& correspondences,
                     const CameraIntrinsics& camera_intrinsics, Pose* pose) = 0;

 protected:
  std::shared_ptr<LossFunction> loss_function_{nullptr};
};

}  // namespace reprojection_error_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_REPROJECTION_ERROR_MINIMIZER_H_
