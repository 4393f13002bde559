// types.h -- camera intrinsics and 3D-2D correspondence of the reprojection-error minimizer.
// Field-for-field /root/reference/nonlinear_optimizer/reprojection_error_minimizer/types.h:14-28.
#ifndef NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_
#define NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_

#include "nonlinear_optimizer/types.h"

namespace nonlinear_optimizer {
namespace reprojection_error_minimizer {

// Pinhole model of an undistorted (rectified) image.
struct CameraIntrinsics {
  double fx{0.0};
  double fy{0.0};
  double cx{0.0};
  double cy{0.0};
  double inv_fx{0.0};  // the caller sets 1 / fx (the analytic minimizer reads it, ..._analytic.cc:127)
  double inv_fy{0.0};
  int width{0};
  int height{0};
};

struct Correspondence {
  Vec3 local_point{Vec3::Zero()};    // 3-D point in the reference frame
  Vec2 matched_pixel{Vec2::Zero()};  // its observation in the query image
};

}  // namespace reprojection_error_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_
