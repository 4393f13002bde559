// types.h -- camera intrinsics and 3D-2D correspondence of the reprojection-error minimizer.
// Field-for-field /root/reference/nonlinear_optimizer/reprojection_error_minimizer/types.h:14-28.
#ifndef NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_
#define NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_

#include <cstddef>

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

// Fills every field of CameraIntrinsics, the reciprocals included, from the four pinhole numbers.
inline CameraIntrinsics MakeCameraIntrinsics(double fx, double fy, double cx, double cy, int width,
                                             int height) {
  CameraIntrinsics k;
  k.fx = fx;
  k.fy = fy;
  k.cx = cx;
  k.cy = cy;
  k.inv_fx = 1.0 / fx;
  k.inv_fy = 1.0 / fy;
  k.width = width;
  k.height = height;
  return k;
}

// Splits an array of correspondences into the two flat arrays nlo_reproj_upload takes
// (xyz triples and uv pairs); `points` and `pixels` must hold 3 * n and 2 * n doubles.
inline void FlattenCorrespondences(const Correspondence* records, std::size_t n, double* points,
                                   double* pixels) {
  for (std::size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) points[3 * i + k] = records[i].local_point(k);
    for (int k = 0; k < 2; ++k) pixels[2 * i + k] = records[i].matched_pixel(k);
  }
}

}  // namespace reprojection_error_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_REPROJECTION_ERROR_MINIMIZER_TYPES_H_
