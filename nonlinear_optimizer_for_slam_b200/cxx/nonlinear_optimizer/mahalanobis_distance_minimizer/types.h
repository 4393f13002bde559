// types.h -- NDT cell and correspondence records of the Mahalanobis-distance minimizers.
//
// Field-for-field the records of
// /root/reference/nonlinear_optimizer/mahalanobis_distance_minimizer/types.h:11-26, so a
// std::vector<Correspondence> built for the reference can be handed to the Cuda minimizers.  The
// hot path reads point (3), ndt.mean (3) and ndt.sqrt_information (9): 15 of the record's 38
// scalars; nlo_ndt_upload_aos() gathers exactly those on the device.
#ifndef NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_TYPES_H_
#define NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_TYPES_H_

#include "nonlinear_optimizer/types.h"

namespace nonlinear_optimizer {
namespace mahalanobis_distance_minimizer {

struct NDT {
  int count{0};                                 // points accumulated in the voxel
  Vec3 sum{Vec3::Zero()};                       // sum of points
  Mat3x3 moment{Mat3x3::Identity()};            // sum of p p^T (starts at Identity, as upstream)
  Vec3 mean{Vec3::Zero()};                      // HOT: cell mean
  Mat3x3 information{Mat3x3::Identity()};       // S^T S
  Mat3x3 sqrt_information{Mat3x3::Identity()};  // HOT: S, residual = S (R p + t - mean)
  bool is_valid{false};
  bool is_planar{false};
};

struct Correspondence {
  Vec3 point{Vec3::Zero()};  // HOT: scan point in the sensor frame
  NDT ndt;                   // the matched cell (a full copy per correspondence, as upstream)
};

}  // namespace mahalanobis_distance_minimizer
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_MAHALANOBIS_DISTANCE_MINIMIZER_TYPES_H_
