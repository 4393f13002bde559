// types.h -- value types of the drop-in C++ API.
//
// Mirrors /root/reference/nonlinear_optimizer/types.h:8-58.  With Eigen on the include path these
// ARE the reference's Eigen typedefs; without it (this build image has no Eigen) a minimal,
// layout-compatible stand-in is provided so the API layer and its tests still compile: column-major
// fixed-size storage, exactly Eigen's memory layout for Matrix<double,N,M> and Isometry3d.
#ifndef NONLINEAR_OPTIMIZER_TYPES_H_
#define NONLINEAR_OPTIMIZER_TYPES_H_

#if defined(NLO_USE_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Dense>))
#include <Eigen/Dense>
namespace nonlinear_optimizer {
using Vec2 = Eigen::Matrix<double, 2, 1>;
using Vec3 = Eigen::Matrix<double, 3, 1>;
using Vec6 = Eigen::Matrix<double, 6, 1>;
using Mat2x2 = Eigen::Matrix<double, 2, 2>;
using Mat3x3 = Eigen::Matrix<double, 3, 3>;
using Mat6x6 = Eigen::Matrix<double, 6, 6>;
using Orientation = Eigen::Quaterniond;
using Pose2 = Eigen::Isometry2d;
using Pose = Eigen::Isometry3d;
inline const double* PoseData(const Pose& pose) { return pose.matrix().data(); }
inline double* PoseData(Pose& pose) { return pose.matrix().data(); }
}  // namespace nonlinear_optimizer
#define NLO_HAVE_EIGEN 1
#else
#include <cmath>
#include <cstddef>
namespace nonlinear_optimizer {

template <int Rows, int Cols>
struct Mat {
  double data_[Rows * Cols];  // column-major, as Eigen
  Mat() { for (double& v : data_) v = 0.0; }
  static Mat Zero() { return Mat(); }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < (Rows < Cols ? Rows : Cols); ++i) m(i, i) = 1.0;
    return m;
  }
  double& operator()(int r, int c) { return data_[c * Rows + r]; }
  double operator()(int r, int c) const { return data_[c * Rows + r]; }
  double& operator()(int i) { return data_[i]; }
  double operator()(int i) const { return data_[i]; }
  double* data() { return data_; }
  const double* data() const { return data_; }
  double& x() { return data_[0]; }
  double& y() { return data_[1]; }
  double& z() { return data_[2]; }
  double x() const { return data_[0]; }
  double y() const { return data_[1]; }
  double z() const { return data_[2]; }
};
using Vec2 = Mat<2, 1>;
using Vec3 = Mat<3, 1>;
using Vec6 = Mat<6, 1>;
using Mat2x2 = Mat<2, 2>;
using Mat3x3 = Mat<3, 3>;
using Mat6x6 = Mat<6, 6>;

inline Vec3 MakeVec3(double x, double y, double z) {
  Vec3 v;
  v(0) = x; v(1) = y; v(2) = z;
  return v;
}
inline Vec2 MakeVec2(double x, double y) {
  Vec2 v;
  v(0) = x; v(1) = y;
  return v;
}

// Stand-in for Eigen::Isometry3d: a column-major 4x4 with the accessors the minimizers' callers use.
struct Pose {
  double m_[16];
  Pose() { SetIdentity(); }
  static Pose Identity() { return Pose(); }
  void SetIdentity() {
    for (double& v : m_) v = 0.0;
    m_[0] = m_[5] = m_[10] = m_[15] = 1.0;
  }
  double& operator()(int r, int c) { return m_[4 * c + r]; }
  double operator()(int r, int c) const { return m_[4 * c + r]; }
  Vec3 translation() const { return MakeVec3(m_[12], m_[13], m_[14]); }
  void set_translation(const Vec3& t) { m_[12] = t(0); m_[13] = t(1); m_[14] = t(2); }
  Mat3x3 linear() const {
    Mat3x3 R;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R(r, c) = (*this)(r, c);
    return R;
  }
  Mat3x3 rotation() const { return linear(); }
  void set_linear(const Mat3x3& R) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) (*this)(r, c) = R(r, c);
  }
  Vec3 operator*(const Vec3& p) const {
    Vec3 out;
    for (int r = 0; r < 3; ++r)
      out(r) = (*this)(r, 0) * p(0) + (*this)(r, 1) * p(1) + (*this)(r, 2) * p(2) + (*this)(r, 3);
    return out;
  }
  Pose inverse() const {  // rigid inverse
    Pose out;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out(r, c) = (*this)(c, r);
    for (int r = 0; r < 3; ++r)
      out(r, 3) = -(out(r, 0) * m_[12] + out(r, 1) * m_[13] + out(r, 2) * m_[14]);
    return out;
  }
  static Pose FromYaw(double yaw, const Vec3& t) {
    Pose p;
    p(0, 0) = std::cos(yaw); p(0, 1) = -std::sin(yaw);
    p(1, 0) = std::sin(yaw); p(1, 1) = std::cos(yaw);
    p.set_translation(t);
    return p;
  }
};
inline const double* PoseData(const Pose& pose) { return pose.m_; }
inline double* PoseData(Pose& pose) { return pose.m_; }

}  // namespace nonlinear_optimizer
#define NLO_HAVE_EIGEN 0
#endif

#endif  // NONLINEAR_OPTIMIZER_TYPES_H_
