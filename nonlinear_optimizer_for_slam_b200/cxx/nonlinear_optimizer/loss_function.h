// loss_function.h -- robust loss interface of the drop-in C++ API.
//
// Source-compatible with /root/reference/nonlinear_optimizer/loss_function.h:11-77: the same class
// names, constructor arguments, exceptions (std::out_of_range on bad parameters, :24-25,:53-54)
// and the scalar `Evaluate(squared_residual, output)` contract
//     output[0] = rho(s)        output[1] = weight applied to J^T J and J^T r
// Two additions, both forced by the device path:
//   * a virtual host call cannot run inside a CUDA kernel and the reference classes keep their
//     parameters private, so every loss also DESCRIBES itself (DeviceKind / DeviceParams); the
//     kernels inline the matching functor.  A user subclass that does not override them reports
//     kCustom and the Cuda minimizers refuse it (Solve returns false) instead of silently
//     computing something else.
//   * CauchyLossFunction (BASELINE.json config 3); the reference has no Cauchy loss.
// The reference's second pure virtual, Evaluate(simd::Scalar, simd::Scalar*), needs the external
// simd_helper library; it is declared only when that header can be found and the device path
// never calls it.
#ifndef NONLINEAR_OPTIOMIZER_LOSS_FUNCTION_H_
#define NONLINEAR_OPTIOMIZER_LOSS_FUNCTION_H_

#include <cmath>
#include <stdexcept>

#if defined(__has_include)
#if __has_include(<simd_helper/simd_helper.h>)
#include <simd_helper/simd_helper.h>
#define NLO_HAVE_SIMD_HELPER 1
#endif
#endif

#ifdef NLO_HAVE_SIMD_HELPER
#define NLO_SIMD_EVALUATE_PURE \
  virtual void Evaluate(const simd::Scalar& squared_residual, simd::Scalar* output) = 0;
#define NLO_SIMD_EVALUATE_UNUSED                                                        \
  void Evaluate(const simd::Scalar& squared_residual, simd::Scalar* output) final {     \
    (void)squared_residual;                                                             \
    (void)output;                                                                       \
  }
#else
#define NLO_SIMD_EVALUATE_PURE
#define NLO_SIMD_EVALUATE_UNUSED
#endif

namespace nonlinear_optimizer {

class LossFunction {
 public:
  enum DeviceLossKind { kNone = 0, kExponential = 1, kHuber = 2, kCauchy = 3, kCustom = -1 };

  LossFunction() = default;
  virtual ~LossFunction() = default;

  virtual void Evaluate(const double squared_residual, double* output) = 0;
  NLO_SIMD_EVALUATE_PURE

  // Which device functor reproduces Evaluate(), and its (at most two) parameters.
  virtual int DeviceKind() const { return kCustom; }
  virtual void DeviceParams(double params[2]) const { params[0] = params[1] = 0.0; }
};

// rho(s) = c1 (1 - exp(-c2 s)),  weight = 2 c1 c2 exp(-c2 s)
class ExponentialLossFunction : public LossFunction {
 public:
  ExponentialLossFunction(const double c1, const double c2) : c1_(c1), c2_(c2) {
    if (c1 < 0.0) throw std::out_of_range("`c1_` should be positive number.");
    if (c2 < 0.0) throw std::out_of_range("`c2_` should be positive number.");
  }

  void Evaluate(const double squared_residual, double* output) final {
    const double decay = std::exp(-c2_ * squared_residual);
    const double weight = (2.0 * c1_ * c2_) * decay;
    output[0] = c1_ - c1_ * decay;
    output[1] = weight;
    output[2] = -2.0 * c2_ * weight;
  }
#ifdef NLO_HAVE_SIMD_HELPER
  void Evaluate(const simd::Scalar& squared_residual, simd::Scalar* output) final {
    const simd::Scalar decay = simd::exp((-c2_) * squared_residual);
    output[0] = c1_ - c1_ * decay;
    output[1] = (2.0 * c1_ * c2_) * decay;
    output[2] = (-2.0 * c2_) * output[1];
  }
#endif

  int DeviceKind() const final { return kExponential; }
  void DeviceParams(double params[2]) const final {
    params[0] = c1_;
    params[1] = c2_;
  }

 private:
  const double c1_;
  const double c2_;
};

// s <= t^2: rho = s, weight = 1;  s > t^2: rho = 2 t sqrt(s) - t^2, weight = t / sqrt(s)
class HuberLossFunction : public LossFunction {
 public:
  explicit HuberLossFunction(const double threshold) : threshold_(threshold) {
    if (threshold <= 0.0) throw std::out_of_range("threshold value should be larger than zero.");
  }

  void Evaluate(const double squared_residual, double* output) final {
    const double squared_threshold = threshold_ * threshold_;
    const bool outlier = squared_residual > squared_threshold;
    const double residual = outlier ? std::sqrt(squared_residual) : 0.0;
    output[0] = outlier ? 2.0 * threshold_ * residual - squared_threshold : squared_residual;
    output[1] = outlier ? threshold_ / residual : 1.0;
  }
  NLO_SIMD_EVALUATE_UNUSED  // the reference's SIMD overload is an empty stub too (:68-72)

  int DeviceKind() const final { return kHuber; }
  void DeviceParams(double params[2]) const final {
    params[0] = threshold_;
    params[1] = 0.0;
  }

 private:
  const double threshold_;
};

// rho(s) = c^2 log(1 + s / c^2),  weight = rho'(s) = 1 / (1 + s / c^2)   (Ceres' CauchyLoss)
class CauchyLossFunction : public LossFunction {
 public:
  explicit CauchyLossFunction(const double c) : c_(c) {
    if (c <= 0.0) throw std::out_of_range("c should be larger than zero.");
  }

  void Evaluate(const double squared_residual, double* output) final {
    const double u = squared_residual / (c_ * c_);
    output[0] = (c_ * c_) * std::log1p(u);
    output[1] = 1.0 / (1.0 + u);
  }
  NLO_SIMD_EVALUATE_UNUSED

  int DeviceKind() const final { return kCauchy; }
  void DeviceParams(double params[2]) const final {
    params[0] = c_;
    params[1] = 0.0;
  }

 private:
  const double c_;
};

}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIOMIZER_LOSS_FUNCTION_H_
