// cuda_backend.h -- glue between the C++ minimizer classes and the C ABI (include/nlo_cuda.h).
// Header-only so a maintainer of the reference can drop the directory into the tree and link
// libnlo_cuda.so; nothing here touches CUDA types.
#ifndef NONLINEAR_OPTIMIZER_CUDA_BACKEND_H_
#define NONLINEAR_OPTIMIZER_CUDA_BACKEND_H_

#include <cstdint>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "nlo_cuda.h"
#include "nonlinear_optimizer/loss_function.h"
#include "nonlinear_optimizer/options.h"

namespace nonlinear_optimizer {
namespace cuda_backend {

// One context per minimizer instance, like the reference's minimizers hold their own scratch: one
// device + stream, or -- given a device list -- several B200s of this process, over which Solve splits
// the correspondence vector by point range exactly where the reference splits it over its thread pool
// (..._analytic.cc:59-73,104-119).  Minimizers are not re-entrant per instance (same contract as the
// reference).
class Session {
 public:
  explicit Session(int device = 0) : devices_(1, device) {}
  explicit Session(const std::vector<int>& devices) : devices_(devices) {}
  ~Session() {
    if (problem_ != nullptr) nlo_problem_destroy(ctx_, problem_);
    if (ctx_ != nullptr) nlo_context_destroy(ctx_);
  }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;

  bool EnsureContext() {
    if (ctx_ != nullptr) return true;
    int rc = NLO_EINVAL;
    if (devices_.size() == 1) {
      rc = nlo_context_create(devices_[0], &ctx_);
    } else if (!devices_.empty()) {
      std::vector<int32_t> dev(devices_.begin(), devices_.end());
      rc = nlo_context_create_multi(dev.data(), static_cast<int32_t>(dev.size()), &ctx_);
    }
    if (rc != NLO_OK) {
      std::cerr << "nlo_context_create (" << devices_.size() << " device(s), first "
                << (devices_.empty() ? -1 : devices_[0]) << ") failed with " << rc
                << ": no usable sm_100 GPU (there is no CPU fallback)" << std::endl;
      ctx_ = nullptr;
      return false;
    }
    return true;
  }

  // (Re)creates the device problem when the capacity is exceeded; `reproj` selects the family.
  bool EnsureProblem(int64_t n, bool reproj) {
    if (problem_ != nullptr && capacity_ >= n) return true;
    if (problem_ != nullptr) nlo_problem_destroy(ctx_, problem_);
    problem_ = nullptr;
    const int64_t cap = n + n / 4 + 256;
    const int rc = reproj ? nlo_reproj_create(ctx_, cap, &problem_) : nlo_ndt_create(ctx_, cap, &problem_);
    if (rc != NLO_OK) return Report("create problem", rc);
    capacity_ = cap;
    return true;
  }

  bool ApplyLoss(const std::shared_ptr<LossFunction>& loss) {
    int kind = NLO_LOSS_NONE;
    double params[2] = {0.0, 0.0};
    if (loss != nullptr) {
      kind = loss->DeviceKind();
      loss->DeviceParams(params);
      if (kind == LossFunction::kCustom) {
        std::cerr << "Cuda minimizer: the loss function does not describe a device functor "
                     "(override DeviceKind/DeviceParams); refusing to solve" << std::endl;
        return false;
      }
    }
    const int rc = nlo_set_loss(ctx_, kind, params);
    return rc == NLO_OK ? true : Report("nlo_set_loss", rc);
  }

  static nlo_solve_options ToC(const Options& options) {
    nlo_solve_options o;
    o.max_iterations = options.max_iterations;
    o.reserved = 0;
    o.parameter_tolerance = options.convergence_handle.parameter_tolerance;
    o.gradient_tolerance = options.convergence_handle.gradient_tolerance;
    return o;
  }

  bool Report(const char* what, int rc) {
    std::cerr << "Cuda minimizer: " << what << " failed (" << rc << "): "
              << (ctx_ ? nlo_last_error(ctx_) : "no context") << std::endl;
    return false;
  }

  nlo_context* ctx() { return ctx_; }
  nlo_problem* problem() { return problem_; }

 private:
  std::vector<int> devices_;
  nlo_context* ctx_{nullptr};
  nlo_problem* problem_{nullptr};
  int64_t capacity_{0};
};

}  // namespace cuda_backend
}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_CUDA_BACKEND_H_
