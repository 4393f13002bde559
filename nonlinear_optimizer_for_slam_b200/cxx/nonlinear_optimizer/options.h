// options.h -- solver options of the drop-in C++ API.
//
// Source-compatible with /root/reference/nonlinear_optimizer/options.h:6-28: the same field names,
// defaults and enum values, so `options.convergence_handle.parameter_tolerance = ...` written for
// the reference compiles unchanged.  As in the reference, the Solve() bodies read only
// max_iterations, convergence_handle.parameter_tolerance and convergence_handle.gradient_tolerance;
// minimizer_type, linear_solver_type, function_tolerance and optimization_handle are accepted and
// ignored (the lambda bounds 1e-6 / 1e-2 are constants inside each Solve,
// mahalanobis_distance_minimizer_analytic.cc:81-82).
#ifndef NONLINEAR_OPTIMIZER_OPTIONS_H_
#define NONLINEAR_OPTIMIZER_OPTIONS_H_

namespace nonlinear_optimizer {

enum class MinimizerType : int {
  kGaussNewton = 0,
  kGradientDescent = 1,
  kQuasiNewton = 2,
  kLevenbergMarquardt = 3,
};

enum class LinearSolverType : int {
  kDenseQR = 0,
  kDenseCholesky = 1,
  kSparseCholesky = 2,
};

struct ConvergenceHandle {
  double function_tolerance = 1e-6;   // unused by every Solve, kept for source compatibility
  double gradient_tolerance = 1e-6;   // stop when ||J^T W r|| falls below this
  double parameter_tolerance = 1e-6;  // stop when ||step|| falls below this
};

struct OptimizationHandle {
  double min_lambda = 1e-6;  // documented only: the solvers hard-code the same bounds
  double max_lambda = 1e-2;
};

struct Options {
  int max_iterations = 40;
  MinimizerType minimizer_type = MinimizerType::kGaussNewton;
  LinearSolverType linear_solver_type = LinearSolverType::kDenseQR;
  ConvergenceHandle convergence_handle;
  OptimizationHandle optimization_handle;
};

}  // namespace nonlinear_optimizer

#endif  // NONLINEAR_OPTIMIZER_OPTIONS_H_
