/*
 * nlo_oracle.cc -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See nlo_oracle.h.
 *
 * Dependency-free double-precision restatement of the reference's scalar minimizers.
 * Every function cites the reference file:line (relative to
 * /root/reference/nonlinear_optimizer/) it follows.  The Eigen operations the reference leans
 * on (Quaterniond(Matrix3d), toRotationMatrix, quaternion product, normalize, fixed-size
 * inverse, Isometry2d::rotate) are restated from their textbook definitions.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product library never links or calls it.
 */
#include "nlo_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

/* ---- tiny fixed-size helpers (row-major) ---- */
inline void MatVec3(const double M[9], const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2];
}

/* C = A * B, 3x3 row-major */
inline void MatMat3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

inline void Skew(const double v[3], double S[9]) {
  S[0] = 0.0;   S[1] = -v[2]; S[2] = v[1];
  S[3] = v[2];  S[4] = 0.0;   S[5] = -v[0];
  S[6] = -v[1]; S[7] = v[0];  S[8] = 0.0;
}

/* Hamilton product a*b, (x,y,z,w) storage -- Eigen::Quaternion::operator* */
inline void QuatMul(const double a[4], const double b[4], double out[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
  const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
  out[3] = aw * bw - ax * bx - ay * by - az * bz;
  out[0] = aw * bx + ax * bw + ay * bz - az * by;
  out[1] = aw * by + ay * bw + az * bx - ax * bz;
  out[2] = aw * bz + az * bw + ax * by - ay * bx;
}

inline void QuatNormalize(double q[4]) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

/* General n x n inverse times vector by Gauss-Jordan elimination with partial pivoting
 * (Eigen's Matrix<double,6,6>::inverse() is a partial-pivot LU; same arithmetic class). */
template <int N>
inline void SolveDense(const double A_in[N * N], const double rhs[N], double x[N]) {
  double A[N][N + 1];
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) A[i][j] = A_in[N * i + j];
    A[i][N] = rhs[i];
  }
  for (int c = 0; c < N; ++c) {
    int piv = c;
    double best = std::fabs(A[c][c]);
    for (int r = c + 1; r < N; ++r)
      if (std::fabs(A[r][c]) > best) { best = std::fabs(A[r][c]); piv = r; }
    if (piv != c)
      for (int j = 0; j <= N; ++j) std::swap(A[c][j], A[piv][j]);
    const double inv = 1.0 / A[c][c];
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = A[r][c] * inv;
      if (f == 0.0) continue;
      for (int j = c; j <= N; ++j) A[r][j] -= f * A[c][j];
    }
  }
  for (int i = 0; i < N; ++i) x[i] = A[i][N] / A[i][i];
}

/* Index of (row,col), row<=col, in the packed row-major upper triangle of an n x n matrix */
inline int UpIdx(int n, int row, int col) { return row * n - row * (row - 1) / 2 + (col - row); }

template <int N>
inline void Unpack(const double* Hup, double H[N * N]) {
  int k = 0;
  for (int r = 0; r < N; ++r)
    for (int c = r; c < N; ++c) {
      H[N * r + c] = Hup[k];
      H[N * c + r] = Hup[k]; /* ReflectHessian: ..._analytic.cc:220-227 */
      ++k;
    }
}

/* Accumulator that is either double or long double. */
template <typename Acc, int NH, int NG>
struct Sums {
  Acc H[NH];
  Acc g[NG];
  Acc cost;
  Sums() { for (auto& v : H) v = 0; for (auto& v : g) v = 0; cost = 0; }
};

/* Shared per-correspondence accumulation: ..._analytic.cc:27-47 (and the 3dof / reprojection
 * twins ..._analytic_3dof.cc:46-67, reprojection_error_minimizer_analytic.cc:40-62).
 * J is ROWS x DIM row-major. */
template <typename Acc, int ROWS, int DIM, int NH>
inline void Accumulate(const double* J, const double* r, int loss_kind, const double loss_params[2],
                       Sums<Acc, NH, DIM>* sums) {
  double local_gradient[DIM];
  for (int c = 0; c < DIM; ++c) {
    double s = 0.0;
    for (int k = 0; k < ROWS; ++k) s += J[DIM * k + c] * r[k];
    local_gradient[c] = s;
  }
  double local_hessian[NH];
  {
    int idx = 0;
    for (int row = 0; row < DIM; ++row)
      for (int col = row; col < DIM; ++col) {
        double s = 0.0;
        for (int k = 0; k < ROWS; ++k) s += J[DIM * k + row] * J[DIM * k + col];
        local_hessian[idx++] = s;
      }
  }
  double squared_residual = 0.0;
  for (int k = 0; k < ROWS; ++k) squared_residual += r[k] * r[k];
  if (loss_kind != NLO_ORACLE_LOSS_NONE) {
    double loss_output[3] = {0.0, 0.0, 0.0};
    nlo_oracle_loss(loss_kind, loss_params, squared_residual, loss_output);
    const double weight = loss_output[1];
    for (int c = 0; c < DIM; ++c) sums->g[c] += weight * local_gradient[c];
    for (int k = 0; k < NH; ++k) sums->H[k] += weight * local_hessian[k];
    sums->cost += loss_output[0];
  } else {
    for (int c = 0; c < DIM; ++c) sums->g[c] += local_gradient[c];
    for (int k = 0; k < NH; ++k) sums->H[k] += local_hessian[k];
    sums->cost += squared_residual;
  }
}

template <typename Acc>
void Ndt6Range(int64_t begin, int64_t end, const double* point, const double* mean,
               const double* sqrt_info, const double R[9], const double t[3], int loss_kind,
               const double loss_params[2], double H21[21], double g[6], double* cost) {
  Sums<Acc, 21, 6> sums;
  double J[18], r[3];
  for (int64_t i = begin; i < end; ++i) {
    nlo_oracle_ndt6_jacobian_residual(R, t, point + 3 * i, mean + 3 * i, sqrt_info + 9 * i, J, r);
    Accumulate<Acc, 3, 6, 21>(J, r, loss_kind, loss_params, &sums);
  }
  for (int k = 0; k < 21; ++k) H21[k] = static_cast<double>(sums.H[k]);
  for (int k = 0; k < 6; ++k) g[k] = static_cast<double>(sums.g[k]);
  *cost = static_cast<double>(sums.cost);
}

template <typename Acc>
void Ndt3Range(int64_t begin, int64_t end, const double* point, const double* mean,
               const double* sqrt_info, const double R2[4], const double t2[2], int loss_kind,
               const double loss_params[2], double H6[6], double g[3], double* cost) {
  Sums<Acc, 6, 3> sums;
  double J[9], r[3];
  for (int64_t i = begin; i < end; ++i) {
    nlo_oracle_ndt3_jacobian_residual(R2, t2, point + 3 * i, mean + 3 * i, sqrt_info + 9 * i, J, r);
    Accumulate<Acc, 3, 3, 6>(J, r, loss_kind, loss_params, &sums);
  }
  for (int k = 0; k < 6; ++k) H6[k] = static_cast<double>(sums.H[k]);
  for (int k = 0; k < 3; ++k) g[k] = static_cast<double>(sums.g[k]);
  *cost = static_cast<double>(sums.cost);
}

template <typename Acc>
void ReprojRange(int64_t begin, int64_t end, const double* local_point, const double* pixel,
                 const double intrinsics[6], const double R[9], const double t[3], int loss_kind,
                 const double loss_params[2], double H21[21], double g[6], double* cost) {
  Sums<Acc, 21, 6> sums;
  double J[12], r[2];
  for (int64_t i = begin; i < end; ++i) {
    nlo_oracle_reproj_jacobian_residual(R, t, local_point + 3 * i, pixel + 2 * i, intrinsics, J, r);
    Accumulate<Acc, 2, 6, 21>(J, r, loss_kind, loss_params, &sums);
  }
  for (int k = 0; k < 21; ++k) H21[k] = static_cast<double>(sums.H[k]);
  for (int k = 0; k < 6; ++k) g[k] = static_cast<double>(sums.g[k]);
  *cost = static_cast<double>(sums.cost);
}

inline void PoseToState6(const double pose[16], double state6[9]) {
  /* Pose is column-major 4x4: element (r,c) at pose[4*c + r]. */
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = pose[4 * c + r];
  state6[0] = pose[12];
  state6[1] = pose[13];
  state6[2] = pose[14];
  nlo_oracle_rotmat_to_quat(R, state6 + 3);
  state6[7] = 0.001;                               /* lambda: ..._analytic.cc:89 */
  state6[8] = std::numeric_limits<double>::max();  /* previous_cost: :90 */
}

inline void State6ToPose(const double state6[9], double pose[16]) {
  double R[9];
  nlo_oracle_quat_to_rotmat(state6 + 3, R);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) pose[4 * c + r] = R[3 * r + c];
  pose[12] = state6[0];
  pose[13] = state6[1];
  pose[14] = state6[2];
}

inline void WriteTrace6(double* trace, int iteration, const double H21[21], const double g[6],
                        double cost, const double state6[9]) {
  if (trace == nullptr) return;
  double* row = trace + static_cast<int64_t>(iteration) * NLO_ORACLE_TRACE6;
  std::memcpy(row, H21, 21 * sizeof(double));
  std::memcpy(row + 21, g, 6 * sizeof(double));
  row[27] = cost;
  std::memcpy(row + 28, state6, 8 * sizeof(double)); /* t3 | q4 | lambda */
}

}  // namespace

extern "C" {

/* loss_function.h:28-33 (Exponential), :57-66 (Huber).  Cauchy is NOT in the reference; it is
 * the Ceres convention rho = c^2 log(1 + s/c^2), weight = rho' = 1 / (1 + s/c^2). */
void nlo_oracle_loss(int kind, const double params[2], double squared_residual, double out[3]) {
  switch (kind) {
    case NLO_ORACLE_LOSS_EXPONENTIAL: {
      const double c1 = params[0], c2 = params[1];
      const double two_c1c2 = 2.0 * c1 * c2;
      const double exp_term = std::exp(-c2 * squared_residual);
      out[0] = c1 - c1 * exp_term;
      out[1] = two_c1c2 * exp_term;
      out[2] = -2.0 * c2 * out[1];
      break;
    }
    case NLO_ORACLE_LOSS_HUBER: {
      const double threshold = params[0];
      const double squared_threshold = threshold * threshold;
      if (squared_residual > squared_threshold) {
        const double residual = std::sqrt(squared_residual);
        out[0] = 2.0 * threshold * residual - squared_threshold;
        out[1] = threshold / residual;
      } else {
        out[0] = squared_residual;
        out[1] = 1.0;
      }
      out[2] = 0.0;
      break;
    }
    case NLO_ORACLE_LOSS_CAUCHY: {
      const double c2 = params[0] * params[0];
      const double u = squared_residual / c2;
      out[0] = c2 * std::log1p(u);
      out[1] = 1.0 / (1.0 + u);
      out[2] = 0.0;
      break;
    }
    default:
      out[0] = squared_residual;
      out[1] = 1.0;
      out[2] = 0.0;
  }
}

/* Eigen::Quaterniond(Matrix3d) -- the trace / largest-diagonal branch selection of Eigen's
 * quaternionbase_assign_impl (used at ..._analytic.cc:87). */
void nlo_oracle_rotmat_to_quat(const double R[9], double q[4]) {
  auto m = [&](int r, int c) { return R[3 * r + c]; };
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
}

/* Eigen::Quaterniond::toRotationMatrix (used at ..._analytic.cc:99,154). */
void nlo_oracle_quat_to_rotmat(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

/* mahalanobis_distance_minimizer.cc:20-33 / reprojection_error_minimizer.h:35-52 */
void nlo_oracle_compute_quaternion(const double w[3], double q[4]) {
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (theta < 1e-6) {
    q[3] = 1.0;
    q[0] = 0.5 * w[0];
    q[1] = 0.5 * w[1];
    q[2] = 0.5 * w[2];
  } else {
    const double half_theta = theta * 0.5;
    const double sin_half_theta_divided_theta = std::sin(half_theta) / theta;
    q[3] = std::cos(half_theta);
    q[0] = sin_half_theta_divided_theta * w[0];
    q[1] = sin_half_theta_divided_theta * w[1];
    q[2] = sin_half_theta_divided_theta * w[2];
  }
}

/* mahalanobis_distance_minimizer_analytic.cc:159-185 */
void nlo_oracle_ndt6_jacobian_residual(const double R[9], const double t[3], const double p[3],
                                       const double mean[3], const double S[9], double J[18],
                                       double r[3]) {
  double p_warped[3], e[3];
  MatVec3(R, p, p_warped);
  for (int i = 0; i < 3; ++i) {
    p_warped[i] += t[i];
    e[i] = p_warped[i] - mean[i];
  }
  MatVec3(S, e, r);
  double skew_p[9], R_skew_p[9], SR[9];
  Skew(p, skew_p);
  MatMat3(R, skew_p, R_skew_p);
  MatMat3(S, R_skew_p, SR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = S[3 * i + j];
      J[6 * i + 3 + j] = -SR[3 * i + j];
    }
}

void nlo_oracle_ndt6_assemble(int64_t begin, int64_t end, const double* point, const double* mean,
                              const double* sqrt_info, const double R[9], const double t[3],
                              int loss_kind, const double loss_params[2], int long_double_accum,
                              double H21[21], double g[6], double* cost) {
  if (long_double_accum)
    Ndt6Range<long double>(begin, end, point, mean, sqrt_info, R, t, loss_kind, loss_params, H21,
                           g, cost);
  else
    Ndt6Range<double>(begin, end, point, mean, sqrt_info, R, t, loss_kind, loss_params, H21, g,
                      cost);
}

/* mahalanobis_distance_minimizer_analytic.cc:122-148 */
int nlo_oracle_gn6_step(const double H21[21], const double g[6], double cost,
                        double parameter_tolerance, double gradient_tolerance, double state6[9]) {
  constexpr double min_lambda = 1e-6; /* :81 */
  constexpr double max_lambda = 1e-2; /* :82 */
  double H[36];
  Unpack<6>(H21, H);
  double& lambda = state6[7];
  double& previous_cost = state6[8];
  for (int k = 0; k < 6; ++k) H[6 * k + k] *= 1.0 + lambda; /* :126 */
  double neg_g[6], step[6];
  for (int k = 0; k < 6; ++k) neg_g[k] = -g[k];
  SolveDense<6>(H, neg_g, step); /* :129 */
  for (int k = 0; k < 3; ++k) state6[k] += step[k]; /* :134 */
  double dq[4], qn[4];
  nlo_oracle_compute_quaternion(step + 3, dq);
  QuatMul(state6 + 3, dq, qn); /* :135 */
  QuatNormalize(qn);           /* :136 */
  std::memcpy(state6 + 3, qn, sizeof(qn));
  double step_norm = 0.0, grad_norm = 0.0;
  for (int k = 0; k < 6; ++k) {
    step_norm += step[k] * step[k];
    grad_norm += g[k] * g[k];
  }
  if (std::sqrt(step_norm) < parameter_tolerance) return 1; /* :139 */
  if (std::sqrt(grad_norm) < gradient_tolerance) return 1;  /* :142 */
  lambda *= (cost > previous_cost ? 2.0 : 0.6);             /* :146 */
  lambda = std::min(std::max(lambda, min_lambda), max_lambda);
  previous_cost = cost;
  return 0;
}

/* mahalanobis_distance_minimizer_analytic.cc:54-157 */
int nlo_oracle_ndt6_solve(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, int loss_kind, const double loss_params[2],
                          int max_iterations, double parameter_tolerance,
                          double gradient_tolerance, int num_threads, double pose[16],
                          int* iterations, double* final_cost, double* trace) {
  /* Executor split, :59-73: num_batch = int(max(1, N/T)); thread idx owns
   * [idx*num_batch, min((idx+1)*num_batch, N)); the N mod T tail is dropped. */
  std::vector<std::pair<int64_t, int64_t>> ranges;
  if (num_threads > 0) {
    const int64_t num_batch = static_cast<int64_t>(
        std::max(1.0, static_cast<double>(n) / static_cast<double>(num_threads)));
    for (int idx = 0; idx < num_threads; ++idx) {
      const int64_t b = std::min<int64_t>(static_cast<int64_t>(idx) * num_batch, n);
      const int64_t e = std::min<int64_t>(static_cast<int64_t>(idx + 1) * num_batch, n);
      ranges.emplace_back(b, e);
    }
  }
  double state6[9];
  PoseToState6(pose, state6);
  int iteration = 0;
  for (; iteration < max_iterations; ++iteration) {
    double R[9];
    nlo_oracle_quat_to_rotmat(state6 + 3, R);
    double H21[21], g[6], cost = 0.0;
    if (num_threads <= 0) {
      Ndt6Range<double>(0, n, point, mean, sqrt_info, R, state6, loss_kind, loss_params, H21, g,
                        &cost);
    } else {
      struct Part { double H[21]; double g[6]; double cost; };
      std::vector<Part> parts(ranges.size());
      std::vector<std::thread> workers;
      for (size_t k = 0; k < ranges.size(); ++k)
        workers.emplace_back([&, k]() {
          Ndt6Range<double>(ranges[k].first, ranges[k].second, point, mean, sqrt_info, R, state6,
                            loss_kind, loss_params, parts[k].H, parts[k].g, &parts[k].cost);
        });
      for (auto& w : workers) w.join();
      std::fill(H21, H21 + 21, 0.0);
      std::fill(g, g + 6, 0.0);
      for (const auto& part : parts) { /* :114-119, thread order */
        for (int k = 0; k < 6; ++k) g[k] += part.g[k];
        for (int k = 0; k < 21; ++k) H21[k] += part.H[k];
        cost += part.cost;
      }
    }
    const int converged =
        nlo_oracle_gn6_step(H21, g, cost, parameter_tolerance, gradient_tolerance, state6);
    WriteTrace6(trace, iteration, H21, g, cost, state6);
    if (converged) break;
  }
  if (iterations) *iterations = iteration;
  if (final_cost) *final_cost = state6[8]; /* "COST: previous_cost" :150 */
  State6ToPose(state6, pose);
  return 1; /* Solve always returns true, :156 */
}

/* mahalanobis_distance_minimizer_analytic_3dof.cc:110-139 */
void nlo_oracle_ndt3_jacobian_residual(const double R2[4], const double t2[2], const double p[3],
                                       const double mean[3], const double S[9], double J[9],
                                       double r[3]) {
  const double ux = p[0], uy = p[1];
  const double wx = R2[0] * ux + R2[1] * uy + t2[0];
  const double wy = R2[2] * ux + R2[3] * uy + t2[1];
  const double e[3] = {wx - mean[0], wy - mean[1], p[2] - mean[2]};
  MatVec3(S, e, r);
  const double k0 = -R2[0] * uy + R2[1] * ux; /* :133-135 */
  const double k1 = -R2[2] * uy + R2[3] * ux;
  for (int i = 0; i < 3; ++i) {
    J[3 * i + 0] = S[3 * i + 0];
    J[3 * i + 1] = S[3 * i + 1];
    J[3 * i + 2] = S[3 * i + 0] * k0 + S[3 * i + 1] * k1;
  }
}

void nlo_oracle_ndt3_assemble(int64_t begin, int64_t end, const double* point, const double* mean,
                              const double* sqrt_info, const double R2[4], const double t2[2],
                              int loss_kind, const double loss_params[2], int long_double_accum,
                              double H6[6], double g[3], double* cost) {
  if (long_double_accum)
    Ndt3Range<long double>(begin, end, point, mean, sqrt_info, R2, t2, loss_kind, loss_params, H6,
                           g, cost);
  else
    Ndt3Range<double>(begin, end, point, mean, sqrt_info, R2, t2, loss_kind, loss_params, H6, g,
                      cost);
}

/* mahalanobis_distance_minimizer_analytic_3dof.cc:69-99 */
int nlo_oracle_gn3_step(const double H6[6], const double g[3], double cost,
                        double parameter_tolerance, double gradient_tolerance, double state3[8]) {
  constexpr double min_lambda = 1e-6; /* :17 */
  constexpr double max_lambda = 1e-2; /* :18 */
  double H[9];
  Unpack<3>(H6, H);
  double& lambda = state3[6];
  double& previous_cost = state3[7];
  for (int k = 0; k < 3; ++k) H[3 * k + k] *= 1.0 + lambda; /* :74 */
  /* Eigen's 3x3 inverse() is the cofactor formula (:77). */
  const double c00 = H[4] * H[8] - H[5] * H[7];
  const double c01 = H[5] * H[6] - H[3] * H[8];
  const double c02 = H[3] * H[7] - H[4] * H[6];
  const double det = H[0] * c00 + H[1] * c01 + H[2] * c02;
  const double inv_det = 1.0 / det;
  double inv[9];
  inv[0] = c00 * inv_det;
  inv[1] = (H[2] * H[7] - H[1] * H[8]) * inv_det;
  inv[2] = (H[1] * H[5] - H[2] * H[4]) * inv_det;
  inv[3] = c01 * inv_det;
  inv[4] = (H[0] * H[8] - H[2] * H[6]) * inv_det;
  inv[5] = (H[2] * H[3] - H[0] * H[5]) * inv_det;
  inv[6] = c02 * inv_det;
  inv[7] = (H[1] * H[6] - H[0] * H[7]) * inv_det;
  inv[8] = (H[0] * H[4] - H[1] * H[3]) * inv_det;
  double step[3];
  for (int i = 0; i < 3; ++i)
    step[i] = inv[3 * i] * (-g[0]) + inv[3 * i + 1] * (-g[1]) + inv[3 * i + 2] * (-g[2]);
  state3[0] += step[0]; /* :82 */
  state3[1] += step[1];
  /* Isometry2d::rotate(angle): linear = linear * Rotation2D(angle)  (:83) */
  const double c = std::cos(step[2]), s = std::sin(step[2]);
  const double a = state3[2], b = state3[3], cc = state3[4], d = state3[5];
  state3[2] = a * c + b * s;
  state3[3] = -a * s + b * c;
  state3[4] = cc * c + d * s;
  state3[5] = -cc * s + d * c;
  const double step_norm = std::sqrt(step[0] * step[0] + step[1] * step[1] + step[2] * step[2]);
  const double grad_norm = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (step_norm < parameter_tolerance) return 1; /* :86 */
  if (grad_norm < gradient_tolerance) return 1;  /* :89 */
  lambda *= (cost > previous_cost ? 2.0 : 0.6);  /* :93-97 */
  lambda = std::min(std::max(lambda, min_lambda), max_lambda);
  previous_cost = cost;
  return 0;
}

/* mahalanobis_distance_minimizer_analytic_3dof.cc:14-108 */
int nlo_oracle_ndt3_solve(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, int loss_kind, const double loss_params[2],
                          int max_iterations, double parameter_tolerance,
                          double gradient_tolerance, double pose[16], int* iterations,
                          double* final_cost, double* trace) {
  double state3[8];
  state3[0] = pose[12]; /* :24 */
  state3[1] = pose[13];
  state3[2] = pose[0];  /* R2 row-major from the column-major 4x4: (0,0) */
  state3[3] = pose[4];  /* (0,1) */
  state3[4] = pose[1];  /* (1,0) */
  state3[5] = pose[5];  /* (1,1) */
  state3[6] = 0.001;
  state3[7] = std::numeric_limits<double>::max();
  const int64_t n_used = (n / 4) * 4; /* :33-36 */
  int iteration = 0;
  for (; iteration < max_iterations; ++iteration) {
    double H6[6], g[3], cost = 0.0;
    Ndt3Range<double>(0, n_used, point, mean, sqrt_info, state3 + 2, state3, loss_kind,
                      loss_params, H6, g, &cost);
    const int converged =
        nlo_oracle_gn3_step(H6, g, cost, parameter_tolerance, gradient_tolerance, state3);
    if (trace) {
      double* row = trace + static_cast<int64_t>(iteration) * NLO_ORACLE_TRACE3;
      std::memcpy(row, H6, 6 * sizeof(double));
      std::memcpy(row + 6, g, 3 * sizeof(double));
      row[9] = cost;
      std::memcpy(row + 10, state3, 7 * sizeof(double)); /* t2 | R2 | lambda */
    }
    if (converged) break;
  }
  if (iterations) *iterations = iteration;
  if (final_cost) *final_cost = state3[7];
  pose[12] = state3[0]; /* :104-105: only xy and the 2x2 block are written back */
  pose[13] = state3[1];
  pose[0] = state3[2];
  pose[4] = state3[3];
  pose[1] = state3[4];
  pose[5] = state3[5];
  return 1;
}

/* reprojection_error_minimizer_analytic.cc:107-162 */
void nlo_oracle_reproj_jacobian_residual(const double R[9], const double t[3], const double X[3],
                                         const double pixel[2], const double intrinsics[6],
                                         double J[12], double r[2]) {
  constexpr double kMinDepth = 0.03; /* :111 */
  double Xw[3];
  MatVec3(R, X, Xw);
  for (int i = 0; i < 3; ++i) Xw[i] += t[i];
  if (Xw[2] < kMinDepth) { /* :119-123 */
    for (int i = 0; i < 12; ++i) J[i] = 0.0;
    r[0] = r[1] = 0.0;
    return;
  }
  const double cx = intrinsics[2], cy = intrinsics[3];
  const double inv_fx = intrinsics[4], inv_fy = intrinsics[5];
  const double inverse_zw = 1.0 / Xw[2];
  r[0] = Xw[0] * inverse_zw - inv_fx * (pixel[0] - cx);
  r[1] = Xw[1] * inverse_zw - inv_fy * (pixel[1] - cy);
  const double squared_inverse_zw = inverse_zw * inverse_zw;
  double dK[6] = {inverse_zw, 0.0, -Xw[0] * squared_inverse_zw,
                  0.0, inverse_zw, -Xw[1] * squared_inverse_zw}; /* :141-146 */
  double skew_X[9], R_skew[9];
  Skew(X, skew_X);
  MatMat3(R, skew_X, R_skew);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = dK[3 * i + j];
      J[6 * i + 3 + j] = -(dK[3 * i] * R_skew[j] + dK[3 * i + 1] * R_skew[3 + j] +
                           dK[3 * i + 2] * R_skew[6 + j]);
    }
}

void nlo_oracle_reproj_assemble(int64_t begin, int64_t end, const double* local_point,
                                const double* pixel, const double intrinsics[6], const double R[9],
                                const double t[3], int loss_kind, const double loss_params[2],
                                int long_double_accum, double H21[21], double g[6], double* cost) {
  if (long_double_accum)
    ReprojRange<long double>(begin, end, local_point, pixel, intrinsics, R, t, loss_kind,
                             loss_params, H21, g, cost);
  else
    ReprojRange<double>(begin, end, local_point, pixel, intrinsics, R, t, loss_kind, loss_params,
                        H21, g, cost);
}

/* reprojection_error_minimizer_analytic.cc:12-105 (iteration body identical to the NDT 6-DoF
 * one: same lambda constants :15-16,:23 and schedule :93-99). */
int nlo_oracle_reproj_solve(int64_t n, const double* local_point, const double* pixel,
                            const double intrinsics[6], int loss_kind,
                            const double loss_params[2], int max_iterations,
                            double parameter_tolerance, double gradient_tolerance,
                            double pose[16], int* iterations, double* final_cost, double* trace) {
  double state6[9];
  PoseToState6(pose, state6);
  int iteration = 0;
  for (; iteration < max_iterations; ++iteration) {
    double R[9];
    nlo_oracle_quat_to_rotmat(state6 + 3, R);
    double H21[21], g[6], cost = 0.0;
    ReprojRange<double>(0, n, local_point, pixel, intrinsics, R, state6, loss_kind, loss_params,
                        H21, g, &cost);
    const int converged =
        nlo_oracle_gn6_step(H21, g, cost, parameter_tolerance, gradient_tolerance, state6);
    WriteTrace6(trace, iteration, H21, g, cost, state6);
    if (converged) break;
  }
  if (iterations) *iterations = iteration;
  if (final_cost) *final_cost = state6[8];
  State6ToPose(state6, pose);
  return 1;
}

/* reprojection_error_minimizer/tests/simple_optimization_test.cc:115-135.  The loop variables
 * are accumulated doubles (x += 0.1), restated exactly. */
int64_t nlo_oracle_pnp_reference_points(double* xyz, int64_t capacity) {
  const double z = 3.0, x_min = -1.5, x_max = 1.5, y_min = -1.0, y_max = 1.0, point_step = 0.1;
  int64_t count = 0;
  for (double x = x_min; x <= x_max; x += point_step)
    for (double y = y_min; y <= y_max; y += point_step) {
      if (xyz != nullptr && count < capacity) {
        xyz[3 * count] = x;
        xyz[3 * count + 1] = y;
        xyz[3 * count + 2] = z;
      }
      ++count;
    }
  return count;
}

/* mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:170-204 */
int64_t nlo_oracle_room_points(double* xyz, int64_t capacity) {
  const double width = 5.0, length = 7.0, height = 2.5, point_step = 0.01;
  int64_t count = 0;
  auto push = [&](double x, double y, double z) {
    if (xyz != nullptr && count < capacity) {
      xyz[3 * count] = x;
      xyz[3 * count + 1] = y;
      xyz[3 * count + 2] = z;
    }
    ++count;
  };
  double x, y, z;
  z = 0.0;
  for (x = -length / 2.0; x <= length / 2.0; x += point_step)
    for (y = -width / 2.0; y <= width / 2.0; y += point_step) push(x, y, z);
  y = -width / 2.0;
  for (x = -length / 2.0; x <= length / 2.0; x += point_step)
    for (z = 0.0; z <= height; z += point_step) {
      push(x, y, z);
      push(x, -y, z);
    }
  x = -length / 2.0;
  for (y = -width / 2.0; y <= width / 2.0; y += point_step)
    for (z = 0.0; z <= height; z += point_step) {
      push(-x, y, z);
      push(x, y, z);
    }
  return count;
}

/* mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:282-294.  The reference
 * computes the pairing in `int` before widening; restated with the same types. */
uint64_t nlo_oracle_voxel_key(const double point[3], double inverse_voxel_resolution) {
  int x_key = static_cast<int>(std::floor(point[0] * inverse_voxel_resolution));
  int y_key = static_cast<int>(std::floor(point[1] * inverse_voxel_resolution));
  int z_key = static_cast<int>(std::floor(point[2] * inverse_voxel_resolution));
  x_key = x_key >= 0 ? 2 * x_key : -2 * x_key - 1;
  y_key = y_key >= 0 ? 2 * y_key : -2 * y_key - 1;
  z_key = z_key >= 0 ? 2 * z_key : -2 * z_key - 1;
  const uint64_t xy_key = (x_key + y_key) * (x_key + y_key + 1) / 2 + y_key;
  const uint64_t xyz_key = (xy_key + z_key) * (xy_key + z_key + 1) / 2 + z_key;
  return xyz_key;
}

}  /* extern "C" */
