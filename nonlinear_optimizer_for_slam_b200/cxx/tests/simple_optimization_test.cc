// simple_optimization_test.cc -- the reference's solver "tests" as real assertions, through the
// drop-in C++ classes.  Mirrors
//   reprojection_error_minimizer/tests/simple_optimization_test.cc  (630-point PnP fixture)
//   mahalanobis_distance_minimizer/tests/simple_optimization_test.cc (room -> NDT map -> register)
//   mahalanobis_distance_minimizer/tests/3dof_6dof_comparison_test.cc (planar variant)
// of /root/reference/nonlinear_optimizer.  The reference prints poses for a human to compare with
// "True pose"; here each case checks the numbers.  Needs a B200; exits non-zero on any failure.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <memory>
#include <unordered_map>
#include <vector>

#include "nonlinear_optimizer/mahalanobis_distance_minimizer/mahalanobis_distance_minimizer_cuda.h"
#include "nonlinear_optimizer/mahalanobis_distance_minimizer/ndt_registration_cuda.h"
#include "nonlinear_optimizer/reprojection_error_minimizer/reprojection_error_minimizer_cuda.h"

using namespace nonlinear_optimizer;

namespace {

int g_failures = 0;
#define CHECK_NEAR(a, b, tol)                                                              \
  do {                                                                                     \
    const double _a = (a), _b = (b);                                                       \
    if (!(std::fabs(_a - _b) <= (tol))) {                                                  \
      std::fprintf(stderr, "FAIL %s:%d  %s = %.12g, expected %.12g +- %g\n", __FILE__,     \
                   __LINE__, #a, _a, _b, static_cast<double>(tol));                        \
      ++g_failures;                                                                        \
    }                                                                                      \
  } while (0)
#define CHECK_TRUE(c)                                                                      \
  do {                                                                                     \
    if (!(c)) {                                                                            \
      std::fprintf(stderr, "FAIL %s:%d  %s\n", __FILE__, __LINE__, #c);                    \
      ++g_failures;                                                                        \
    }                                                                                      \
  } while (0)

Pose YawPose(double x, double y, double z, double yaw) {
  Pose p = Pose::Identity();
  const double c = std::cos(yaw), s = std::sin(yaw);
  double* m = PoseData(p);  // column-major 4x4
  m[0] = c; m[1] = s; m[4] = -s; m[5] = c;
  m[12] = x; m[13] = y; m[14] = z;
  return p;
}

void Apply(const Pose& T, const double in[3], double out[3]) {
  const double* m = PoseData(T);
  for (int r = 0; r < 3; ++r) out[r] = m[r] * in[0] + m[4 + r] * in[1] + m[8 + r] * in[2] + m[12 + r];
}

Pose Inverse(const Pose& T) {
  Pose out = Pose::Identity();
  const double* m = PoseData(T);
  double* o = PoseData(out);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) o[4 * c + r] = m[4 * r + c];
  for (int r = 0; r < 3; ++r) o[12 + r] = -(o[r] * m[12] + o[4 + r] * m[13] + o[8 + r] * m[14]);
  return out;
}

double YawOf(const Pose& T) { return std::atan2(PoseData(T)[1], PoseData(T)[0]); }

// ---------------------------------------------------------------- PnP (reprojection error)
void TestReprojection() {
  using namespace reprojection_error_minimizer;
  CameraIntrinsics K;
  K.fx = K.fy = 525.0; K.cx = 320.0; K.cy = 240.0;
  K.inv_fx = 1.0 / K.fx; K.inv_fy = 1.0 / K.fy; K.width = 640; K.height = 480;
  const Pose true_pose = YawPose(-0.1, 0.123, -0.5, 0.1);
  const Pose true_inv = Inverse(true_pose);
  std::vector<Correspondence> correspondences;
  for (double x = -1.5; x <= 1.5; x += 0.1)
    for (double y = -1.0; y <= 1.0; y += 0.1) {
      const double X[3] = {x, y, 3.0};
      double q[3];
      Apply(true_inv, X, q);
      Correspondence corr;
      corr.local_point(0) = X[0]; corr.local_point(1) = X[1]; corr.local_point(2) = X[2];
      corr.matched_pixel(0) = K.fx * q[0] / q[2] + K.cx;
      corr.matched_pixel(1) = K.fy * q[1] / q[2] + K.cy;
      correspondences.push_back(corr);
    }
  std::cerr << "# points: " << correspondences.size() << std::endl;
  CHECK_TRUE(correspondences.size() == 630);

  std::unique_ptr<ReprojectionErrorMinimizer> optimizer = std::make_unique<ReprojectionErrorMinimizerCuda>();
  optimizer->SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
  Options options;
  Pose pose = Pose::Identity();
  CHECK_TRUE(optimizer->Solve(options, correspondences, K, &pose));
  const auto& res = static_cast<ReprojectionErrorMinimizerCuda*>(optimizer.get())->last_result();
  CHECK_TRUE(res.iterations == 6);                       // results/reproj_amd64.txt:5
  CHECK_NEAR(res.final_cost, 2.33228e-11, 5e-17);
  const Pose inv = Inverse(pose);                        // the reference prints Solve(...).inverse()
  CHECK_NEAR(PoseData(inv)[12], -0.1, 1e-6);             // results/reproj_amd64.txt:10
  CHECK_NEAR(PoseData(inv)[13], 0.123, 1e-6);
  CHECK_NEAR(PoseData(inv)[14], -0.5, 1e-6);
  CHECK_NEAR(YawOf(inv), 0.1, 1e-6);

  // a loss the device cannot express must be refused, not silently replaced
  struct Custom : LossFunction {
    void Evaluate(const double s, double* out) override { out[0] = s; out[1] = 1.0; }
  };
  optimizer->SetLossFunction(std::make_shared<Custom>());
  Pose pose2 = Pose::Identity();
  CHECK_TRUE(!optimizer->Solve(options, correspondences, K, &pose2));
  bool threw = false;
  try { HuberLossFunction bad(0.0); } catch (const std::out_of_range&) { threw = true; }
  CHECK_TRUE(threw);                                     // loss_function.h:53-54
}

// ---------------------------------------------------------------- NDT map fixture
struct Cell {
  int count = 0;
  double sum[3] = {0, 0, 0};
  double moment[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // starts at Identity (types.h:14)
  mahalanobis_distance_minimizer::NDT ndt;
};

// cyclic Jacobi for a symmetric 3x3: A = V diag(w) V^T, eigenvalues sorted ascending
void SymEig3(const double A_in[9], double w[3], double V[9]) {
  double A[9];
  for (int i = 0; i < 9; ++i) { A[i] = A_in[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
    if (off < 1e-30) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (std::fabs(A[3 * p + q]) < 1e-300) continue;
        const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * A[3 * p + q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - s * akq;
          A[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - s * aqk;
          A[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (w[j] > w[j + 1]) {
        std::swap(w[j], w[j + 1]);
        for (int k = 0; k < 3; ++k) std::swap(V[3 * k + j], V[3 * k + j + 1]);
      }
}

uint64_t VoxelKey(const double p[3], double inv_res) {  // tests/simple_optimization_test.cc:282-294
  int k[3];
  for (int a = 0; a < 3; ++a) {
    k[a] = static_cast<int>(std::floor(p[a] * inv_res));
    k[a] = k[a] >= 0 ? 2 * k[a] : -2 * k[a] - 1;
  }
  const uint64_t xy = static_cast<uint64_t>(k[0] + k[1]) * (k[0] + k[1] + 1) / 2 + k[1];
  return (xy + k[2]) * (xy + k[2] + 1) / 2 + k[2];
}

using NdtMap = std::unordered_map<uint64_t, Cell>;

std::vector<Vec3> RoomPoints(double step) {  // :170-204, coarser step to keep the test quick
  std::vector<Vec3> pts;
  auto push = [&](double x, double y, double z) {
    Vec3 v; v(0) = x; v(1) = y; v(2) = z; pts.push_back(v);
  };
  for (double x = -3.5; x <= 3.5; x += step)
    for (double y = -2.5; y <= 2.5; y += step) push(x, y, 0.0);
  for (double x = -3.5; x <= 3.5; x += step)
    for (double z = 0.0; z <= 2.5; z += step) { push(x, -2.5, z); push(x, 2.5, z); }
  for (double y = -2.5; y <= 2.5; y += step)
    for (double z = 0.0; z <= 2.5; z += step) { push(3.5, y, z); push(-3.5, y, z); }
  return pts;
}

void BuildNdtMap(const std::vector<Vec3>& points, double voxel, NdtMap* map) {  // :236-280
  const double inv = 1.0 / voxel;
  for (const Vec3& p : points) {
    const double q[3] = {p(0), p(1), p(2)};
    Cell& c = (*map)[VoxelKey(q, inv)];
    ++c.count;
    for (int a = 0; a < 3; ++a) {
      c.sum[a] += q[a];
      for (int b = 0; b < 3; ++b) c.moment[3 * a + b] += q[a] * q[b];
    }
  }
  for (auto& kv : *map) {
    Cell& c = kv.second;
    if (c.count < 5) continue;
    double mean[3], cov[9];
    for (int a = 0; a < 3; ++a) mean[a] = c.sum[a] / c.count;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) cov[3 * a + b] = c.moment[3 * a + b] / c.count - mean[a] * mean[b];
    double w[3], V[9];
    SymEig3(cov, w, V);
    if (w[2] < 0.01) continue;
    w[0] = std::max(w[0], 0.01 * w[2]);
    w[1] = std::max(w[1], 0.01 * w[2]);
    auto& ndt = c.ndt;
    ndt.count = c.count;
    ndt.is_valid = true;
    for (int a = 0; a < 3; ++a) ndt.mean(a) = mean[a];
    // sqrt_information = diag(w^-1/2) * V  (V, not V^T, as the reference writes it, :275-276)
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col) ndt.sqrt_information(r, col) = V[3 * r + col] / std::sqrt(w[r]);
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += ndt.sqrt_information(k, r) * ndt.sqrt_information(k, col);
        ndt.information(r, col) = s;
      }
  }
}

// Association as in the reference's MatchPointCloud (:296-342): for every warped point the (at
// most) two nearest valid cell means within 1.0 m, one correspondence per hit.  The reference
// asks a flann KD-tree; 96 cells are few enough for an exhaustive search here.
std::vector<mahalanobis_distance_minimizer::Correspondence> Match(const NdtMap& map, double /*voxel*/,
                                                                  const std::vector<Vec3>& local,
                                                                  const Pose& pose) {
  std::vector<const mahalanobis_distance_minimizer::NDT*> cells;
  for (const auto& kv : map)
    if (kv.second.ndt.is_valid) cells.push_back(&kv.second.ndt);
  std::vector<mahalanobis_distance_minimizer::Correspondence> out;
  out.reserve(2 * local.size());
  for (const Vec3& p : local) {
    const double q[3] = {p(0), p(1), p(2)};
    double w[3];
    Apply(pose, q, w);
    const mahalanobis_distance_minimizer::NDT* best[2] = {nullptr, nullptr};
    double best_d2[2] = {1.0, 1.0};  // squared radius 1.0
    for (const auto* ndt : cells) {
      const double dx = w[0] - ndt->mean(0), dy = w[1] - ndt->mean(1), dz = w[2] - ndt->mean(2);
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best_d2[0]) {
        best_d2[1] = best_d2[0]; best[1] = best[0];
        best_d2[0] = d2; best[0] = ndt;
      } else if (d2 < best_d2[1]) {
        best_d2[1] = d2; best[1] = ndt;
      }
    }
    for (int k = 0; k < 2; ++k) {
      if (best[k] == nullptr) continue;
      mahalanobis_distance_minimizer::Correspondence c;
      c.point = p;
      c.ndt = *best[k];
      out.push_back(c);
    }
  }
  return out;
}

void TestMahalanobis(bool planar) {
  using namespace mahalanobis_distance_minimizer;
  constexpr double kVoxel = 1.0;
  const std::vector<Vec3> global = RoomPoints(0.02);
  NdtMap map;
  BuildNdtMap(global, kVoxel, &map);
  int valid = 0;
  for (auto& kv : map) valid += kv.second.ndt.is_valid ? 1 : 0;
  std::cerr << "# points: " << global.size() << ", Ndt map size: " << map.size() << " (" << valid
            << " valid)" << std::endl;
  CHECK_TRUE(map.size() == 96);  // results/maha_amd64_simple.txt:2

  // 3dof_6dof_comparison_test.cc:77-80 (planar) / simple_optimization_test.cc:85-88
  const Pose true_pose = planar ? YawPose(-0.15, 0.05, 0.0, 0.2) : YawPose(-0.2, 0.123, 0.3, 0.1);
  const Pose true_inv = Inverse(true_pose);
  std::vector<Vec3> local;
  for (size_t i = 0; i < global.size(); i += 7) {  // thin the scan
    const double q[3] = {global[i](0), global[i](1), global[i](2)};
    double l[3];
    Apply(true_inv, q, l);
    Vec3 v; v(0) = l[0]; v(1) = l[1]; v(2) = l[2];
    local.push_back(v);
  }

  std::unique_ptr<MahalanobisDistanceMinimizer> optimizer;
  if (planar) optimizer = std::make_unique<MahalanobisDistanceMinimizerCuda3DOF>();
  else optimizer = std::make_unique<MahalanobisDistanceMinimizerCuda>();
  optimizer->SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
  optimizer->SetMultiThreadExecutor(std::make_shared<MultiThreadExecutor>(2));  // accepted, unused
  Options options;
  Pose pose = Pose::Identity();
  int outer = 0;
  for (; outer < 10; ++outer) {  // simple_optimization_test.cc:481-501
    const auto correspondences = Match(map, kVoxel, local, pose);
    const Pose last = pose;
    CHECK_TRUE(optimizer->Solve(options, correspondences, &pose));
    double dt = 0.0;
    for (int k = 12; k < 15; ++k) dt += std::pow(PoseData(pose)[k] - PoseData(last)[k], 2);
    if (std::sqrt(dt) < 1e-5 && std::fabs(YawOf(pose) - YawOf(last)) < 1e-5) break;
  }
  std::cerr << "outer iterations: " << outer << "  pose: " << PoseData(pose)[12] << " "
            << PoseData(pose)[13] << " " << PoseData(pose)[14] << " yaw " << YawOf(pose) << std::endl;
  const double* t = PoseData(true_pose);
  CHECK_NEAR(PoseData(pose)[12], t[12], 1e-2);  // the reference's own runs stop ~4 mm from truth
  CHECK_NEAR(PoseData(pose)[13], t[13], 1e-2);
  if (!planar) CHECK_NEAR(PoseData(pose)[14], t[14], 1e-2);
  else CHECK_NEAR(PoseData(pose)[14], 0.0, 0.0);  // z untouched, ..._analytic_3dof.cc:104-105
  CHECK_NEAR(YawOf(pose), YawOf(true_pose), 5e-3);
}

// The same registration with map building, matching and the outer loop on the device.
void TestDeviceRegistration(bool planar) {
  using namespace mahalanobis_distance_minimizer;
  const std::vector<Vec3> global = RoomPoints(0.02);
  const Pose true_pose = planar ? YawPose(-0.15, 0.05, 0.0, 0.2) : YawPose(-0.2, 0.123, 0.3, 0.1);
  const Pose true_inv = Inverse(true_pose);
  std::vector<Vec3> local;
  for (size_t i = 0; i < global.size(); i += 7) {
    const double q[3] = {global[i](0), global[i](1), global[i](2)};
    double l[3];
    Apply(true_inv, q, l);
    Vec3 v; v(0) = l[0]; v(1) = l[1]; v(2) = l[2];
    local.push_back(v);
  }
  NdtRegistrationCuda reg;
  reg.SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
  CHECK_TRUE(reg.BuildMap(global, 1.0));
  CHECK_TRUE(reg.SetScan(local));
  Options options;
  Pose pose = Pose::Identity();
  CHECK_TRUE(reg.Register(options, &pose, planar));
  std::cerr << "device registration: outer " << reg.last_result().outer_iterations << ", pose "
            << PoseData(pose)[12] << " " << PoseData(pose)[13] << " " << PoseData(pose)[14] << " yaw "
            << YawOf(pose) << ", " << reg.last_result().device_ms << " ms on the device" << std::endl;
  const double* t = PoseData(true_pose);
  CHECK_NEAR(PoseData(pose)[12], t[12], 2e-3);  // proper S = diag V^T: converges onto the truth
  CHECK_NEAR(PoseData(pose)[13], t[13], 2e-3);
  if (!planar) CHECK_NEAR(PoseData(pose)[14], t[14], 2e-3);
  CHECK_NEAR(YawOf(pose), YawOf(true_pose), 1e-3);
  CHECK_TRUE(reg.last_result().outer_iterations >= 2 && reg.last_result().outer_iterations <= 10);

  // the same registration against the voxel-hash form of the map: same correspondences, so the
  // same rounds and the same pose up to the rounding of the map's atomic sums
  const int dense_outer = reg.last_result().outer_iterations;
  CHECK_TRUE(reg.BuildMap(global, 1.0, /*reference_literal_sqrt_information=*/false, /*voxel_hash=*/true));
  Pose hashed_pose = Pose::Identity();
  CHECK_TRUE(reg.Register(options, &hashed_pose, planar));
  CHECK_TRUE(reg.last_result().outer_iterations == dense_outer);
  for (int k = 0; k < 16; ++k) CHECK_NEAR(PoseData(hashed_pose)[k], PoseData(pose)[k], 1e-8);
}

// Solve() over several B200s of this process (a device list) against the same Solve() on one:
// the split by point range sits inside Solve, where the reference splits over its thread pool
// (..._analytic.cc:59-73,104-119).  On a one-GPU box the list names device 0 twice: two shards, two
// iteration kernels and the same peer-memory all-reduce between them, on one device.
void TestShardedSolve(bool planar) {
  using namespace mahalanobis_distance_minimizer;
  const int visible = nlo_visible_device_count();
  CHECK_TRUE(visible >= 1);
  std::vector<int> devices;
  const int D = visible >= 2 ? std::min(visible, 8) : 2;
  for (int r = 0; r < D; ++r) devices.push_back(visible >= 2 ? r : 0);
  const size_t n = visible >= 2 ? 4000003 : 600001;  // not a multiple of 4 * D: the planar tail rule is global
  const Pose true_pose = planar ? YawPose(-0.15, 0.05, 0.0, 0.2) : YawPose(-0.2, 0.123, 0.3, 0.1);
  const Pose true_inv = Inverse(true_pose);
  std::vector<Correspondence> correspondences(n);
  uint64_t state = 88172645463325252ull;
  auto uniform = [&]() {
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    return static_cast<double>(state >> 11) * (1.0 / 9007199254740992.0);
  };
  for (size_t i = 0; i < n; ++i) {
    // a point on the floor or on the y = -2.5 wall, its 0.5 m planar cell, 1 cm noise across the surface
    const bool floor_hit = uniform() < 0.6;
    double w[3] = {-3.5 + 7.0 * uniform(), floor_hit ? -2.5 + 5.0 * uniform() : -2.5,
                   floor_hit ? 0.0 : 2.5 * uniform()};
    const int normal = floor_hit ? 2 : 1;
    Correspondence& c = correspondences[i];
    for (int k = 0; k < 3; ++k) c.ndt.mean(k) = (k == normal) ? w[k] : (std::floor(w[k] * 2.0) + 0.5) * 0.5;
    w[normal] += 0.02 * (uniform() - 0.5);
    double l[3];
    Apply(true_inv, w, l);
    for (int k = 0; k < 3; ++k) c.point(k) = l[k];
    c.ndt.is_valid = true;
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col)
        c.ndt.sqrt_information(r, col) = (r == col) ? (r == normal ? 69.3 : 6.93) : 0.0;
    // an off-diagonal term so that the column-major ingest matters
    c.ndt.sqrt_information(0, 1) = 0.37;
  }
  Options options;
  auto solve = [&](std::unique_ptr<MahalanobisDistanceMinimizer> optimizer, Pose* pose, nlo_solve_result* res) {
    optimizer->SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
    *pose = Pose::Identity();
    CHECK_TRUE(optimizer->Solve(options, correspondences, pose));
    *res = planar ? static_cast<MahalanobisDistanceMinimizerCuda3DOF*>(optimizer.get())->last_result()
                  : static_cast<MahalanobisDistanceMinimizerCuda*>(optimizer.get())->last_result();
  };
  Pose one, many;
  nlo_solve_result res_one{}, res_many{};
  if (planar) {
    solve(std::make_unique<MahalanobisDistanceMinimizerCuda3DOF>(0), &one, &res_one);
    solve(std::make_unique<MahalanobisDistanceMinimizerCuda3DOF>(devices), &many, &res_many);
  } else {
    solve(std::make_unique<MahalanobisDistanceMinimizerCuda>(0), &one, &res_one);
    solve(std::make_unique<MahalanobisDistanceMinimizerCuda>(devices), &many, &res_many);
  }
  std::cerr << "sharded Solve over " << D << " device(s) (" << visible << " visible), " << n
            << " correspondences: iterations " << res_one.iterations << " / " << res_many.iterations
            << ", cost " << res_one.final_cost << " / " << res_many.final_cost << std::endl;
  CHECK_TRUE(res_one.iterations == res_many.iterations);
  CHECK_TRUE(res_one.iterations >= 2);
  CHECK_NEAR(res_many.final_cost, res_one.final_cost, 1e-9 * std::fabs(res_one.final_cost));
  // the shards change only the order of the fp64 sums (every device of the sharded solve itself
  // ends bit-identical: nlo_ndt*_solve fails with NLO_ECOMM otherwise)
  for (int k = 0; k < 16; ++k) CHECK_NEAR(PoseData(many)[k], PoseData(one)[k], 1e-9);
  CHECK_NEAR(PoseData(one)[12], PoseData(true_pose)[12], 5e-3);
  CHECK_NEAR(PoseData(one)[13], PoseData(true_pose)[13], 5e-3);
}

}  // namespace

int main(int, char**) {
  std::cerr << "Start ReprojectionErrorMinimizerCuda" << std::endl;
  TestReprojection();
  std::cerr << "Start MahalanobisDistanceMinimizerCuda" << std::endl;
  TestMahalanobis(false);
  std::cerr << "Start MahalanobisDistanceMinimizerCuda3DOF" << std::endl;
  TestMahalanobis(true);
  std::cerr << "Start NdtRegistrationCuda" << std::endl;
  TestDeviceRegistration(false);
  TestDeviceRegistration(true);
  std::cerr << "Start sharded Solve (device list)" << std::endl;
  TestShardedSolve(false);
  TestShardedSolve(true);
  if (g_failures == 0) std::cerr << "ALL CXX TESTS PASSED" << std::endl;
  return g_failures == 0 ? 0 : 1;
}
