// dropin_bench.cc -- end-to-end timing of the reference-facing call itself:
//
//     MahalanobisDistanceMinimizerCuda::Solve(options, std::vector<Correspondence>, &pose)
//
// exactly as a caller of the reference's MahalanobisDistanceMinimizer::Solve
// (mahalanobis_distance_minimizer.h:31-33) makes it: the correspondences are the reference's own
// 304-byte AoS records in a pageable std::vector; the timed region covers ingest (host gather ->
// pinned ring -> PCIe -> device repack), the device-resident Gauss-Newton loop and the pose
// read-back.  Prints ONE JSON line on stdout.
//
//   dropin_bench --n 16777216 [--devices 0,1,...] [--planar] [--max-iterations 40] [--force-iterations]
//                [--reps 3] [--loss exponential|huber|none]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "nonlinear_optimizer/mahalanobis_distance_minimizer/mahalanobis_distance_minimizer_cuda.h"

using namespace nonlinear_optimizer;
using namespace nonlinear_optimizer::mahalanobis_distance_minimizer;

namespace {

uint64_t SplitMix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
double U01(uint64_t bits) { return static_cast<double>(bits >> 11) * (1.0 / 9007199254740992.0); }

// A synthetic scan of the reference's 7 x 5 x 2.5 m room (tests/simple_optimization_test.cc:170-204)
// seen from `true_pose`, every point associated with a planar NDT cell of a 0.5 m voxel grid on
// the surface it came from: mean = voxel centre on the surface, sqrt_information = diag(1/sigma) with
// sigma = 0.144 m in the plane (uniform over 0.5 m) and 1 % of that variance across it
// (UpdateNdtMap's clamp, :271-272).  Deterministic in (seed, i).
void FillRecord(uint64_t seed, uint64_t i, const double Rt[9], const double tt[3], Correspondence* c) {
  uint64_t h = SplitMix(seed ^ SplitMix(i));
  const double u0 = U01(h); h = SplitMix(h);
  const double u1 = U01(h); h = SplitMix(h);
  const double u2 = U01(h); h = SplitMix(h);
  const double n0 = U01(h); h = SplitMix(h);
  const double n1 = U01(h);
  double w[3];
  int normal;  // axis across the surface
  const double a = u0 * 95.0;
  if (a < 35.0) { w[0] = -3.5 + 7.0 * u1; w[1] = -2.5 + 5.0 * u2; w[2] = 0.0; normal = 2; }
  else if (a < 52.5) { w[0] = -3.5 + 7.0 * u1; w[1] = -2.5; w[2] = 2.5 * u2; normal = 1; }
  else if (a < 70.0) { w[0] = -3.5 + 7.0 * u1; w[1] = 2.5; w[2] = 2.5 * u2; normal = 1; }
  else if (a < 82.5) { w[0] = -3.5; w[1] = -2.5 + 5.0 * u1; w[2] = 2.5 * u2; normal = 0; }
  else { w[0] = 3.5; w[1] = -2.5 + 5.0 * u1; w[2] = 2.5 * u2; normal = 0; }
  double mean[3];
  for (int k = 0; k < 3; ++k) mean[k] = (k == normal) ? w[k] : (std::floor(w[k] * 2.0) + 0.5) * 0.5;
  // measurement noise across the surface (Box-Muller), 1 cm
  w[normal] += 0.01 * std::sqrt(-2.0 * std::log(std::max(n0, 1e-300))) * std::cos(6.283185307179586 * n1);
  // sensor frame: l = R_true^T (w - t_true)
  const double d[3] = {w[0] - tt[0], w[1] - tt[1], w[2] - tt[2]};
  for (int r = 0; r < 3; ++r) c->point(r) = Rt[r] * d[0] + Rt[3 + r] * d[1] + Rt[6 + r] * d[2];
  c->ndt.count = 100;
  c->ndt.is_valid = true;
  for (int k = 0; k < 3; ++k) c->ndt.mean(k) = mean[k];
  const double s_in = 1.0 / 0.1443, s_across = 1.0 / 0.01443;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col)
      c->ndt.sqrt_information(r, col) = (r == col) ? (r == normal ? s_across : s_in) : 0.0;
}

std::vector<int> ParseDevices(const char* s) {
  std::vector<int> out;
  while (*s) {
    out.push_back(std::atoi(s));
    const char* comma = std::strchr(s, ',');
    if (comma == nullptr) break;
    s = comma + 1;
  }
  return out;
}

double NowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
  int64_t n = 1 << 20;
  std::vector<int> devices = {0};
  bool planar = false, force = false;
  int max_iterations = 40, reps = 3;
  std::string loss = "exponential";
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "--n") n = std::atoll(next());
    else if (a == "--devices") devices = ParseDevices(next());
    else if (a == "--planar") planar = true;
    else if (a == "--force-iterations") force = true;
    else if (a == "--max-iterations") max_iterations = std::atoi(next());
    else if (a == "--reps") reps = std::atoi(next());
    else if (a == "--loss") loss = next();
    else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  // true pose of the reference's fixtures (simple_optimization_test.cc:85-88; 3dof_6dof_comparison_test.cc:77-80)
  const double yaw = planar ? 0.2 : 0.1;
  const double tt[3] = {planar ? -0.15 : -0.2, planar ? 0.05 : 0.123, planar ? 0.0 : 0.3};
  const double Rt[9] = {std::cos(yaw), -std::sin(yaw), 0.0, std::sin(yaw), std::cos(yaw), 0.0, 0.0, 0.0, 1.0};

  const double t_gen = NowMs();
  std::vector<Correspondence> correspondences(static_cast<size_t>(n));
  {
    const int T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
      pool.emplace_back([&, t]() {
        for (int64_t i = n * t / T; i < n * (t + 1) / T; ++i)
          FillRecord(1004, static_cast<uint64_t>(i), Rt, tt, &correspondences[static_cast<size_t>(i)]);
      });
    for (auto& th : pool) th.join();
  }
  const double gen_ms = NowMs() - t_gen;

  std::unique_ptr<MahalanobisDistanceMinimizer> optimizer;
  if (planar) optimizer = std::make_unique<MahalanobisDistanceMinimizerCuda3DOF>(devices);
  else optimizer = std::make_unique<MahalanobisDistanceMinimizerCuda>(devices);
  if (loss == "exponential") optimizer->SetLossFunction(std::make_shared<ExponentialLossFunction>(1.0, 1.0));
  else if (loss == "huber") optimizer->SetLossFunction(std::make_shared<HuberLossFunction>(1.0));
  Options options;
  options.max_iterations = max_iterations;
  if (force) {  // norm < 0 is never true: every Solve runs max_iterations iterations
    options.convergence_handle.parameter_tolerance = 0.0;
    options.convergence_handle.gradient_tolerance = 0.0;
  }
  auto last = [&]() -> const nlo_solve_result& {
    return planar ? static_cast<MahalanobisDistanceMinimizerCuda3DOF*>(optimizer.get())->last_result()
                  : static_cast<MahalanobisDistanceMinimizerCuda*>(optimizer.get())->last_result();
  };
  auto ingest = [&](double* total, double* gather) {
    if (planar) static_cast<MahalanobisDistanceMinimizerCuda3DOF*>(optimizer.get())->last_ingest_ms(total, gather);
    else static_cast<MahalanobisDistanceMinimizerCuda*>(optimizer.get())->last_ingest_ms(total, gather);
  };

  // warm-up Solve: context creation, module load, pinned ring, device problem
  Pose pose = Pose::Identity();
  const double t_first = NowMs();
  if (!optimizer->Solve(options, correspondences, &pose)) return 1;
  const double first_ms = NowMs() - t_first;

  std::vector<double> wall, ing, gat, dev;
  int iterations = 0;
  for (int r = 0; r < reps; ++r) {
    pose = Pose::Identity();
    const double t0 = NowMs();
    if (!optimizer->Solve(options, correspondences, &pose)) return 1;
    wall.push_back(NowMs() - t0);
    double ti = 0.0, tg = 0.0;
    ingest(&ti, &tg);
    ing.push_back(ti);
    gat.push_back(tg);
    dev.push_back(last().device_ms);
    iterations = last().iterations + (force ? 0 : 1);  // a converged Solve assembled `iterations + 1` times
  }
  auto median = [](std::vector<double> v) {
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
  };
  const double w = median(wall), in = median(ing), g = median(gat), d = median(dev);
  const double h2d = 120.0 * static_cast<double>(n);  // 15 doubles per record cross PCIe
  const double* P = PoseData(pose);
  std::printf(
      "{\"bench\": \"dropin_solve\", \"kind\": \"%s\", \"n\": %lld, \"devices\": %zu, \"loss\": \"%s\", "
      "\"record_bytes\": %zu, \"host_memory\": \"pageable std::vector<Correspondence>\", "
      "\"max_iterations\": %d, \"forced\": %s, \"assemblies_per_solve\": %d, "
      "\"solve_wall_ms\": %.4f, \"first_solve_wall_ms\": %.3f, \"ingest_ms\": %.4f, \"host_gather_ms\": %.4f, "
      "\"device_loop_ms\": %.4f, \"h2d_bytes_per_solve\": %.0f, \"pcie_gbs\": %.3f, "
      "\"aos_read_gbs\": %.3f, \"gpoints_s\": %.4f, \"solves_per_s\": %.3f, \"generate_ms\": %.1f, "
      "\"pose_t\": [%.9f, %.9f, %.9f], \"final_cost\": %.9g}\n",
      planar ? "ndt3" : "ndt6", static_cast<long long>(n), devices.size(), loss.c_str(), sizeof(Correspondence),
      max_iterations, force ? "true" : "false", iterations, w, first_ms, in, g, d, h2d,
      h2d / (in * 1e-3) / 1e9, static_cast<double>(sizeof(Correspondence)) * n / (in * 1e-3) / 1e9,
      static_cast<double>(n) * iterations / (w * 1e-3) / 1e9, 1e3 / w, gen_ms, P[12], P[13], P[14],
      last().final_cost);
  return 0;
}
