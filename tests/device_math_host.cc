// device_math_host.cc -- the per-correspondence math, the rotation to the canonical frame and the
// damped step of nonlinear_optimizer_for_slam_b200/csrc/nlo_device.cuh compiled FOR THE HOST (test
// infrastructure; nothing of the product links it).  tests/test_device_math_host.py generates
// nlo_device_host.cuh from the product header -- the only edit: the two MUFU seed instructions of
// FastRcp / FastSqrt (inline PTX) become a float reciprocal / reciprocal square root, the Newton steps
// behind them stay -- builds this file with g++ and compares it with the oracle, so that the CPU suite
// checks the very source the kernels are built from (residuals, Jacobian terms, losses, the 28 / 10
// accumulators, CanonPlan encoding, LDL^T step, quaternion update, lambda schedule) without a GPU.
// The warp-cooperative CanonicalRotate is replayed phase by phase over 32 "lanes".
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_runtime.h>  // __device__ / __forceinline__ as plain attributes for g++

using std::isfinite;
static inline void __syncwarp() {}
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

#include "nlo_device_host.cuh"

namespace {

using namespace nlo;

// CanonicalRotate (nlo_device.cuh), its __syncwarp-separated phases executed lane after lane
void CanonicalRotateHost(double* total, const double* R) {
  double scratch[9] = {0};
  CanonPlan plans[32];
  double v1[32], v2[6];
  for (int lane = 0; lane < 32; ++lane) plans[lane] = MakeCanonPlan(lane);
  for (int lane = 0; lane < 32; ++lane) {
    v1[lane] = 0.0;
    if (plans[lane].dest1 >= 0) v1[lane] = static_cast<double>(plans[lane].sign1) * PlanDot(plans[lane].d1, total, R);
    if (plans[lane].dest1 >= 32) scratch[plans[lane].dest1 - 32] = v1[lane];
  }
  for (int lane = 0; lane < 6; ++lane) v2[lane] = PlanDot(plans[lane].d2, scratch, R);
  for (int lane = 0; lane < 32; ++lane)
    if (plans[lane].dest1 >= 0 && plans[lane].dest1 < 32) total[plans[lane].dest1] = v1[lane];
  for (int lane = 0; lane < 6; ++lane) total[15 + lane] = v2[lane];
}

template <int LOSS>
void AssembleLoss(int kind, int64_t n, const double* a, const double* b, const double* S, const double* R,
                  const double* t, double p0, double p1, const double* K, double* out) {
  double acc[32];
  for (double& x : acc) x = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    if (kind == kReproj) {
      const double v[5] = {a[3 * i], a[3 * i + 1], a[3 * i + 2], b[2 * i], b[2 * i + 1]};
      ReprojPoint<LOSS>(v, R, t, K, p0, p1, true, acc);
    } else {
      double v[12] = {a[3 * i], a[3 * i + 1], a[3 * i + 2], b[3 * i], b[3 * i + 1], b[3 * i + 2]};
      InformationFromSqrt(S + 9 * i, v + 6);  // what every ingest kernel stores
      if (kind == kNdt6) Ndt6Point<LOSS>(v, R, t, p0, p1, true, acc);
      else Ndt3Point<LOSS>(v, R, t, p0, p1, true, acc);
    }
  }
  if (kind != kNdt3) CanonicalRotateHost(acc, R);
  for (int k = 0; k < 28; ++k) out[k] = acc[k];
}

void Assemble(int kind, int loss, int64_t n, const double* a, const double* b, const double* S, const double* R,
              const double* t, double p0, double p1, const double* K, double* out) {
  if (loss == kLossCauchy) p1 = 1.0 / (p0 * p0);  // nlo_set_loss: the kernels multiply by 1 / c^2
  switch (loss) {
    case kLossExponential: return AssembleLoss<kLossExponential>(kind, n, a, b, S, R, t, p0, p1, K, out);
    case kLossHuber: return AssembleLoss<kLossHuber>(kind, n, a, b, S, R, t, p0, p1, K, out);
    case kLossCauchy: return AssembleLoss<kLossCauchy>(kind, n, a, b, S, R, t, p0, p1, K, out);
    default: return AssembleLoss<kLossNone>(kind, n, a, b, S, R, t, p0, p1, K, out);
  }
}

// init_states_kernel of nlo_kernels.cu
void InitState(State* s, const double* P, int kind) {
  memset(s, 0, sizeof(State));
  if (kind == kNdt3) {
    s->t[0] = P[12]; s->t[1] = P[13];
    s->R[0] = P[0]; s->R[1] = P[4]; s->R[2] = P[1]; s->R[3] = P[5];
  } else {
    double Rin[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rin[3 * r + c] = P[4 * c + r];
    RotToQuat(Rin, s->q);
    QuatToRot(s->q, s->R);
    s->t[0] = P[12]; s->t[1] = P[13]; s->t[2] = P[14];
  }
  s->lambda = 0.001;
  s->previous_cost = DBL_MAX;
}

}  // namespace

extern "C" {

// kind 0 ndt6 / 1 ndt3 / 2 reprojection.  a = point[3n]; b = mean[3n] (pixel[2n]); S = sqrt_info[9n] row-major
// (unused for kind 2); pose = column-major 4x4; K = fx fy cx cy 1/fx 1/fy.  out = canonical sums (28; 10 for kind 1).
void hm_assemble(int kind, int loss, int64_t n, const double* a, const double* b, const double* S, const double* pose,
                 double p0, double p1, const double* K, double* out) {
  State st;
  InitState(&st, pose, kind);
  Assemble(kind, loss, n, a, b, S, st.R, st.t, p0, p1, K, out);
}

// The whole loop of a Solve: assemble -> Step6 / Step3 until done.  pose in / out; returns the iteration count.
int hm_solve(int kind, int loss, int64_t n, const double* a, const double* b, const double* S, double* pose, double p0,
             double p1, const double* K, int max_iterations, double ptol, double gtol, double* final_cost) {
  State st;
  InitState(&st, pose, kind);
  double sums[28];
  while (!st.done && max_iterations > 0) {
    Assemble(kind, loss, n, a, b, S, st.R, st.t, p0, p1, K, sums);
    if (kind == kNdt3) Step3(sums, &st, ptol, gtol, max_iterations, nullptr);
    else Step6(sums, &st, ptol, gtol, max_iterations, nullptr);
  }
  // finish_states_kernel
  if (kind == kNdt3) {
    pose[12] = st.t[0]; pose[13] = st.t[1];
    pose[0] = st.R[0]; pose[4] = st.R[1]; pose[1] = st.R[2]; pose[5] = st.R[3];
  } else {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) pose[4 * c + r] = st.R[3 * r + c];
    pose[12] = st.t[0]; pose[13] = st.t[1]; pose[14] = st.t[2];
  }
  *final_cost = st.previous_cost;
  return st.status != 0 ? -1 : st.iteration;
}

}  // extern "C"
