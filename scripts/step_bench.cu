// Micro-benchmark of the single-warp damped step (profiling aid, not product code):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I nonlinear_optimizer_for_slam_b200/csrc \
//        scripts/step_bench.cu -o scripts/step_bench && ./scripts/step_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "nlo_device.cuh"

using namespace nlo;

__global__ void bench(const double* sums_in, long long* cycles, double* out, int reps) {
  __shared__ double sums[32];
  __shared__ State st;
  const int lane = threadIdx.x;
  if (lane < 28) sums[lane] = sums_in[lane];
  if (lane == 0) {
    st = State{};
    st.q[3] = 1.0; st.R[0] = st.R[4] = st.R[8] = 1.0; st.lambda = 1e-3; st.previous_cost = 1e300;
  }
  __syncwarp();
  double acc = 0.0;
  // 0: warp solve only
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (lane == 0) {
      double x[6];
      SolveGeneral6(sums, 1.001 + acc * 1e-300, x);
      acc += x[0];
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) cycles[0] = (t1 - t0) / reps;
  // 1: serial in-place LDLT (lane 0)
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (lane == 0) {
      double x[6];
      SolveSpd6(sums, 1.001 + acc * 1e-300, x);
      acc += x[0];
    }
    __syncwarp();
  }
  t1 = clock64();
  if (lane == 0) cycles[1] = (t1 - t0) / reps;
  // 2: the full Step6 (state update included)
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (lane == 0) { st.done = 0; st.iteration = 0; }
    __syncwarp();
    if (lane == 0) Step6(sums, &st, 0.0, 0.0, 1000000, nullptr);
    __syncwarp();
  }
  t1 = clock64();
  if (lane == 0) cycles[2] = (t1 - t0) / reps;
  // 4: the tail alone (pose update etc. for a fixed step)
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (lane == 0) {
      st.done = 0; st.iteration = 0;
      double x[6] = {1e-3 + acc * 1e-300, -2e-3, 5e-4, 1e-3, 2e-3, -1e-3};
      ApplyStep6(sums, x, &st, 0.0, 0.0, 1000000, nullptr);
      acc += st.R[1];
    }
    __syncwarp();
  }
  t1 = clock64();
  if (lane == 0) cycles[4] = (t1 - t0) / reps;
  // 3: canonical rotation, planned two-round form (in place on a scratch copy of the sums)
  __shared__ double canon_sums[32], canon_scratch[9];
  if (lane < 28) canon_sums[lane] = sums[lane];
  const CanonPlan plan = MakeCanonPlan(lane);
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    CanonicalRotate(canon_sums, st.R, canon_scratch, plan, lane);
    acc += canon_sums[lane & 15];
    __syncwarp();
  }
  t1 = clock64();
  if (lane == 0) cycles[3] = (t1 - t0) / reps;
  if (lane == 0) out[0] = acc + st.t[0];
}

int main() {
  // a well conditioned SPD system: H = M^T M + I
  double H[6][6], M[6][6];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) M[i][j] = ((i * 7 + j * 3) % 11) / 11.0 - 0.4;
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
    double s = (i == j) ? 1.0 : 0.0;
    for (int k = 0; k < 6; ++k) s += M[k][i] * M[k][j];
    H[i][j] = s * 1000.0;
  }
  double sums[28];
  int k = 0;
  for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) sums[k++] = H[i][j];
  for (int i = 0; i < 6; ++i) sums[21 + i] = 0.3 * (i + 1);
  sums[27] = 123.0;
  double* d_sums; long long* d_cycles; double* d_out;
  cudaMalloc(&d_sums, sizeof(sums)); cudaMalloc(&d_cycles, 8 * sizeof(long long)); cudaMalloc(&d_out, 8);
  cudaMemcpy(d_sums, sums, sizeof(sums), cudaMemcpyHostToDevice);
  bench<<<1, 32>>>(d_sums, d_cycles, d_out, 200);
  cudaDeviceSynchronize();
  bench<<<1, 32>>>(d_sums, d_cycles, d_out, 2000);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[5];
  cudaMemcpy(c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%s\ncycles per call: pivoted fallback %lld | serial LDLT %lld | full Step6 %lld | canonical %lld | tail %lld\n",
         cudaGetErrorString(e), c[0], c[1], c[2], c[3], c[4]);
  return 0;
}
