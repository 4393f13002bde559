/*
 * nlo_oracle_simd.cc -- CPU TIMING BASELINE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Restates the fastest CPU path the reference publishes: the float, 8-lane AVX2+FMA assembly
 * of MahalanobisDistanceMinimizerAnalyticSIMDVarious::SolveFloatIntrinsicAligned
 * (mahalanobis_distance_minimizer_analytic_simd_various.cc:1300-1429: SoA float planes, one
 * 8-wide lane group per step, the robust loss evaluated per lane through the scalar double
 * function :1387-1400, float lane accumulators reduced horizontally at the end :1430-1447),
 * wrapped in the thread split on 8-lane stride boundaries of
 * mahalanobis_distance_minimizer_analytic_simd.cc:55-76.
 * It exists so bench.py can time "the reference's SIMD multithreaded CPU path" next to the GPU.
 * Float arithmetic => it is NOT the parity oracle (parity is pinned on the double path).
 */
#include <immintrin.h>

#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

#include "nlo_oracle.h"

namespace {

struct LaneSums {
  float H[21];
  float g[6];
  float cost;
};

inline float HorizontalSum(__m256 v) {
  alignas(32) float buf[8];
  _mm256_store_ps(buf, v);
  return buf[0] + buf[1] + buf[2] + buf[3] + buf[4] + buf[5] + buf[6] + buf[7];
}

/* planes[k] for k = 0..14: x y z | mx my mz | s00 s01 s02 s10 ... s22, each n floats. */
void AssembleRange(const float* const planes[15], int64_t begin, int64_t end, const float R[9],
                   const float t[3], int loss_kind, const double loss_params[2], LaneSums* out) {
  __m256 Rv[9], tv[3];
  for (int k = 0; k < 9; ++k) Rv[k] = _mm256_set1_ps(R[k]);
  for (int k = 0; k < 3; ++k) tv[k] = _mm256_set1_ps(t[k]);
  __m256 accH[21], accg[6], acc_cost = _mm256_setzero_ps();
  for (auto& v : accH) v = _mm256_setzero_ps();
  for (auto& v : accg) v = _mm256_setzero_ps();

  for (int64_t i = begin; i + 8 <= end; i += 8) {
    __m256 p[3], mu[3], S[9];
    for (int k = 0; k < 3; ++k) p[k] = _mm256_loadu_ps(planes[k] + i);
    for (int k = 0; k < 3; ++k) mu[k] = _mm256_loadu_ps(planes[3 + k] + i);
    for (int k = 0; k < 9; ++k) S[k] = _mm256_loadu_ps(planes[6 + k] + i);

    __m256 e[3], r[3];
    for (int a = 0; a < 3; ++a) {
      const __m256 pw = _mm256_fmadd_ps(
          Rv[3 * a], p[0], _mm256_fmadd_ps(Rv[3 * a + 1], p[1], _mm256_fmadd_ps(Rv[3 * a + 2], p[2], tv[a])));
      e[a] = _mm256_sub_ps(pw, mu[a]);
    }
    for (int a = 0; a < 3; ++a)
      r[a] = _mm256_fmadd_ps(S[3 * a], e[0],
                             _mm256_fmadd_ps(S[3 * a + 1], e[1], _mm256_mul_ps(S[3 * a + 2], e[2])));

    /* A = -R [p]x : row a is p x R_a  (= -(R_a x p)) */
    __m256 A[9];
    for (int a = 0; a < 3; ++a) {
      const __m256 r0 = Rv[3 * a], r1 = Rv[3 * a + 1], r2 = Rv[3 * a + 2];
      A[3 * a + 0] = _mm256_fmsub_ps(r2, p[1], _mm256_mul_ps(r1, p[2]));
      A[3 * a + 1] = _mm256_fmsub_ps(r0, p[2], _mm256_mul_ps(r2, p[0]));
      A[3 * a + 2] = _mm256_fmsub_ps(r1, p[0], _mm256_mul_ps(r0, p[1]));
    }
    __m256 J[18]; /* 3 x 6 */
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c) {
        J[6 * a + c] = S[3 * a + c];
        J[6 * a + 3 + c] = _mm256_fmadd_ps(
            S[3 * a], A[c], _mm256_fmadd_ps(S[3 * a + 1], A[3 + c], _mm256_mul_ps(S[3 * a + 2], A[6 + c])));
      }

    const __m256 sq = _mm256_fmadd_ps(r[0], r[0], _mm256_fmadd_ps(r[1], r[1], _mm256_mul_ps(r[2], r[2])));
    __m256 loss = sq, weight = _mm256_set1_ps(1.0f);
    if (loss_kind != NLO_ORACLE_LOSS_NONE) {
      alignas(32) float sq_buf[8], loss_buf[8], weight_buf[8];
      _mm256_store_ps(sq_buf, sq);
      for (int k = 0; k < 8; ++k) {
        double out3[3] = {0.0, 0.0, 0.0};
        nlo_oracle_loss(loss_kind, loss_params, sq_buf[k], out3);
        loss_buf[k] = static_cast<float>(out3[0]);
        weight_buf[k] = static_cast<float>(out3[1]);
      }
      loss = _mm256_load_ps(loss_buf);
      weight = _mm256_load_ps(weight_buf);
    }
    for (int c = 0; c < 6; ++c) {
      const __m256 jr = _mm256_fmadd_ps(J[c], r[0], _mm256_fmadd_ps(J[6 + c], r[1], _mm256_mul_ps(J[12 + c], r[2])));
      accg[c] = _mm256_add_ps(accg[c], _mm256_mul_ps(weight, jr));
    }
    int idx = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) {
        const __m256 jj = _mm256_fmadd_ps(
            J[a], J[b], _mm256_fmadd_ps(J[6 + a], J[6 + b], _mm256_mul_ps(J[12 + a], J[12 + b])));
        accH[idx] = _mm256_add_ps(accH[idx], _mm256_mul_ps(weight, jj));
        ++idx;
      }
    acc_cost = _mm256_add_ps(acc_cost, loss);
  }
  for (int k = 0; k < 21; ++k) out->H[k] = HorizontalSum(accH[k]);
  for (int k = 0; k < 6; ++k) out->g[k] = HorizontalSum(accg[k]);
  out->cost = HorizontalSum(acc_cost);
}

}  // namespace

extern "C" {

/* AoS(double, oracle array convention) -> 15 float planes; the per-Solve conversion the reference
 * performs at ..._simd_various.cc:1252-1266.  planes must hold 15*n floats (plane k at k*n). */
void nlo_oracle_simd_pack(int64_t n, const double* point, const double* mean,
                          const double* sqrt_info, float* planes) {
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) planes[k * n + i] = static_cast<float>(point[3 * i + k]);
    for (int k = 0; k < 3; ++k) planes[(3 + k) * n + i] = static_cast<float>(mean[3 * i + k]);
    for (int k = 0; k < 9; ++k) planes[(6 + k) * n + i] = static_cast<float>(sqrt_info[9 * i + k]);
  }
}

/* One float/AVX2 assembly pass over floor(n/8)*8 correspondences with `num_threads` std::threads
 * (0 or 1 = calling thread only).  Partials are added in thread order into doubles. */
void nlo_oracle_simd_ndt6_assemble(int64_t n, const float* planes, const double R_rowmajor[9],
                                   const double t[3], int loss_kind, const double loss_params[2],
                                   int num_threads, double H21[21], double g[6], double* cost) {
  const float* plane_ptr[15];
  for (int k = 0; k < 15; ++k) plane_ptr[k] = planes + static_cast<int64_t>(k) * n;
  float Rf[9], tf[3];
  for (int k = 0; k < 9; ++k) Rf[k] = static_cast<float>(R_rowmajor[k]);
  for (int k = 0; k < 3; ++k) tf[k] = static_cast<float>(t[k]);
  const int T = std::max(1, num_threads);
  const int64_t num_stride = n / 8;
  const int64_t stride_per_thread = num_stride / T; /* ..._analytic_simd.cc:59-68 */
  std::vector<LaneSums> parts(T);
  std::vector<std::thread> workers;
  for (int k = 0; k < T; ++k) {
    const int64_t b = 8 * stride_per_thread * k;
    const int64_t e = (T == 1) ? 8 * num_stride : 8 * stride_per_thread * (k + 1);
    if (T == 1) {
      AssembleRange(plane_ptr, b, e, Rf, tf, loss_kind, loss_params, &parts[k]);
    } else {
      workers.emplace_back([&, k, b, e]() {
        AssembleRange(plane_ptr, b, e, Rf, tf, loss_kind, loss_params, &parts[k]);
      });
    }
  }
  for (auto& w : workers) w.join();
  for (int k = 0; k < 21; ++k) H21[k] = 0.0;
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  *cost = 0.0;
  for (const auto& part : parts) {
    for (int k = 0; k < 21; ++k) H21[k] += part.H[k];
    for (int k = 0; k < 6; ++k) g[k] += part.g[k];
    *cost += part.cost;
  }
}

/* Threaded scalar-double assembly (the executor split of ..._analytic.cc:59-73,104-119) as a
 * single pass, for timing the double path on all host cores. */
void nlo_oracle_ndt6_assemble_threads(int64_t n, const double* point, const double* mean,
                                      const double* sqrt_info, const double R_rowmajor[9],
                                      const double t[3], int loss_kind,
                                      const double loss_params[2], int num_threads,
                                      double H21[21], double g[6], double* cost) {
  const int T = std::max(1, num_threads);
  const int64_t num_batch =
      static_cast<int64_t>(std::max(1.0, static_cast<double>(n) / static_cast<double>(T)));
  struct Part { double H[21]; double g[6]; double cost; };
  std::vector<Part> parts(T);
  std::vector<std::thread> workers;
  for (int k = 0; k < T; ++k) {
    const int64_t b = std::min<int64_t>(static_cast<int64_t>(k) * num_batch, n);
    const int64_t e = std::min<int64_t>(static_cast<int64_t>(k + 1) * num_batch, n);
    workers.emplace_back([&, k, b, e]() {
      nlo_oracle_ndt6_assemble(b, e, point, mean, sqrt_info, R_rowmajor, t, loss_kind,
                               loss_params, 0, parts[k].H, parts[k].g, &parts[k].cost);
    });
  }
  for (auto& w : workers) w.join();
  for (int k = 0; k < 21; ++k) H21[k] = 0.0;
  for (int k = 0; k < 6; ++k) g[k] = 0.0;
  *cost = 0.0;
  for (const auto& part : parts) {
    for (int k = 0; k < 21; ++k) H21[k] += part.H[k];
    for (int k = 0; k < 6; ++k) g[k] += part.g[k];
    *cost += part.cost;
  }
}

}  /* extern "C" */
