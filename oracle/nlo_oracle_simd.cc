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
 *
 * Also here, for the secondary configs: the float SIMD twins of the planar minimizer
 * (mahalanobis_distance_minimizer_analytic_3dof_simd.cc:85-176) and of the reprojection minimizer
 * (reprojection_error_minimizer/reprojection_error_minimizer_analytic_simd.cc:55-156).  The reference
 * runs both on ONE thread (neither touches the executor); num_threads > 1 applies the 8-lane thread
 * split of ..._analytic_simd.cc:55-76 to them as a favour to the CPU side.  Their quirks are kept:
 * the reprojection twin hands ||r|| -- `r__.norm()`, not the squared norm -- to the loss (:70) and
 * gates on z > 0 instead of z >= 0.03 (:64).
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

/* planar twin: ..._analytic_3dof_simd.cc:85-157.  H: 00 01 02 11 12 22. */
struct LaneSums3 {
  float H[6];
  float g[3];
  float cost;
};

void AssembleRange3(const float* const planes[15], int64_t begin, int64_t end, const float R[4],
                    const float t[2], int loss_kind, const double loss_params[2], LaneSums3* out) {
  const __m256 R00 = _mm256_set1_ps(R[0]), R01 = _mm256_set1_ps(R[1]), R10 = _mm256_set1_ps(R[2]),
               R11 = _mm256_set1_ps(R[3]), t0 = _mm256_set1_ps(t[0]), t1 = _mm256_set1_ps(t[1]);
  __m256 accH[6], accg[3], acc_cost = _mm256_setzero_ps();
  for (auto& v : accH) v = _mm256_setzero_ps();
  for (auto& v : accg) v = _mm256_setzero_ps();
  for (int64_t i = begin; i + 8 <= end; i += 8) {
    __m256 p[3], mu[3], S[9];
    for (int k = 0; k < 3; ++k) p[k] = _mm256_loadu_ps(planes[k] + i);
    for (int k = 0; k < 3; ++k) mu[k] = _mm256_loadu_ps(planes[3 + k] + i);
    for (int k = 0; k < 9; ++k) S[k] = _mm256_loadu_ps(planes[6 + k] + i);
    /* u_warped = R u + t; e = (u_warped, p.z) - mu  (:105-112) */
    __m256 e[3];
    e[0] = _mm256_sub_ps(_mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(R00, p[0]), _mm256_mul_ps(R01, p[1])), t0), mu[0]);
    e[1] = _mm256_sub_ps(_mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(R10, p[0]), _mm256_mul_ps(R11, p[1])), t1), mu[1]);
    e[2] = _mm256_sub_ps(p[2], mu[2]);
    __m256 r[3];
    for (int a = 0; a < 3; ++a)
      r[a] = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(S[3 * a], e[0]), _mm256_mul_ps(S[3 * a + 1], e[1])),
                           _mm256_mul_ps(S[3 * a + 2], e[2]));
    /* R_skew_p (:120-121) and J = [A | A k ; c | c k]  (:123-127) */
    const __m256 k0 = _mm256_sub_ps(_mm256_mul_ps(R01, p[0]), _mm256_mul_ps(R00, p[1]));
    const __m256 k1 = _mm256_sub_ps(_mm256_mul_ps(R11, p[0]), _mm256_mul_ps(R10, p[1]));
    __m256 J[9];
    for (int a = 0; a < 3; ++a) {
      J[3 * a] = S[3 * a];
      J[3 * a + 1] = S[3 * a + 1];
      J[3 * a + 2] = _mm256_add_ps(_mm256_mul_ps(S[3 * a], k0), _mm256_mul_ps(S[3 * a + 1], k1));
    }
    const __m256 sq = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(r[0], r[0]), _mm256_mul_ps(r[1], r[1])),
                                    _mm256_mul_ps(r[2], r[2]));
    __m256 loss = sq, weight = _mm256_set1_ps(1.0f);
    if (loss_kind != NLO_ORACLE_LOSS_NONE) { /* per-lane scalar loss, :130-143 */
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
    for (int c = 0; c < 3; ++c) {
      const __m256 jr = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(J[c], r[0]), _mm256_mul_ps(J[3 + c], r[1])),
                                      _mm256_mul_ps(J[6 + c], r[2]));
      accg[c] = _mm256_add_ps(accg[c], _mm256_mul_ps(jr, weight));
    }
    int idx = 0;
    for (int a = 0; a < 3; ++a)
      for (int b = a; b < 3; ++b) {
        const __m256 jj = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(J[a], J[b]), _mm256_mul_ps(J[3 + a], J[3 + b])),
                                        _mm256_mul_ps(J[6 + a], J[6 + b]));
        accH[idx] = _mm256_add_ps(accH[idx], _mm256_mul_ps(weight, jj));
        ++idx;
      }
    acc_cost = _mm256_add_ps(acc_cost, loss);
  }
  for (int k = 0; k < 6; ++k) out->H[k] = HorizontalSum(accH[k]);
  for (int k = 0; k < 3; ++k) out->g[k] = HorizontalSum(accg[k]);
  out->cost = HorizontalSum(acc_cost);
}

/* reprojection twin: reprojection_error_minimizer_analytic_simd.cc:55-137.  planes: X Y Z px py. */
void AssembleRangeReproj(const float* const planes[5], int64_t begin, int64_t end, const float R[9],
                         const float t[3], const float K[6], int loss_kind, const double loss_params[2],
                         LaneSums* out) {
  __m256 Rv[9], tv[3];
  for (int k = 0; k < 9; ++k) Rv[k] = _mm256_set1_ps(R[k]);
  for (int k = 0; k < 3; ++k) tv[k] = _mm256_set1_ps(t[k]);
  const __m256 inv_fx = _mm256_set1_ps(1.0f / K[0]), inv_fy = _mm256_set1_ps(1.0f / K[1]), /* :29-30 */
               cx = _mm256_set1_ps(K[2]), cy = _mm256_set1_ps(K[3]);
  __m256 accH[21], accg[6], acc_cost = _mm256_setzero_ps();
  for (auto& v : accH) v = _mm256_setzero_ps();
  for (auto& v : accg) v = _mm256_setzero_ps();
  const __m256 zero = _mm256_setzero_ps(), one = _mm256_set1_ps(1.0f);
  for (int64_t i = begin; i + 8 <= end; i += 8) {
    __m256 X[3];
    for (int k = 0; k < 3; ++k) X[k] = _mm256_loadu_ps(planes[k] + i);
    const __m256 px = _mm256_loadu_ps(planes[3] + i), py = _mm256_loadu_ps(planes[4] + i);
    __m256 Xw[3];
    for (int a = 0; a < 3; ++a)
      Xw[a] = _mm256_add_ps(_mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(Rv[3 * a], X[0]), _mm256_mul_ps(Rv[3 * a + 1], X[1])),
                                          _mm256_mul_ps(Rv[3 * a + 2], X[2])), tv[a]);
    const __m256 is_nonzero = _mm256_and_ps(_mm256_cmp_ps(Xw[2], zero, _CMP_GT_OQ), one); /* :64 */
    const __m256 inv_zw = _mm256_div_ps(one, Xw[2]);
    __m256 r[2];
    r[0] = _mm256_sub_ps(_mm256_mul_ps(Xw[0], inv_zw), _mm256_mul_ps(inv_fx, _mm256_sub_ps(px, cx)));
    r[1] = _mm256_sub_ps(_mm256_mul_ps(Xw[1], inv_zw), _mm256_mul_ps(inv_fy, _mm256_sub_ps(py, cy)));
    /* `sq_r__ = r__.norm()` (:70): the NORM goes to the loss, as the reference writes it */
    const __m256 sq = _mm256_sqrt_ps(_mm256_add_ps(_mm256_mul_ps(r[0], r[0]), _mm256_mul_ps(r[1], r[1])));
    __m256 loss = sq, weight = one;
    if (loss_kind != NLO_ORACLE_LOSS_NONE) { /* :73-86 */
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
    weight = _mm256_mul_ps(weight, is_nonzero);
    /* -R [X]x, :89-98 */
    __m256 M[9];
    for (int a = 0; a < 3; ++a) {
      M[3 * a + 0] = _mm256_sub_ps(_mm256_mul_ps(Rv[3 * a + 2], X[1]), _mm256_mul_ps(Rv[3 * a + 1], X[2]));
      M[3 * a + 1] = _mm256_sub_ps(_mm256_mul_ps(Rv[3 * a + 0], X[2]), _mm256_mul_ps(Rv[3 * a + 2], X[0]));
      M[3 * a + 2] = _mm256_sub_ps(_mm256_mul_ps(Rv[3 * a + 1], X[0]), _mm256_mul_ps(Rv[3 * a + 0], X[1]));
    }
    const __m256 inv_zwzw = _mm256_mul_ps(inv_zw, inv_zw);
    const __m256 xw_i = _mm256_mul_ps(Xw[0], inv_zwzw), yw_i = _mm256_mul_ps(Xw[1], inv_zwzw);
    __m256 J[12]; /* 2 x 6, :100-118 */
    J[0] = inv_zw; J[1] = zero; J[2] = _mm256_sub_ps(zero, xw_i);
    J[6] = zero; J[7] = inv_zw; J[8] = _mm256_sub_ps(zero, yw_i);
    for (int c = 0; c < 3; ++c) {
      J[3 + c] = _mm256_sub_ps(_mm256_mul_ps(inv_zw, M[c]), _mm256_mul_ps(xw_i, M[6 + c]));
      J[9 + c] = _mm256_sub_ps(_mm256_mul_ps(inv_zw, M[3 + c]), _mm256_mul_ps(yw_i, M[6 + c]));
    }
    for (int c = 0; c < 6; ++c) {
      const __m256 jr = _mm256_add_ps(_mm256_mul_ps(J[c], r[0]), _mm256_mul_ps(J[6 + c], r[1]));
      accg[c] = _mm256_add_ps(accg[c], _mm256_mul_ps(jr, weight));
    }
    int idx = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) {
        for (int kk = 0; kk < 2; ++kk) /* :124-130: one weighted product per residual row */
          accH[idx] = _mm256_add_ps(accH[idx], _mm256_mul_ps(_mm256_mul_ps(J[6 * kk + a], J[6 * kk + b]), weight));
        ++idx;
      }
    acc_cost = _mm256_add_ps(acc_cost, loss);
  }
  for (int k = 0; k < 21; ++k) out->H[k] = HorizontalSum(accH[k]);
  for (int k = 0; k < 6; ++k) out->g[k] = HorizontalSum(accg[k]);
  out->cost = HorizontalSum(acc_cost);
}

/* [b, e) of thread k of T on 8-lane stride boundaries (..._analytic_simd.cc:59-68; T == 1: everything) */
inline void StrideRange(int64_t n, int k, int T, int64_t* b, int64_t* e) {
  const int64_t num_stride = n / 8;
  const int64_t per = num_stride / T;
  *b = 8 * per * k;
  *e = (T == 1) ? 8 * num_stride : 8 * per * (k + 1);
}

}  // namespace

extern "C" {

void nlo_oracle_simd_ndt3_assemble(int64_t n, const float* planes, const double R2_rowmajor[4],
                                   const double t2[2], int loss_kind, const double loss_params[2],
                                   int num_threads, double H6[6], double g[3], double* cost) {
  const float* plane_ptr[15];
  for (int k = 0; k < 15; ++k) plane_ptr[k] = planes + static_cast<int64_t>(k) * n;
  float Rf[4], tf[2];
  for (int k = 0; k < 4; ++k) Rf[k] = static_cast<float>(R2_rowmajor[k]);
  for (int k = 0; k < 2; ++k) tf[k] = static_cast<float>(t2[k]);
  const int T = std::max(1, num_threads);
  std::vector<LaneSums3> parts(T);
  std::vector<std::thread> workers;
  for (int k = 0; k < T; ++k) {
    int64_t b, e;
    StrideRange(n, k, T, &b, &e);
    if (T == 1) AssembleRange3(plane_ptr, b, e, Rf, tf, loss_kind, loss_params, &parts[k]);
    else workers.emplace_back([&, k, b, e]() { AssembleRange3(plane_ptr, b, e, Rf, tf, loss_kind, loss_params, &parts[k]); });
  }
  for (auto& w : workers) w.join();
  for (int k = 0; k < 6; ++k) H6[k] = 0.0;
  for (int k = 0; k < 3; ++k) g[k] = 0.0;
  *cost = 0.0;
  for (const auto& part : parts) {
    for (int k = 0; k < 6; ++k) H6[k] += part.H[k];
    for (int k = 0; k < 3; ++k) g[k] += part.g[k];
    *cost += part.cost;
  }
}

/* local_point[3n], pixel[2n] (double) -> 5 float planes X Y Z px py (plane k at k*n): the per-Solve
 * conversion of reprojection_error_minimizer_analytic_simd.cc:19-27 */
void nlo_oracle_simd_reproj_pack(int64_t n, const double* local_point, const double* pixel, float* planes) {
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) planes[k * n + i] = static_cast<float>(local_point[3 * i + k]);
    for (int k = 0; k < 2; ++k) planes[(3 + k) * n + i] = static_cast<float>(pixel[2 * i + k]);
  }
}

void nlo_oracle_simd_reproj_assemble(int64_t n, const float* planes, const double intrinsics[6],
                                     const double R_rowmajor[9], const double t[3], int loss_kind,
                                     const double loss_params[2], int num_threads, double H21[21],
                                     double g[6], double* cost) {
  const float* plane_ptr[5];
  for (int k = 0; k < 5; ++k) plane_ptr[k] = planes + static_cast<int64_t>(k) * n;
  float Rf[9], tf[3], Kf[6];
  for (int k = 0; k < 9; ++k) Rf[k] = static_cast<float>(R_rowmajor[k]);
  for (int k = 0; k < 3; ++k) tf[k] = static_cast<float>(t[k]);
  for (int k = 0; k < 6; ++k) Kf[k] = static_cast<float>(intrinsics[k]);
  const int T = std::max(1, num_threads);
  std::vector<LaneSums> parts(T);
  std::vector<std::thread> workers;
  for (int k = 0; k < T; ++k) {
    int64_t b, e;
    StrideRange(n, k, T, &b, &e);
    if (T == 1) AssembleRangeReproj(plane_ptr, b, e, Rf, tf, Kf, loss_kind, loss_params, &parts[k]);
    else workers.emplace_back([&, k, b, e]() { AssembleRangeReproj(plane_ptr, b, e, Rf, tf, Kf, loss_kind, loss_params, &parts[k]); });
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
