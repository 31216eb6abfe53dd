// solve.cuh — K4: the O(1) tail of one ICP iteration, executed by one lane of the last CTA of a sweep.
//
//   umeyama_solve      pcl::registration::TransformationEstimationSVD -> Eigen::umeyama (no scaling)
//                      from the 16 running sums: centroids, 3x3 cross-covariance,
//                      3x3 SVD by one-sided Jacobi rotations, reflection fix, R = U S V^T, t = dm - R sm.
//   p2p_finish_iteration   `final = transformation_ * final; ++nr_iterations_;` followed by
//                      pcl::registration::DefaultConvergenceCriteria::hasConverged (PCL 1.8.x).
// Reference call sites: icp.align() at src/icpslam/icp_odometer.cpp:198 and
// src/icpslam/octree_mapper.cpp:114; hasConverged() at icp_odometer.cpp:201 / octree_mapper.cpp:117.
#pragma once
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
         M[2] * (M[3] * M[7] - M[4] * M[6]);
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// A (row-major 3x3) = U diag(s) V^T, s descending, U and V full orthogonal matrices.
__device__ void svd3_jacobi(const double* A, double* U, double* s, double* V) {
  double B[9], W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    B[i] = A[i];
    W[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      double al = 0, be = 0, ga = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        al += B[3 * i + p] * B[3 * i + p];
        be += B[3 * i + q] * B[3 * i + q];
        ga += B[3 * i + p] * B[3 * i + q];
      }
      if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
      rotated = true;
      double zeta = (be - al) / (2.0 * ga);
      double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double bp = B[3 * i + p], bq = B[3 * i + q];
        B[3 * i + p] = c * bp - sn * bq;
        B[3 * i + q] = sn * bp + c * bq;
        double wp = W[3 * i + p], wq = W[3 * i + q];
        W[3 * i + p] = c * wp - sn * wq;
        W[3 * i + q] = sn * wp + c * wq;
      }
    }
    if (!rotated) break;
  }
  double nrm[3];
  int ord[3] = {0, 1, 2};
#pragma unroll
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  // stable descending sort of three
#define B2_SWAP_IF(a, b)                                                          \
  if (nrm[ord[b]] > nrm[ord[a]]) {                                                \
    int tmp = ord[a];                                                             \
    ord[a] = ord[b];                                                              \
    ord[b] = tmp;                                                                 \
  }
  B2_SWAP_IF(0, 1) B2_SWAP_IF(1, 2) B2_SWAP_IF(0, 1)
#undef B2_SWAP_IF
  double u[3][3], v[3][3];
  for (int j = 0; j < 3; ++j) {
    int c = ord[j];
    s[j] = nrm[c];
    for (int i = 0; i < 3; ++i) v[j][i] = W[3 * i + c];
  }
  const double tiny = 1e-300 + s[0] * 1e-14;
  int rank = 0;
  for (int j = 0; j < 3; ++j)
    if (s[j] > tiny) {
      int c = ord[j];
      for (int i = 0; i < 3; ++i) u[j][i] = B[3 * i + c] / s[j];
      rank = j + 1;
    }
  if (rank == 0) {
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) u[j][i] = (i == j) ? 1.0 : 0.0;
  } else if (rank == 1) {
    double e[3] = {0, 0, 0};
    int m = 0;
    for (int i = 1; i < 3; ++i)
      if (fabs(u[0][i]) < fabs(u[0][m])) m = i;
    e[m] = 1.0;
    cross3(u[0], e, u[1]);
    double n1 = sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
    for (int i = 0; i < 3; ++i) u[1][i] /= n1;
    cross3(u[0], u[1], u[2]);
  } else if (rank == 2) {
    cross3(u[0], u[1], u[2]);
    double n2 = sqrt(u[2][0] * u[2][0] + u[2][1] * u[2][1] + u[2][2] * u[2][2]);
    for (int i = 0; i < 3; ++i) u[2][i] /= n2;
  }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      U[3 * i + j] = u[j][i];
      V[3 * i + j] = v[j][i];
    }
}

// sums: [0] n, [1..3] sum(src), [4..6] sum(dst), [7..15] sum(dst_r * src_c) row-major, [16] sum(d2)
__device__ void umeyama_solve(const double* S, double* T16) {
  const double n = S[0];
  double sm[3], dm[3], sigma[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    sm[i] = S[1 + i] / n;
    dm[i] = S[4 + i] / n;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) sigma[3 * r + c] = S[7 + 3 * r + c] / n - dm[r] * sm[c];
  double U[9], sv[3], V[9];
  svd3_jacobi(sigma, U, sv, V);
  double sg[3] = {1.0, 1.0, 1.0};
  if (det3(U) * det3(V) < 0) sg[2] = -1.0;
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double a = 0;
      for (int k = 0; k < 3; ++k) a += U[3 * r + k] * sg[k] * V[3 * c + k];
      R[3 * r + c] = a;
    }
  for (int i = 0; i < 16; ++i) T16[i] = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[4 * r + c] = R[3 * r + c];
    T16[4 * r + 3] = dm[r] - (R[3 * r] * sm[0] + R[3 * r + 1] * sm[1] + R[3 * r + 2] * sm[2]);
  }
  T16[15] = 1.0;
}

// Eigen Matrix4f product: each entry accumulated k = 0..3 in float.
__device__ __forceinline__ void mat4f_mul(const float* A, const float* B, float* C) {
  float t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s = fadd(s, fmul(A[4 * r + k], B[4 * k + c]));
      t[4 * r + c] = s;
    }
  for (int i = 0; i < 16; ++i) C[i] = t[i];
}

// One lane: consume the reduced sums of a point-to-point sweep and advance the loop state.
__device__ void p2p_finish_iteration(const double* S, IcpState* st, const IcpConfig& cfg) {
  const int n_corr = (int)S[0];
  st->n_corr = n_corr;
  if (*(volatile int*)&st->pad) {  // a sweep thread met a non-finite coordinate
    st->status = -6;               // B2ICP_ERR_NONFINITE_INPUT
    st->converged = 0;
    st->done = 1;
    return;
  }
  if (n_corr < cfg.min_corr) {  // `Not enough correspondences found` -> converged_ = false; break
    st->status = -4;            // B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES
    st->converged = 0;
    st->done = 1;
    for (int i = 0; i < 16; ++i) st->Tinc[i] = (i % 5 == 0) ? 1.f : 0.f;
    return;
  }
  double Td[16];
  umeyama_solve(S, Td);
  float Tinc[16];
  for (int i = 0; i < 16; ++i) Tinc[i] = (float)Td[i];  // transformation_ is a Matrix4f
  float fin[16];
  for (int i = 0; i < 16; ++i) fin[i] = st->final_T[i];
  mat4f_mul(Tinc, fin, fin);
  for (int i = 0; i < 16; ++i) {
    st->final_T[i] = fin[i];
    st->Tinc[i] = Tinc[i];
  }
  const int iters = st->iter + 1;
  st->iter = iters;
  const double mse = S[16] / S[0];
  st->mse = mse;
  int conv = 0;
  if (iters >= cfg.max_iterations) {
    conv = 1;
  } else {
    double cos_angle = 0.5 * (double)fsub(fadd(fadd(Tinc[0], Tinc[5]), Tinc[10]), 1.0f);
    double tr2 = (double)fadd(fadd(fmul(Tinc[3], Tinc[3]), fmul(Tinc[7], Tinc[7])), fmul(Tinc[11], Tinc[11]));
    if (cos_angle >= cfg.rot_thresh && tr2 <= cfg.trans_thresh) {
      conv = 1;
    } else if (fabs(mse - st->prev_mse) < cfg.mse_abs) {
      conv = 1;
    } else if (fabs(mse - st->prev_mse) / st->prev_mse < cfg.mse_rel) {
      conv = 1;
    } else {
      st->prev_mse = mse;
    }
  }
  if (conv) {
    st->converged = 1;
    st->done = 1;
  }
}

}  // namespace b2
