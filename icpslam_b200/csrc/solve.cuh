// solve.cuh — K4: the O(1) tail of one ICP iteration, executed by one lane of the last CTA of a sweep.
//
//   umeyama_solve      pcl::registration::TransformationEstimationSVD -> Eigen::umeyama (no scaling)
//                      from the 16 running sums: centroids, 3x3 cross-covariance,
//                      3x3 SVD by one-sided Jacobi rotations, reflection fix, R = U S V^T, t = dm - R sm.
//   p2p_finish_iteration   `final = transformation_ * final; ++nr_iterations_;` followed by
//                      pcl::registration::DefaultConvergenceCriteria::hasConverged (PCL 1.8.x).
// Reference call sites: icp.align() at src/icpslam/icp_odometer.cpp:198 and
// src/icpslam/octree_mapper.cpp:114; hasConverged() at icp_odometer.cpp:201 / octree_mapper.cpp:117.
//
// Everything here is a serial fp64 dependency chain on the critical path of every iteration, so the
// 3x3 matrices live in named registers (no dynamically indexed arrays -> no local memory) and each
// Jacobi rotation costs one sqrt, one division and one rsqrt.
#pragma once
#include "common.cuh"

namespace b2 {

struct Mat3 {  // row-major 3x3 in registers
  double a00, a01, a02, a10, a11, a12, a20, a21, a22;
};

__device__ __forceinline__ double det3(const Mat3& M) {
  return M.a00 * (M.a11 * M.a22 - M.a12 * M.a21) - M.a01 * (M.a10 * M.a22 - M.a12 * M.a20) +
         M.a02 * (M.a10 * M.a21 - M.a11 * M.a20);
}

// One one-sided Jacobi rotation making columns (p, q) of B orthogonal; the same rotation is applied
// to the columns of W.  Columns are passed by reference as three scalars each.
__device__ __forceinline__ bool jacobi_rotate(double& bp0, double& bp1, double& bp2, double& bq0, double& bq1,
                                              double& bq2, double& wp0, double& wp1, double& wp2, double& wq0,
                                              double& wq1, double& wq2) {
  const double al = bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
  const double be = bq0 * bq0 + bq1 * bq1 + bq2 * bq2;
  const double ga = bp0 * bq0 + bp1 * bq1 + bp2 * bq2;
  if (ga == 0.0 || ga * ga <= 1e-30 * (al * be)) return false;  // |cos| <= 1e-15: below that a rotation is a no-op in fp64
  // t = tan(theta) = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (be - al) / (2 ga)
  //   = 2 ga * sign / (|be - al| + sqrt((be - al)^2 + 4 ga^2))            (one sqrt, one division)
  const double diff = be - al;
  const double sgn = (diff >= 0.0) == (ga >= 0.0) ? 1.0 : -1.0;
  const double t = sgn * 2.0 * fabs(ga) / (fabs(diff) + sqrt(diff * diff + 4.0 * ga * ga));
  const double c = rsqrt(1.0 + t * t), s = c * t;
  double x, y;
  x = c * bp0 - s * bq0; y = s * bp0 + c * bq0; bp0 = x; bq0 = y;
  x = c * bp1 - s * bq1; y = s * bp1 + c * bq1; bp1 = x; bq1 = y;
  x = c * bp2 - s * bq2; y = s * bp2 + c * bq2; bp2 = x; bq2 = y;
  x = c * wp0 - s * wq0; y = s * wp0 + c * wq0; wp0 = x; wq0 = y;
  x = c * wp1 - s * wq1; y = s * wp1 + c * wq1; wp1 = x; wq1 = y;
  x = c * wp2 - s * wq2; y = s * wp2 + c * wq2; wp2 = x; wq2 = y;
  return true;
}

#define B2_SWAP(a, b) { double t__ = a; a = b; b = t__; }

// A = U diag(s) V^T with s0 >= s1 >= s2 >= 0, U and V orthogonal (JacobiSVD FullU|FullV semantics).
__device__ __forceinline__ void svd3_jacobi(const Mat3& A, Mat3& U, double& s0, double& s1, double& s2, Mat3& V) {
  // B = A V accumulates; columns of B: (b00,b10,b20), (b01,b11,b21), (b02,b12,b22)
  double b00 = A.a00, b01 = A.a01, b02 = A.a02, b10 = A.a10, b11 = A.a11, b12 = A.a12, b20 = A.a20, b21 = A.a21,
         b22 = A.a22;
  double w00 = 1, w01 = 0, w02 = 0, w10 = 0, w11 = 1, w12 = 0, w20 = 0, w21 = 0, w22 = 1;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool r0 = jacobi_rotate(b00, b10, b20, b01, b11, b21, w00, w10, w20, w01, w11, w21);  // (0,1)
    bool r1 = jacobi_rotate(b00, b10, b20, b02, b12, b22, w00, w10, w20, w02, w12, w22);  // (0,2)
    bool r2 = jacobi_rotate(b01, b11, b21, b02, b12, b22, w01, w11, w21, w02, w12, w22);  // (1,2)
    if (!(r0 || r1 || r2)) break;
  }
  double n0 = sqrt(b00 * b00 + b10 * b10 + b20 * b20);
  double n1 = sqrt(b01 * b01 + b11 * b11 + b21 * b21);
  double n2 = sqrt(b02 * b02 + b12 * b12 + b22 * b22);
  // stable descending order of the three columns (bubble network, strict >)
  if (n1 > n0) { B2_SWAP(n0, n1) B2_SWAP(b00, b01) B2_SWAP(b10, b11) B2_SWAP(b20, b21) B2_SWAP(w00, w01) B2_SWAP(w10, w11) B2_SWAP(w20, w21) }
  if (n2 > n1) { B2_SWAP(n1, n2) B2_SWAP(b01, b02) B2_SWAP(b11, b12) B2_SWAP(b21, b22) B2_SWAP(w01, w02) B2_SWAP(w11, w12) B2_SWAP(w21, w22) }
  if (n1 > n0) { B2_SWAP(n0, n1) B2_SWAP(b00, b01) B2_SWAP(b10, b11) B2_SWAP(b20, b21) B2_SWAP(w00, w01) B2_SWAP(w10, w11) B2_SWAP(w20, w21) }
  s0 = n0; s1 = n1; s2 = n2;
  V = Mat3{w00, w01, w02, w10, w11, w12, w20, w21, w22};
  const double tiny = 1e-300 + n0 * 1e-14;
  double u00 = 1, u10 = 0, u20 = 0, u01 = 0, u11 = 1, u21 = 0, u02 = 0, u12 = 0, u22 = 1;
  if (n0 > tiny) {
    const double i0 = 1.0 / n0;
    u00 = b00 * i0; u10 = b10 * i0; u20 = b20 * i0;
    if (n1 > tiny) {
      const double i1 = 1.0 / n1;
      u01 = b01 * i1; u11 = b11 * i1; u21 = b21 * i1;
    } else {
      // rank 1: any unit vector orthogonal to u0 (cross with the axis of its smallest component)
      double ex = 0, ey = 0, ez = 0;
      const double ax = fabs(u00), ay = fabs(u10), az = fabs(u20);
      if (ax <= ay && ax <= az) ex = 1; else if (ay <= az) ey = 1; else ez = 1;
      u01 = u10 * ez - u20 * ey; u11 = u20 * ex - u00 * ez; u21 = u00 * ey - u10 * ex;
      const double in = rsqrt(u01 * u01 + u11 * u11 + u21 * u21);
      u01 *= in; u11 *= in; u21 *= in;
    }
    if (n2 > tiny) {
      const double i2 = 1.0 / n2;
      u02 = b02 * i2; u12 = b12 * i2; u22 = b22 * i2;
    } else {
      // rank <= 2 (planar clouds): complete with u0 x u1
      u02 = u10 * u21 - u20 * u11; u12 = u20 * u01 - u00 * u21; u22 = u00 * u11 - u10 * u01;
      const double in = rsqrt(u02 * u02 + u12 * u12 + u22 * u22);
      u02 *= in; u12 *= in; u22 *= in;
    }
  }
  U = Mat3{u00, u01, u02, u10, u11, u12, u20, u21, u22};
}
#undef B2_SWAP

// sums: [0] n, [1..3] sum(src), [4..6] sum(dst), [7..15] sum(dst_r * src_c) row-major, [16] sum(d2)
// T16: row-major 4x4 double
__device__ __forceinline__ void umeyama_solve(const double* S, double* T16) {
  const double inv_n = 1.0 / S[0];
  const double sm0 = S[1] * inv_n, sm1 = S[2] * inv_n, sm2 = S[3] * inv_n;
  const double dm0 = S[4] * inv_n, dm1 = S[5] * inv_n, dm2 = S[6] * inv_n;
  Mat3 sigma;
  sigma.a00 = S[7] * inv_n - dm0 * sm0;  sigma.a01 = S[8] * inv_n - dm0 * sm1;  sigma.a02 = S[9] * inv_n - dm0 * sm2;
  sigma.a10 = S[10] * inv_n - dm1 * sm0; sigma.a11 = S[11] * inv_n - dm1 * sm1; sigma.a12 = S[12] * inv_n - dm1 * sm2;
  sigma.a20 = S[13] * inv_n - dm2 * sm0; sigma.a21 = S[14] * inv_n - dm2 * sm1; sigma.a22 = S[15] * inv_n - dm2 * sm2;
  Mat3 U, V;
  double s0, s1, s2;
  svd3_jacobi(sigma, U, s0, s1, s2, V);
  const double g2 = (det3(U) * det3(V) < 0) ? -1.0 : 1.0;  // Eigen >= 3.3 reflection fix on the last column
  // R = U diag(1,1,g2) V^T
  const double r00 = U.a00 * V.a00 + U.a01 * V.a01 + g2 * U.a02 * V.a02;
  const double r01 = U.a00 * V.a10 + U.a01 * V.a11 + g2 * U.a02 * V.a12;
  const double r02 = U.a00 * V.a20 + U.a01 * V.a21 + g2 * U.a02 * V.a22;
  const double r10 = U.a10 * V.a00 + U.a11 * V.a01 + g2 * U.a12 * V.a02;
  const double r11 = U.a10 * V.a10 + U.a11 * V.a11 + g2 * U.a12 * V.a12;
  const double r12 = U.a10 * V.a20 + U.a11 * V.a21 + g2 * U.a12 * V.a22;
  const double r20 = U.a20 * V.a00 + U.a21 * V.a01 + g2 * U.a22 * V.a02;
  const double r21 = U.a20 * V.a10 + U.a21 * V.a11 + g2 * U.a22 * V.a12;
  const double r22 = U.a20 * V.a20 + U.a21 * V.a21 + g2 * U.a22 * V.a22;
  T16[0] = r00; T16[1] = r01; T16[2] = r02;  T16[3] = dm0 - (r00 * sm0 + r01 * sm1 + r02 * sm2);
  T16[4] = r10; T16[5] = r11; T16[6] = r12;  T16[7] = dm1 - (r10 * sm0 + r11 * sm1 + r12 * sm2);
  T16[8] = r20; T16[9] = r21; T16[10] = r22; T16[11] = dm2 - (r20 * sm0 + r21 * sm1 + r22 * sm2);
  T16[12] = 0.0; T16[13] = 0.0; T16[14] = 0.0; T16[15] = 1.0;
}

// Eigen Matrix4f product: each entry accumulated k = 0..3 in float.
__device__ __forceinline__ void mat4f_mul(const float* A, const float* B, float* C) {
  float t[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) s = fadd(s, fmul(A[4 * r + k], B[4 * k + c]));
      t[4 * r + c] = s;
    }
#pragma unroll
  for (int i = 0; i < 16; ++i) C[i] = t[i];
}

// One lane: consume the reduced sums of a point-to-point sweep and advance the loop state.
__device__ __noinline__ void p2p_finish_iteration(const double* S, IcpState* st, const IcpConfig& cfg) {
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
    return;
  }
  double Td[16];
  umeyama_solve(S, Td);
  float Tinc[16], fin[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    Tinc[i] = (float)Td[i];  // transformation_ is a Matrix4f
    fin[i] = st->final_T[i];
  }
  mat4f_mul(Tinc, fin, fin);
#pragma unroll
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
    const double cos_angle = 0.5 * (double)fsub(fadd(fadd(Tinc[0], Tinc[5]), Tinc[10]), 1.0f);
    const double tr2 = (double)fadd(fadd(fmul(Tinc[3], Tinc[3]), fmul(Tinc[7], Tinc[7])), fmul(Tinc[11], Tinc[11]));
    const double prev = st->prev_mse;
    if (cos_angle >= cfg.rot_thresh && tr2 <= cfg.trans_thresh) {
      conv = 1;
    } else if (fabs(mse - prev) < cfg.mse_abs) {
      conv = 1;
    } else if (fabs(mse - prev) / prev < cfg.mse_rel) {
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
