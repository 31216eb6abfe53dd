// gicp.cuh — the device half of pcl::GeneralizedIterativeClosestPoint, the class the reference
// instantiates (reference src/icpslam/icp_odometer.cpp:188, src/icpslam/octree_mapper.cpp:104;
// SURVEY.md App. A.2):
//   K5  knn_cov_kernel      computeCovariances: k nearest neighbours (k = 20, the point itself included),
//                           double covariance from float products, 3x3 SVD, eigenvalues -> (1, 1, eps)
//   G2  gicp_corr_kernel    per outer iteration: query = transformation_ * (guess * p) in float, exact
//                           1-NN (nn.cuh), strict gate d2 < max^2, M = (R C1 R^T + C2)^-1
//   G3  gicp_fdf_kernel     one evaluation of the BFGS cost functor (f, df in one pass): 13 sums
// The BFGS recursion itself is O(1) scalar work and runs on the host (gicp_host.inl), one fiber per scan.
//
// PCL's GICP does not survive a change in the last bit of f (its line search ends on a round-off
// test), so everything here is written to be BIT-IDENTICAL to a fixed arithmetic definition:
//   * every double operation is an explicit __dmul_rn / __dadd_rn / __dsub_rn / __ddiv_rn / __dsqrt_rn in
//     the order of the C expression it restates (no FMA contraction, no reassociation);
//   * neighbour sums run in (d2, index) order; cross-point sums use a fixed tree: 256-point CTAs, xor
//     butterfly 16,8,4,2,1 inside each warp, the 8 warp sums added in order, CTA sums added in order
//     (gicp_sum_kernel).
// Every kernel takes an array of per-scan tasks (blockIdx.y = scan): the scans of a batch advance together, one
// launch per evaluation ROUND instead of one per scan (gicp_host.inl).
#pragma once
#include "common.cuh"
#include "nn.cuh"

namespace b2 {

constexpr int kGicpThreads = 256;
constexpr int kGicpSums = 14;  // f, g0..g2, R00..R22, count
constexpr int kMaxK = 32;      // k_correspondences <= 32

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ---- 3x3 helpers, each a literal restatement of its C expression (left-to-right) -----------------
// C = A * B :   t = A[3r]*B[c] + A[3r+1]*B[3+c] + A[3r+2]*B[6+c]
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
  double t[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      t[3 * r + c] = dadd(dadd(dmul(A[3 * r], B[c]), dmul(A[3 * r + 1], B[3 + c])), dmul(A[3 * r + 2], B[6 + c]));
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}
// C = A * B^T
__device__ __forceinline__ void mat3_mul_bt(const double* A, const double* B, double* C) {
  double t[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      t[3 * r + c] =
          dadd(dadd(dmul(A[3 * r], B[3 * c]), dmul(A[3 * r + 1], B[3 * c + 1])), dmul(A[3 * r + 2], B[3 * c + 2]));
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}
// Eigen's closed-form inverse: cofactors times 1/det
__device__ __forceinline__ void mat3_inv(const double* M, double* I) {
  const double c00 = dsub(dmul(M[4], M[8]), dmul(M[5], M[7]));
  const double c01 = dsub(dmul(M[5], M[6]), dmul(M[3], M[8]));
  const double c02 = dsub(dmul(M[3], M[7]), dmul(M[4], M[6]));
  const double det = dadd(dadd(dmul(M[0], c00), dmul(M[1], c01)), dmul(M[2], c02));
  const double id = ddiv(1.0, det);
  I[0] = dmul(c00, id);
  I[1] = dmul(dsub(dmul(M[2], M[7]), dmul(M[1], M[8])), id);
  I[2] = dmul(dsub(dmul(M[1], M[5]), dmul(M[2], M[4])), id);
  I[3] = dmul(c01, id);
  I[4] = dmul(dsub(dmul(M[0], M[8]), dmul(M[2], M[6])), id);
  I[5] = dmul(dsub(dmul(M[2], M[3]), dmul(M[0], M[5])), id);
  I[6] = dmul(c02, id);
  I[7] = dmul(dsub(dmul(M[1], M[6]), dmul(M[0], M[7])), id);
  I[8] = dmul(dsub(dmul(M[0], M[4]), dmul(M[1], M[3])), id);
}

__device__ __forceinline__ void cross3r(const double* a, const double* b, double* c) {
  c[0] = dsub(dmul(a[1], b[2]), dmul(a[2], b[1]));
  c[1] = dsub(dmul(a[2], b[0]), dmul(a[0], b[2]));
  c[2] = dsub(dmul(a[0], b[1]), dmul(a[1], b[0]));
}
__device__ __forceinline__ double norm3r(const double* a) {
  return __dsqrt_rn(dadd(dadd(dmul(a[0], a[0]), dmul(a[1], a[1])), dmul(a[2], a[2])));
}

// U of the SVD A = U diag(s) V^T (s descending), by the fixed one-sided Jacobi recipe of the arithmetic
// definition: pairs (0,1), (0,2), (1,2); rotate unless ga == 0 or |ga| <= 1e-15 sqrt(al be)
// (where a rotation stops changing an fp64 column; Eigen::JacobiSVD itself stops at 2 eps);
// zeta = (be - al) / (2 ga); t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)); c = 1 / sqrt(1 + t^2); s = c t.
// one sweep over the pairs (0,1), (0,2), (1,2) of the recipe; false = no pair was rotated (converged)
__device__ __forceinline__ bool svd3_sweep(double* B) {
  bool rotated = false;
#pragma unroll
  for (int pq = 0; pq < 3; ++pq) {
    const int p = (pq == 2) ? 1 : 0;
    const int q = (pq == 0) ? 1 : 2;
    double al = 0, be = 0, ga = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      al = dadd(al, dmul(B[3 * i + p], B[3 * i + p]));
      be = dadd(be, dmul(B[3 * i + q], B[3 * i + q]));
      ga = dadd(ga, dmul(B[3 * i + p], B[3 * i + q]));
    }
    if (ga == 0.0 || fabs(ga) <= dmul(1e-15, __dsqrt_rn(dmul(al, be)))) continue;
    rotated = true;
    const double zeta = ddiv(dsub(be, al), dmul(2.0, ga));
    const double t = ddiv(zeta >= 0 ? 1.0 : -1.0, dadd(fabs(zeta), __dsqrt_rn(dadd(1.0, dmul(zeta, zeta)))));
    const double c = ddiv(1.0, __dsqrt_rn(dadd(1.0, dmul(t, t))));
    const double sn = dmul(c, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double bp = B[3 * i + p], bq = B[3 * i + q];
      B[3 * i + p] = dsub(dmul(c, bp), dmul(sn, bq));
      B[3 * i + q] = dadd(dmul(sn, bp), dmul(c, bq));
    }
  }
  return rotated;
}
constexpr int kSvdMaxSweeps = 60;

// the rest of the recipe once the sweeps are over: column norms, stable descending order, rank completion
__device__ void svd3_finish(const double* B, double* U) {
  double nrm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    nrm[j] = __dsqrt_rn(dadd(dadd(dmul(B[j], B[j]), dmul(B[3 + j], B[3 + j])), dmul(B[6 + j], B[6 + j])));
  // stable descending order of the three columns
  int o0 = 0, o1 = 1, o2 = 2;
  if (nrm[o1] > nrm[o0]) { int t = o0; o0 = o1; o1 = t; }
  if (nrm[o2] > nrm[o1]) { int t = o1; o1 = o2; o2 = t; }
  if (nrm[o1] > nrm[o0]) { int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
  double s[3], u[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) s[j] = nrm[ord[j]];
  const double tiny = dadd(1e-300, dmul(s[0], 1e-14));
  int rank = 0;
  for (int j = 0; j < 3; ++j)
    if (s[j] > tiny) {
      const int c = ord[j];
      for (int i = 0; i < 3; ++i) u[j][i] = ddiv(B[3 * i + c], s[j]);
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
    cross3r(u[0], e, u[1]);
    const double n1 = norm3r(u[1]);
    for (int i = 0; i < 3; ++i) u[1][i] = ddiv(u[1][i], n1);
    cross3r(u[0], u[1], u[2]);
  } else if (rank == 2) {
    cross3r(u[0], u[1], u[2]);
    const double n2 = norm3r(u[2]);
    for (int i = 0; i < 3; ++i) u[2][i] = ddiv(u[2][i], n2);
  }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) U[3 * i + j] = u[j][i];
}
__device__ void svd3_u(const double* A, double* U) {
  double B[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) B[i] = A[i];
  for (int sweep = 0; sweep < kSvdMaxSweeps; ++sweep)
    if (!svd3_sweep(B)) break;
  svd3_finish(B, U);
}

// ---- K5: k nearest neighbours + regularised covariance ---------------------------------------------
// One thread per point of `cloud` (original order); `g` is the grid built over the same cloud.  The k best
// (d2, index) keys are kept sorted in a per-thread array; rings of cells are visited until the distance to
// the outside of the visited block exceeds the k-th best distance.  Points that exhaust `max_rings` are
// queued for the exhaustive fallback (same arithmetic, every point scanned).
// KS > 0: k is the compile-time constant KS, the list lives in registers (every index static) and an insertion is a
// branch-free bubble of the candidate through the list — the same straight-line code for every lane.  KS == 0: k is
// a run-time value, the list is a local-memory array shifted by a loop (the general path).
template <int KS>
__device__ __forceinline__ void knn_offer(unsigned long long* keys, int k, unsigned long long c) {
  if (KS > 0) {
    if (c >= keys[KS - 1]) return;
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      const bool lt = c < keys[j];
      const unsigned long long hi = lt ? keys[j] : c;
      keys[j] = lt ? c : keys[j];
      c = hi;
    }
  } else {
    if (c >= keys[k - 1]) return;
    int p = k - 1;
    while (p > 0 && keys[p - 1] > c) {
      keys[p] = keys[p - 1];
      --p;
    }
    keys[p] = c;
  }
}

template <int KS>
__device__ __forceinline__ void knn_scan(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz,
                                         unsigned long long* keys, int k) {
  for (int j = s; j < e; ++j) {
    const float4 p = __ldg(pts + j);
    knn_offer<KS>(keys, k, pack_key(sqdist3(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w)));
  }
}

// first half of computeCovariances for one point: the raw covariance of its k neighbours (9 doubles, row-major)
template <int KS>
__device__ __forceinline__ void cov_from_keys(const float4* __restrict__ cloud, const unsigned long long* keys, int k_,
                                              double* __restrict__ out9) {
  const int k = KS > 0 ? KS : k_;
  double mean[3] = {0, 0, 0};
  double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll
  for (int j = 0; j < (KS > 0 ? KS : kMaxK); ++j) {
    if (KS == 0 && j >= k) break;
    const float4 pt = __ldg(cloud + key_idx(keys[j]));
    mean[0] = dadd(mean[0], (double)pt.x);
    mean[1] = dadd(mean[1], (double)pt.y);
    mean[2] = dadd(mean[2], (double)pt.z);
    c00 = dadd(c00, (double)fmul(pt.x, pt.x));  // float product, double accumulation (PCL)
    c10 = dadd(c10, (double)fmul(pt.y, pt.x));
    c11 = dadd(c11, (double)fmul(pt.y, pt.y));
    c20 = dadd(c20, (double)fmul(pt.z, pt.x));
    c21 = dadd(c21, (double)fmul(pt.z, pt.y));
    c22 = dadd(c22, (double)fmul(pt.z, pt.z));
  }
  const double kd = (double)k;
  for (int d = 0; d < 3; ++d) mean[d] = ddiv(mean[d], kd);
  double C[9];
  C[0] = dsub(ddiv(c00, kd), dmul(mean[0], mean[0]));
  C[3] = dsub(ddiv(c10, kd), dmul(mean[1], mean[0]));
  C[4] = dsub(ddiv(c11, kd), dmul(mean[1], mean[1]));
  C[6] = dsub(ddiv(c20, kd), dmul(mean[2], mean[0]));
  C[7] = dsub(ddiv(c21, kd), dmul(mean[2], mean[1]));
  C[8] = dsub(ddiv(c22, kd), dmul(mean[2], mean[2]));
  C[1] = C[3];
  C[2] = C[6];
  C[5] = C[7];
  for (int i = 0; i < 9; ++i) out9[i] = C[i];
}

// second half of computeCovariances for one point: U of the covariance's SVD, eigenvalues replaced by (1, 1, eps)
__device__ __forceinline__ void cov_regularise(const double* U, double eps, double* __restrict__ out9) {
  double o[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int col = 0; col < 3; ++col) {
    const double v = col == 2 ? eps : 1.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o[3 * r + c] = dadd(o[3 * r + c], dmul(dmul(v, U[3 * r + col]), U[3 * c + col]));
  }
  for (int i = 0; i < 9; ++i) out9[i] = o[i];
}

// One task = one cloud of a batch (blockIdx.y selects it): the covariances of every scan of a GICP batch are one launch.
struct KnnTask {
  GridView g;           // grid over the cloud
  const float4* cloud;  // the cloud in original order
  double* cov;          // [n][9]
  int n;
  int pad;
};

// k nearest neighbours + covariance of point i of task t (one thread); queued on `unresolved_list` when its k
// neighbours do not lie within max_rings cells of its grid.
template <int KS>
__device__ __forceinline__ void knn_cov_point(const KnnTask& t, int task_id, int i, int k_, double eps, int max_rings,
                                              int2* __restrict__ unresolved_list, unsigned int* __restrict__ unresolved_count) {
  const GridView& g = t.g;
  const float4* __restrict__ cloud = t.cloud;
  const float4 q = __ldg(cloud + i);
  constexpr int kList = KS > 0 ? KS : kMaxK;
  const int k = KS > 0 ? KS : k_;
  unsigned long long keys[kList];
#pragma unroll
  for (int j = 0; j < kList; ++j) keys[j] = kInfKey;
  const int cx = cell_coord(q.x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(q.y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(q.z, g.oz, g.inv_cell, g.nz);
  bool resolved = false;
  for (int R = 0;; ++R) {
    if (R > max_rings) break;
    const float thr = key_d2(KS > 0 ? keys[kList - 1] : keys[k - 1]);
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
    const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      const float lz = slab_gap(q.z, g.oz, g.cell, z, z, g.slack);
      const float lz2 = fmul(lz, lz);
      if (lz2 > thr) continue;
      const bool zface = (z == cz - R) || (z == cz + R);
      for (int y = y0; y <= y1; ++y) {
        const float ly = slab_gap(q.y, g.oy, g.cell, y, y, g.slack);
        if (fadd(fmul(ly, ly), lz2) > key_d2(KS > 0 ? keys[kList - 1] : keys[k - 1])) continue;
        const int row = (z * g.ny + y) * g.nx;
        if (zface || y == cy - R || y == cy + R) {
          knn_scan<KS>(g.pts, __ldg(g.cell_start + row + xa), __ldg(g.cell_start + row + xb + 1), q.x, q.y, q.z, keys, k);
        } else {
          if (cx - R >= 0)
            knn_scan<KS>(g.pts, __ldg(g.cell_start + row + cx - R), __ldg(g.cell_start + row + cx - R + 1), q.x, q.y, q.z, keys, k);
          if (R > 0 && cx + R <= g.nx - 1)
            knn_scan<KS>(g.pts, __ldg(g.cell_start + row + cx + R), __ldg(g.cell_start + row + cx + R + 1), q.x, q.y, q.z, keys, k);
        }
      }
    }
    // can anything outside the block [c - R, c + R]^3 still enter the list?
    float ex = INFINITY;
    if (cx - R > 0) ex = fminf(ex, q.x - (g.ox + (float)(cx - R) * g.cell));
    if (cx + R < g.nx - 1) ex = fminf(ex, (g.ox + (float)(cx + R + 1) * g.cell) - q.x);
    if (cy - R > 0) ex = fminf(ex, q.y - (g.oy + (float)(cy - R) * g.cell));
    if (cy + R < g.ny - 1) ex = fminf(ex, (g.oy + (float)(cy + R + 1) * g.cell) - q.y);
    if (cz - R > 0) ex = fminf(ex, q.z - (g.oz + (float)(cz - R) * g.cell));
    if (cz + R < g.nz - 1) ex = fminf(ex, (g.oz + (float)(cz + R + 1) * g.cell) - q.z);
    if (ex == INFINITY) {
      resolved = true;
      break;
    }
    ex = fmaxf(ex - g.slack, 0.0f);
    if (fmul(ex, ex) > key_d2(KS > 0 ? keys[kList - 1] : keys[k - 1])) {
      resolved = true;
      break;
    }
  }
  if (!resolved) {
    unresolved_list[atomicAdd(unresolved_count, 1u)] = make_int2(task_id, i);
    return;
  }
  cov_from_keys<KS>(cloud, keys, k, t.cov + (size_t)9 * i);
}



// every point of every cloud of the batch (blockIdx.y = cloud), on the cloud's FINE grid
template <int KS>
__global__ void __launch_bounds__(128) knn_cov_kernel(const KnnTask* __restrict__ tasks, int k, double eps, int max_rings,
                                                      int2* __restrict__ unresolved_list,
                                                      unsigned int* __restrict__ unresolved_count) {
  const KnnTask& t = tasks[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= t.n) return;
  // threads take the points in CELL order (the grid is over the same cloud): the lanes of a warp walk the same
  // cells, their loads coalesce and their trip counts agree; the result lands at the point's original index
  const int i = __float_as_int(__ldg(t.g.pts + j).w);
  knn_cov_point<KS>(t, (int)blockIdx.y, i, k, eps, max_rings, unresolved_list, unresolved_count);
}

// second pass: the points the fine grid could not settle (the sparse far field of a sweep: their 20 neighbours lie
// metres away, dozens of fine cells) on a COARSE grid over the same cloud (cell x 4: the same ring budget reaches 4x
// as far for the same number of rows).  The k-NN is exact on either grid, so the result does not depend on which
// pass settled a point.
template <int KS>
__global__ void __launch_bounds__(128) knn_cov_list_kernel(const KnnTask* __restrict__ coarse_tasks, int k, double eps, int max_rings,
                                                           const int2* __restrict__ list, const unsigned int* __restrict__ count,
                                                           int2* __restrict__ unresolved_list,
                                                           unsigned int* __restrict__ unresolved_count) {
  const unsigned int total = *count;
  for (unsigned int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    const int2 e = list[w];
    knn_cov_point<KS>(coarse_tasks[e.x], e.x, e.y, k, eps, max_rings, unresolved_list, unresolved_count);
  }
}

// Exhaustive fallback for the points whose k neighbours do not lie within `max_rings` cells (the sparse fringe of a
// sweep): one CTA per queued point.  Every thread keeps the k best keys of its share of the cloud, then the CTA
// extracts the k smallest of all of them in ascending order (k rounds of a block-wide minimum over the threads' list
// heads).  Keys (d2, index) are unique, so the result is the same sorted list a sequential scan would produce.
// (Round 1 gave every queued point to ONE thread: 65 536 sequential offers, 10.8 ms per cloud whatever the count.)
constexpr int kKnnFbThreads = 256;
__global__ void __launch_bounds__(kKnnFbThreads) knn_cov_fallback(const KnnTask* __restrict__ tasks, int k, double eps,
                                                                  const int2* __restrict__ list,
                                                                  const unsigned int* __restrict__ count) {
  __shared__ unsigned long long s_min[kKnnFbThreads / 32];
  __shared__ unsigned long long s_win;
  __shared__ unsigned long long s_keys[kMaxK];
  const unsigned int total = *count;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned int w = blockIdx.x; w < total; w += gridDim.x) {
    const KnnTask& t = tasks[list[w].x];
    const int i = list[w].y;
    const float4 q = __ldg(t.cloud + i);
    unsigned long long keys[kMaxK];
    for (int j = 0; j < kMaxK; ++j) keys[j] = kInfKey;
    for (int j = threadIdx.x; j < t.g.n; j += kKnnFbThreads) {
      const float4 p = __ldg(t.g.pts + j);
      knn_offer<0>(keys, k, pack_key(sqdist3(q.x, q.y, q.z, p.x, p.y, p.z), __float_as_int(p.w)));
    }
    int head = 0;
    for (int r = 0; r < k; ++r) {
      unsigned long long mine = head < k ? keys[head] : kInfKey;
      unsigned long long m = warp_min_key(mine);
      if (lane == 0) s_min[warp] = m;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long best = s_min[0];
        for (int x = 1; x < kKnnFbThreads / 32; ++x) best = s_min[x] < best ? s_min[x] : best;
        s_win = best;
        s_keys[r] = best;
      }
      __syncthreads();
      if (mine == s_win && mine != kInfKey) ++head;  // keys are unique: exactly one thread owns the winner
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long out[kMaxK];
      for (int j = 0; j < k; ++j) out[j] = s_keys[j];
      cov_from_keys<0>(t.cloud, out, k, t.cov + (size_t)9 * i);
    }
    __syncthreads();
  }
}

// Last stage of the covariances: every raw covariance left in `cov` by the three passes above becomes
// U diag(1, 1, eps) U^T in place.  The Jacobi recipe needs anything from a few sweeps to its cap of 60 depending on
// the matrix, so a thread per point leaves most lanes of a warp waiting for its slowest matrix (ncu: 8 of 32 lanes
// active, 47 % of the covariance kernel's instructions).  Here the lanes of a warp advance their matrices ONE sweep at a
// time and a lane whose matrix has converged finishes it and takes the next point of its cloud from a counter
// (warp-aggregated atomic), so the sweep code always runs on a full warp.  The arithmetic of a point is the same
// sequence of operations as svd3_u().
__global__ void __launch_bounds__(128) cov_svd_kernel(const KnnTask* __restrict__ tasks, double eps,
                                                      unsigned int* __restrict__ counters) {
  const KnnTask& t = tasks[blockIdx.y];
  unsigned int* ctr = counters + blockIdx.y;
  const int lane = threadIdx.x & 31;
  const unsigned int n = (unsigned int)t.n;
  double B[9];
  unsigned int i = 0;
  int sweep = 0;
  bool active = false, done = false;
  for (;;) {
    const unsigned int need = __ballot_sync(0xFFFFFFFFu, !active && !done);
    if (need) {
      const int leader = __ffs(need) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(ctr, (unsigned int)__popc(need));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (!active && !done) {
        i = base + (unsigned int)__popc(need & ((1u << lane) - 1u));
        if (i < n) {
          const double* C = t.cov + (size_t)9 * i;
#pragma unroll
          for (int e = 0; e < 9; ++e) B[e] = C[e];
          sweep = 0;
          active = true;
        } else {
          done = true;
        }
      }
    }
    if (__all_sync(0xFFFFFFFFu, done)) break;
    bool rotated = false;
    if (active) rotated = svd3_sweep(B);
    if (active && (!rotated || ++sweep == kSvdMaxSweeps)) {
      double U[9];
      svd3_finish(B, U);
      cov_regularise(U, eps, t.cov + (size_t)9 * i);
      active = false;
    }
  }
}

// ---- G2: correspondences + Mahalanobis matrices of one outer iteration -------------------------------
struct GicpIterArgs {
  float guess[16];   // base_transformation_
  float T[16];       // transformation_
  double R[9];       // top-left 3x3 of double(transformation_) * double(guess)
  double max2;       // corr_dist_threshold_^2, gate is STRICT <
  float bound2;
  int max_rings;
  int use_seed;
};

// One task = one scan of a GICP batch (blockIdx.y selects it): every scan of the batch advances with the same launch.
struct GicpCorrTask {
  GridView g;
  const float4* src;
  const double* cov_src;
  const double* cov_tgt;
  double* mahal;
  int* corr_idx;
  float* corr_d2;
  int* corr_pos;
  int n;
  int pad;
  GicpIterArgs a;
};

__global__ void __launch_bounds__(kSweepThreads, kSweepMinCtas) gicp_corr_kernel(const GicpCorrTask* __restrict__ tasks) {
  __shared__ NNScratch<kSweepThreads> sc;
  const GicpCorrTask& t = tasks[blockIdx.y];
  const GicpIterArgs& a = t.a;
  const int i = blockIdx.x * kSweepThreads + threadIdx.x;
  if (i >= t.n) return;
  const float4 p = __ldg(t.src + i);
  const float4 q0 = xform_f(a.guess, p.x, p.y, p.z);
  const float4 q = xform_f(a.T, q0.x, q0.y, q0.z);
  NNResult r;
  r.key = kInfKey;
  r.pos = -1;
  if (isfinite(q.x) && isfinite(q.y) && isfinite(q.z))
    r = grid_nn<kSweepThreads>(t.g, q.x, q.y, q.z, a.bound2, a.max_rings, a.use_seed ? t.corr_pos[i] : -1, sc);
  const float d2 = key_d2(r.key);
  const bool keep = (r.key != kInfKey) && ((double)d2 < a.max2);
  const int ti = key_idx(r.key);
  t.corr_idx[i] = keep ? ti : -1;
  t.corr_d2[i] = d2;
  t.corr_pos[i] = r.pos;
  if (!keep) return;
  double C1[9], C2[9], M[9], temp[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    C1[e] = t.cov_src[(size_t)9 * i + e];
    C2[e] = t.cov_tgt[(size_t)9 * ti + e];
  }
  mat3_mul(a.R, C1, M);
  mat3_mul_bt(M, a.R, temp);
#pragma unroll
  for (int e = 0; e < 9; ++e) temp[e] = dadd(temp[e], C2[e]);
  mat3_inv(temp, M);
#pragma unroll
  for (int e = 0; e < 9; ++e) t.mahal[(size_t)e * t.n + i] = M[e];  // nine planes of n: the cost functor reads them coalesced
}

// ---- G3: one evaluation of the cost functor (OptimizationFunctorWithIndices::fdf) ---------------------
struct GicpEvalArgs {
  float Tx[16];    // base_transformation_ with applyState(x) applied (built on the host)
  float base[16];  // base_transformation_
};

struct GicpFdfTask {
  const float4* src;
  const float4* tgt;
  const int* corr_idx;
  const double* mahal;
  double* partials;      // [ceil(n / 256)][kGicpSums]
  double* sums;          // [kGicpSums]: the CTA sums added in CTA order; may point into mapped pinned host memory
  unsigned int* ticket;  // CTAs of this scan that have stored their partials (zero between launches)
  int n;
  int pad;
  GicpEvalArgs a;
};

// The tasks of one launch travel as a kernel parameter (no task upload in front of every evaluation).
constexpr int kGicpParamTasks = 16;
struct GicpFdfBatch {
  GicpFdfTask t[kGicpParamTasks];
};

__global__ void __launch_bounds__(kGicpThreads) gicp_fdf_kernel(const __grid_constant__ GicpFdfBatch batch) {
  __shared__ double smem[kGicpThreads / 32][kGicpSums];
  __shared__ bool s_last;
  const GicpFdfTask& t = batch.t[blockIdx.y];
  const GicpEvalArgs& a = t.a;
  const int n = t.n;
  if (blockIdx.x * kGicpThreads >= n) return;
  const int i = blockIdx.x * kGicpThreads + threadIdx.x;
  double v[kGicpSums];
#pragma unroll
  for (int c = 0; c < kGicpSums; ++c) v[c] = 0.0;
  const int ti = i < n ? t.corr_idx[i] : -1;
  if (ti >= 0) {
    const float4 ps = __ldg(t.src + i);
    const float4 pt = __ldg(t.tgt + ti);
    const float4 pp = xform_f(a.Tx, ps.x, ps.y, ps.z);
    const double r0 = (double)fsub(pp.x, pt.x), r1 = (double)fsub(pp.y, pt.y), r2 = (double)fsub(pp.z, pt.z);
    double M[9];  // plane e of the scan's Mahalanobis matrices at [e * n + i]
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = __ldg(t.mahal + (size_t)e * n + i);
    const double t0 = dadd(dadd(dmul(M[0], r0), dmul(M[1], r1)), dmul(M[2], r2));
    const double t1 = dadd(dadd(dmul(M[3], r0), dmul(M[4], r1)), dmul(M[5], r2));
    const double t2 = dadd(dadd(dmul(M[6], r0), dmul(M[7], r1)), dmul(M[8], r2));
    v[0] = dadd(dadd(dmul(r0, t0), dmul(r1, t1)), dmul(r2, t2));
    v[1] = t0;
    v[2] = t1;
    v[3] = t2;
    const float4 pb = xform_f(a.base, ps.x, ps.y, ps.z);
    const double b0 = (double)pb.x, b1 = (double)pb.y, b2 = (double)pb.z;
    v[4] = dmul(b0, t0); v[5] = dmul(b0, t1); v[6] = dmul(b0, t2);
    v[7] = dmul(b1, t0); v[8] = dmul(b1, t1); v[9] = dmul(b1, t2);
    v[10] = dmul(b2, t0); v[11] = dmul(b2, t1); v[12] = dmul(b2, t2);
    v[13] = 1.0;
  }
  // the fixed tree: xor butterfly 16, 8, 4, 2, 1 inside the warp, warp sums in order, CTA sums in order (last
  // CTA, below).  The butterfly runs as a reduce-scatter: at distance o a lane keeps half of its columns and
  // hands the other half to its partner, so column c ends up complete in one lane after 16 + 8 + 4 + 2 + 1
  // exchanges instead of 14 x 5.  Every sum is the same pair of operands as in the plain butterfly (IEEE addition
  // commutes), so the value of each column is bit-identical to it.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double w16[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) w16[c] = c < kGicpSums ? v[c] : 0.0;
  // distance 16: lanes with bit 4 clear keep columns 0..7, the others 8..15
  double w8[8];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double give = up ? w16[c] : w16[8 + c];
      const double keep = up ? w16[8 + c] : w16[c];
      w8[c] = dadd(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 16));
    }
  }
  double w4[4];
  {
    const bool up = (lane & 8) != 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double give = up ? w8[c] : w8[4 + c];
      const double keep = up ? w8[4 + c] : w8[c];
      w4[c] = dadd(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 8));
    }
  }
  double w2[2];
  {
    const bool up = (lane & 4) != 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double give = up ? w4[c] : w4[2 + c];
      const double keep = up ? w4[2 + c] : w4[c];
      w2[c] = dadd(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 4));
    }
  }
  double w1;
  {
    const bool up = (lane & 2) != 0;
    const double give = up ? w2[0] : w2[1];
    const double keep = up ? w2[1] : w2[0];
    w1 = dadd(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 2));
  }
  w1 = dadd(w1, __shfl_xor_sync(0xFFFFFFFFu, w1, 1));
  // lane l now holds column 8 * bit4 + 4 * bit3 + 2 * bit2 + bit1 of the warp
  {
    const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    if ((lane & 1) == 0 && col < kGicpSums) smem[warp][col] = w1;
  }
  __syncthreads();
  if (threadIdx.x < kGicpSums) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kGicpThreads / 32; ++w) s = dadd(s, smem[w][threadIdx.x]);
    t.partials[(size_t)blockIdx.x * kGicpSums + threadIdx.x] = s;
    __threadfence();
  }
  __syncthreads();
  // The last level of the tree, by whichever CTA of the scan finishes last: the CTA sums added in CTA order
  // (a launch-independent order), 14 doubles per scan written where the host reads them.
  const int nblk = (n + kGicpThreads - 1) / kGicpThreads;
  if (threadIdx.x == 0) s_last = atomicAdd(t.ticket, 1u) == (unsigned int)(nblk - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // all threads fetch the CTA sums (kGicpThreads CTAs per trip, coalesced, every load in flight at once); then
  // thread c adds column c in CTA order out of shared memory: the order of the adds is the definition's, only
  // the loads are no longer one dependent round trip per eight CTAs
  __shared__ double s_part[kGicpThreads][kGicpSums];
  double tot = 0.0;
  for (int b0 = 0; b0 < nblk; b0 += kGicpThreads) {
    const int cnt = min(kGicpThreads, nblk - b0);
    const double* src = t.partials + (size_t)b0 * kGicpSums;
    double* dst = &s_part[0][0];
    for (int w = threadIdx.x; w < cnt * kGicpSums; w += kGicpThreads) dst[w] = __ldcg(src + w);
    __syncthreads();
    if (threadIdx.x < kGicpSums)
      for (int b = 0; b < cnt; ++b) tot = dadd(tot, s_part[b][threadIdx.x]);
    __syncthreads();
  }
  if (threadIdx.x < kGicpSums) t.sums[threadIdx.x] = tot;
  if (threadIdx.x == 0) *t.ticket = 0;
}

}  // namespace b2
