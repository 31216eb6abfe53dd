/*
 * b2icp_oracle.cpp — CPU oracle for the ICP hot path.  TEST INFRASTRUCTURE ONLY (see
 * b2icp_oracle.h: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file).
 *
 * PARITY UNPINNED.  The reference (YoshuaNava/icpslam) delegates every line of hot-path
 * arithmetic to PCL (un-vendored, nominal pin 1.8.1: reference CMakeLists.txt:7,
 * package.xml:13-14) and ships no tests or golden vectors.  Each function below restates the
 * published PCL / FLANN / Eigen algorithm reached from the reference call sites
 *   src/icpslam/icp_odometer.cpp:188-201,205   and   src/icpslam/octree_mapper.cpp:96,104-117
 * following SURVEY.md Appendix A (section numbers quoted per function).
 *
 * Deliberate, documented deviations from PCL (none changes results beyond float rounding):
 *   - k-NN ties resolve to the smallest target index (FLANN: first visited) — SURVEY.md §8c.
 *   - Umeyama sums are accumulated in double (PCL: float Eigen expressions); everything PCL
 *     *stores* as float (points, Matrix4f transformation_/final_) is stored as float here.
 *   - The 3x3 SVD is a one-sided Jacobi (Eigen: two-sided JacobiSVD); both are exact to ~1e-15.
 *
 * Build: see oracle/Makefile (-O3 -ffp-contract=off so that x*y+z is never fused: baseline
 * x86-64 builds of PCL have no FMA and the float op order is part of the restatement).
 */
#include "b2icp_oracle.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_set>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using clk = std::chrono::steady_clock;
inline double ms_since(clk::time_point t0) {
  return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

int g_threads = 1;

/* FLANN L2_Simple<float> (Appendix A.6): result += diff*diff for x, y, z in that order. */
inline float sqdist3(const float* a, const float* b) {
  float r = 0.0f;
  float d0 = a[0] - b[0];
  r += d0 * d0;
  float d1 = a[1] - b[1];
  r += d1 * d1;
  float d2 = a[2] - b[2];
  r += d2 * d2;
  return r;
}

/* (d2, idx) lexicographic "better than": canonical tie rule = smallest index. */
inline bool better(float d2, int32_t idx, float bd2, int32_t bidx) {
  return d2 < bd2 || (d2 == bd2 && idx < bidx);
}

/* ------------------------------------------------------------------------------------------ */
/* pcl::KdTreeFLANN / flann::KDTreeSingleIndex (Appendix A.6): leaf size 15, reorder=true,
 * exact search (eps = 0), incremental per-dimension lower bounds.  The lower bound is summed in
 * the same association order as sqdist3 so that, by monotonicity of IEEE rounding, it can never
 * exceed the float distance of any point in the subtree: pruning is exact, not approximate. */
struct KdNode {
  int32_t left, right; /* children; left < 0 => leaf */
  int32_t begin, end;  /* leaf range in reordered arrays */
  int32_t dim;
  float divlow, divhigh;
};

struct KdTree {
  std::vector<float> pts; /* reordered xyz, 3 floats per point */
  std::vector<int32_t> ids; /* original index of reordered point */
  std::vector<KdNode> nodes;
  float bbmin[3], bbmax[3];
  size_t n = 0;
  static constexpr int kLeaf = 15;

  int32_t build_rec(const float* xyzw, std::vector<int32_t>& ind, int32_t lo, int32_t hi, float* bmin,
                    float* bmax) {
    int32_t me = (int32_t)nodes.size();
    nodes.push_back(KdNode{-1, -1, lo, hi, 0, 0.f, 0.f});
    if (hi - lo <= kLeaf) {
      /* leaf: tighten the box to its points (FLANN does the same) */
      for (int d = 0; d < 3; ++d) {
        bmin[d] = std::numeric_limits<float>::max();
        bmax[d] = -std::numeric_limits<float>::max();
      }
      for (int32_t i = lo; i < hi; ++i)
        for (int d = 0; d < 3; ++d) {
          float v = xyzw[4 * (size_t)ind[i] + d];
          bmin[d] = std::min(bmin[d], v);
          bmax[d] = std::max(bmax[d], v);
        }
      return me;
    }
    /* split on the widest dimension of the actual extent, at the median element */
    float emin[3], emax[3];
    for (int d = 0; d < 3; ++d) {
      emin[d] = std::numeric_limits<float>::max();
      emax[d] = -std::numeric_limits<float>::max();
    }
    for (int32_t i = lo; i < hi; ++i)
      for (int d = 0; d < 3; ++d) {
        float v = xyzw[4 * (size_t)ind[i] + d];
        emin[d] = std::min(emin[d], v);
        emax[d] = std::max(emax[d], v);
      }
    int dim = 0;
    float span = emax[0] - emin[0];
    for (int d = 1; d < 3; ++d)
      if (emax[d] - emin[d] > span) {
        span = emax[d] - emin[d];
        dim = d;
      }
    int32_t mid = lo + (hi - lo) / 2;
    std::nth_element(ind.begin() + lo, ind.begin() + mid, ind.begin() + hi, [&](int32_t a, int32_t b) {
      float va = xyzw[4 * (size_t)a + dim], vb = xyzw[4 * (size_t)b + dim];
      return va < vb || (va == vb && a < b);
    });
    float lmin[3], lmax[3], rmin[3], rmax[3];
    int32_t l = build_rec(xyzw, ind, lo, mid, lmin, lmax);
    int32_t r = build_rec(xyzw, ind, mid, hi, rmin, rmax);
    nodes[me].left = l;
    nodes[me].right = r;
    nodes[me].dim = dim;
    nodes[me].divlow = lmax[dim];
    nodes[me].divhigh = rmin[dim];
    for (int d = 0; d < 3; ++d) {
      bmin[d] = std::min(lmin[d], rmin[d]);
      bmax[d] = std::max(lmax[d], rmax[d]);
    }
    return me;
  }

  void build(const float* xyzw, size_t n_) {
    n = n_;
    std::vector<int32_t> ind(n);
    std::iota(ind.begin(), ind.end(), 0);
    nodes.clear();
    nodes.reserve(2 * n / kLeaf + 16);
    if (n) build_rec(xyzw, ind, 0, (int32_t)n, bbmin, bbmax);
    pts.resize(3 * n);
    ids = ind;
    for (size_t i = 0; i < n; ++i)
      for (int d = 0; d < 3; ++d) pts[3 * i + d] = xyzw[4 * (size_t)ind[i] + d];
  }

  struct Result {
    int k;
    int count;
    float* d2;
    int32_t* idx;
    inline float worst() const { return count < k ? std::numeric_limits<float>::infinity() : d2[k - 1]; }
    inline void offer(float d, int32_t i) {
      if (count == k && !better(d, i, d2[k - 1], idx[k - 1])) return;
      int pos = count < k ? count++ : k - 1;
      while (pos > 0 && better(d, i, d2[pos - 1], idx[pos - 1])) {
        d2[pos] = d2[pos - 1];
        idx[pos] = idx[pos - 1];
        --pos;
      }
      d2[pos] = d;
      idx[pos] = i;
    }
  };

  void search_rec(int32_t ni, const float* q, float* dists, Result& res) const {
    const KdNode& nd = nodes[ni];
    if (nd.left < 0) {
      for (int32_t i = nd.begin; i < nd.end; ++i) {
        float d = sqdist3(q, &pts[3 * (size_t)i]);
        res.offer(d, ids[i]);
      }
      return;
    }
    int dim = nd.dim;
    float val = q[dim];
    float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int32_t best, other;
    float cut;
    if (diff1 + diff2 < 0) {
      best = nd.left;
      other = nd.right;
      cut = diff2 * diff2; /* distance to the low face of the right child */
    } else {
      best = nd.right;
      other = nd.left;
      cut = diff1 * diff1;
    }
    search_rec(best, q, dists, res);
    float saved = dists[dim];
    /* the far child lies beyond the plane only if q is outside its slab on this axis */
    bool outside = (other == nd.right) ? (val < nd.divhigh) : (val > nd.divlow);
    float nd_d = outside ? std::max(cut, saved) : saved;
    dists[dim] = nd_d;
    float lb = 0.0f;
    lb += dists[0];
    lb += dists[1];
    lb += dists[2];
    if (lb <= res.worst()) search_rec(other, q, dists, res);
    dists[dim] = saved;
  }

  int knn(const float* q, int k, int32_t* idx, float* d2) const {
    Result res{k, 0, d2, idx};
    if (!n) return 0;
    float dists[3];
    for (int d = 0; d < 3; ++d) {
      float v = 0.f;
      if (q[d] < bbmin[d]) v = (q[d] - bbmin[d]) * (q[d] - bbmin[d]);
      if (q[d] > bbmax[d]) v = (q[d] - bbmax[d]) * (q[d] - bbmax[d]);
      dists[d] = v;
    }
    search_rec(0, q, dists, res);
    return res.count;
  }
};

/* ------------------------------------------------------------------------------------------ */
/* Appendix A.5 — float version: ((m0*x + m1*y) + m2*z) + m3, each op rounded to float. */
inline void xform_f(const float* T, const float* p, float* o) {
  float x = p[0], y = p[1], z = p[2];
  o[0] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
  o[1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
  o[2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
  o[3] = 1.0f;
}
inline void xform_d(const double* T, const float* p, float* o) {
  double x = p[0], y = p[1], z = p[2];
  o[0] = (float)(((T[0] * x + T[1] * y) + T[2] * z) + T[3]);
  o[1] = (float)(((T[4] * x + T[5] * y) + T[6] * z) + T[7]);
  o[2] = (float)(((T[8] * x + T[9] * y) + T[10] * z) + T[11]);
  o[3] = 1.0f;
}

inline void mat4f_mul(const float* A, const float* B, float* C) {
  float t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
      t[4 * r + c] = s;
    }
  std::memcpy(C, t, sizeof(t));
}

/* ------------------------------------------------------------------------------------------ */
/* 3x3 SVD A = U diag(s) V^T by one-sided (Hestenes) Jacobi, s sorted descending, U and V
 * completed to full orthogonal matrices (JacobiSVD ComputeFullU|ComputeFullV semantics). */
inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double det3(const double* M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
         M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* `skip`: a pair whose |cos| is at or below it is left alone.  1e-17 (below one ulp: the sweeps run until rounding
 * happens to produce an exact zero, or to the cap) for the Umeyama solve; 1e-15 for the GICP covariances, which is
 * where a rotation stops changing an fp64 column and still tighter than Eigen::JacobiSVD's own 2 eps. */
void svd3_skip(const double* A, double* U, double* s, double* V, double skip) {
  double B[9], W[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(B, A, sizeof(B));
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double al = 0, be = 0, ga = 0;
        for (int i = 0; i < 3; ++i) {
          al += B[3 * i + p] * B[3 * i + p];
          be += B[3 * i + q] * B[3 * i + q];
          ga += B[3 * i + p] * B[3 * i + q];
        }
        if (ga == 0.0 || std::fabs(ga) <= skip * std::sqrt(al * be)) continue;
        rotated = true;
        double zeta = (be - al) / (2.0 * ga);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
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
  for (int j = 0; j < 3; ++j)
    nrm[j] = std::sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  std::sort(ord, ord + 3, [&](int a, int b) { return nrm[a] > nrm[b] || (nrm[a] == nrm[b] && a < b); });
  double u[3][3], v[3][3];
  for (int j = 0; j < 3; ++j) {
    int c = ord[j];
    s[j] = nrm[c];
    for (int i = 0; i < 3; ++i) v[j][i] = W[3 * i + c];
  }
  double tiny = 1e-300 + s[0] * 1e-14;
  int rank = 0;
  for (int j = 0; j < 3; ++j)
    if (s[j] > tiny) {
      int c = ord[j];
      for (int i = 0; i < 3; ++i) u[j][i] = B[3 * i + c] / s[j];
      rank = j + 1;
    }
  if (rank == 0) {
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) u[j][i] = (i == j);
  } else if (rank == 1) {
    /* any orthonormal completion */
    double e[3] = {0, 0, 0};
    int m = 0;
    for (int i = 1; i < 3; ++i)
      if (std::fabs(u[0][i]) < std::fabs(u[0][m])) m = i;
    e[m] = 1.0;
    cross3(u[0], e, u[1]);
    double n1 = std::sqrt(u[1][0] * u[1][0] + u[1][1] * u[1][1] + u[1][2] * u[1][2]);
    for (int i = 0; i < 3; ++i) u[1][i] /= n1;
    cross3(u[0], u[1], u[2]);
  } else if (rank == 2) {
    cross3(u[0], u[1], u[2]);
    double n2 = std::sqrt(u[2][0] * u[2][0] + u[2][1] * u[2][1] + u[2][2] * u[2][2]);
    for (int i = 0; i < 3; ++i) u[2][i] /= n2;
  }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) {
      U[3 * i + j] = u[j][i];
      V[3 * i + j] = v[j][i];
    }
}
void svd3(const double* A, double* U, double* s, double* V) { svd3_skip(A, U, s, V, 1e-17); }

/* Eigen::umeyama(src, dst, with_scaling=false) as called by
 * pcl::registration::TransformationEstimationSVD (Appendix A.3) from the 16 running sums
 * n, sum(src), sum(dst), sum(dst * src^T). */
struct UmeyamaSums {
  double n = 0, s[3] = {0, 0, 0}, d[3] = {0, 0, 0}, ds[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  inline void add(const float* src, const float* dst) {
    n += 1.0;
    for (int i = 0; i < 3; ++i) {
      s[i] += (double)src[i];
      d[i] += (double)dst[i];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) ds[3 * r + c] += (double)dst[r] * (double)src[c];
  }
};

void umeyama_from_sums(const UmeyamaSums& S, double* T16) {
  double sm[3], dm[3], sigma[9];
  for (int i = 0; i < 3; ++i) {
    sm[i] = S.s[i] / S.n;
    dm[i] = S.d[i] / S.n;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) sigma[3 * r + c] = S.ds[3 * r + c] / S.n - dm[r] * sm[c];
  double U[9], sv[3], V[9];
  svd3(sigma, U, sv, V);
  double sg[3] = {1, 1, 1};
  if (det3(U) * det3(V) < 0) sg[2] = -1; /* Eigen >= 3.3 reflection fix */
  double R[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double a = 0;
      for (int k = 0; k < 3; ++k) a += U[3 * r + k] * sg[k] * V[3 * c + k];
      R[3 * r + c] = a;
    }
  for (int i = 0; i < 16; ++i) T16[i] = 0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T16[4 * r + c] = R[3 * r + c];
    T16[4 * r + 3] = dm[r] - (R[3 * r] * sm[0] + R[3 * r + 1] * sm[1] + R[3 * r + 2] * sm[2]);
  }
  T16[15] = 1.0;
}

void nn_sweep(const KdTree& tree, const float* q, size_t n, int32_t* idx, float* d2) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
  for (long i = 0; i < (long)n; ++i) {
    int32_t bi = -1;
    float bd = std::numeric_limits<float>::infinity();
    tree.knn(q + 4 * i, 1, &bi, &bd);
    idx[i] = bi;
    d2[i] = bd;
  }
}

void set_identity_result(b2icp_result* out) {
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < 4; ++i) out->T[5 * i] = 1.0;
  out->fitness = std::numeric_limits<double>::quiet_NaN();
  out->mse_last = std::numeric_limits<double>::quiet_NaN();
}

bool all_finite(const float* xyzw, size_t n) {
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d)
      if (!std::isfinite(xyzw[4 * i + d])) return false;
  return true;
}

/* pcl::IterativeClosestPoint::computeTransformation + DefaultConvergenceCriteria (App. A.3). */
int align_p2p(const b2icp_params* p, const float* src, size_t ns, const float* tgt, size_t nt,
              const float* guess16, b2icp_result* out, float* aligned, int record_iter, int32_t* corr_idx,
              float* corr_d2, b2o_stage_ms* st) {
  auto t0 = clk::now();
  KdTree tree;
  tree.build(tgt, nt); /* Registration::initCompute -> tree_->setInputCloud(target_) */
  st->build += ms_since(t0);

  float final_T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  bool guess_is_identity = true;
  if (guess16) {
    for (int i = 0; i < 16; ++i)
      if (guess16[i] != final_T[i]) guess_is_identity = false;
    std::memcpy(final_T, guess16, sizeof(final_T));
  }
  std::vector<float> cur(4 * ns);
  auto t1 = clk::now();
  if (guess_is_identity) {
    for (size_t i = 0; i < ns; ++i) {
      cur[4 * i] = src[4 * i];
      cur[4 * i + 1] = src[4 * i + 1];
      cur[4 * i + 2] = src[4 * i + 2];
      cur[4 * i + 3] = 1.0f;
    }
  } else {
    for (size_t i = 0; i < ns; ++i) xform_f(final_T, src + 4 * i, &cur[4 * i]);
  }
  st->transform += ms_since(t1);

  std::vector<int32_t> idx(ns);
  std::vector<float> d2(ns);
  const double max2 = p->max_correspondence_distance * p->max_correspondence_distance;
  const double rot_thresh = 1.0 - p->transformation_epsilon;
  const double trans_thresh = p->transformation_epsilon;
  const double mse_abs = 1e-12, mse_rel = p->euclidean_fitness_epsilon;
  double prev_mse = std::numeric_limits<double>::max();
  int iters = 0, converged = 0, status = B2ICP_OK, n_corr = 0;
  double mse = std::numeric_limits<double>::quiet_NaN();
  float Tinc[16];

  do {
    auto tn = clk::now();
    nn_sweep(tree, cur.data(), ns, idx.data(), d2.data());
    st->nn += ms_since(tn);

    auto ts = clk::now();
    /* CorrespondenceEstimation::determineCorrespondences: `if (distance > max_dist_sqr) continue;` */
    UmeyamaSums S;
    double dsum = 0;
    n_corr = 0;
    for (size_t i = 0; i < ns; ++i) {
      bool keep = idx[i] >= 0 && !((double)d2[i] > max2);
      if (keep) {
        S.add(&cur[4 * i], tgt + 4 * (size_t)idx[i]);
        dsum += (double)d2[i];
        ++n_corr;
      }
      if ((record_iter == iters || record_iter < 0) && corr_idx) corr_idx[i] = keep ? idx[i] : -1;
      if ((record_iter == iters || record_iter < 0) && corr_d2) corr_d2[i] = d2[i];
    }
    if (n_corr < 3) { /* min_number_correspondences_ = 3 */
      status = B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES;
      converged = 0;
      st->solve += ms_since(ts);
      break;
    }
    double Td[16];
    umeyama_from_sums(S, Td);
    for (int i = 0; i < 16; ++i) Tinc[i] = (float)Td[i]; /* transformation_ is a Matrix4f */
    st->solve += ms_since(ts);

    auto tt = clk::now();
    for (size_t i = 0; i < ns; ++i) { /* transformCloud(input_transformed, in place) */
      float o[4];
      xform_f(Tinc, &cur[4 * i], o);
      std::memcpy(&cur[4 * i], o, sizeof(o));
    }
    mat4f_mul(Tinc, final_T, final_T); /* final = transformation_ * final */
    st->transform += ms_since(tt);
    ++iters;

    /* DefaultConvergenceCriteria::hasConverged (PCL 1.8.x) */
    mse = dsum / (double)n_corr;
    if (iters >= p->max_iterations) {
      converged = 1;
    } else {
      double cos_angle = 0.5 * (double)(Tinc[0] + Tinc[5] + Tinc[10] - 1.0f);
      double tr2 = (double)(Tinc[3] * Tinc[3] + Tinc[7] * Tinc[7] + Tinc[11] * Tinc[11]);
      if (cos_angle >= rot_thresh && tr2 <= trans_thresh) {
        converged = 1;
      } else if (std::fabs(mse - prev_mse) < mse_abs) {
        converged = 1;
      } else if (std::fabs(mse - prev_mse) / prev_mse < mse_rel) {
        converged = 1;
      } else {
        prev_mse = mse;
      }
    }
  } while (!converged);

  for (int i = 0; i < 16; ++i) out->T[i] = (double)final_T[i];
  out->converged = converged;
  out->iterations = iters;
  out->n_corr_last = n_corr;
  out->status_detail = status;
  out->mse_last = mse;
  if (aligned) {
    auto tt = clk::now();
    for (size_t i = 0; i < ns; ++i) xform_f(final_T, src + 4 * i, aligned + 4 * i);
    st->transform += ms_since(tt);
  }
  return status;
}

}  // namespace

/* GICP lives in gicp_oracle.cpp */
int b2o_align_gicp_impl(const b2icp_params* p, const float* src, size_t ns, const float* tgt, size_t nt,
                        const float* guess16, b2icp_result* out, float* aligned, int record_iter,
                        int32_t* corr_idx, float* corr_d2, b2o_stage_ms* st, int threads);

extern "C" {

int b2o_set_threads(int n) {
#ifdef _OPENMP
  /* the cores this process may run on, not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers,
   * which would silently turn the "all host threads" baseline into a single-thread one) */
  int mx = omp_get_num_procs();
  if (n < 1) n = 1;
  if (n > mx) n = mx;
  g_threads = n;
  omp_set_dynamic(0);
#else
  (void)n;
  g_threads = 1;
#endif
  return g_threads;
}

int b2o_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

int b2o_nn_brute(const float* tgt, size_t nt, const float* q, size_t nq, int32_t* idx, float* d2) {
  if (!tgt || !q || !idx) return B2ICP_ERR_INVALID_ARG;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (long i = 0; i < (long)nq; ++i) {
    int32_t bi = -1;
    float bd = std::numeric_limits<float>::infinity();
    for (size_t j = 0; j < nt; ++j) {
      float d = sqdist3(q + 4 * i, tgt + 4 * j);
      if (d < bd) { /* ascending j + strict < == smallest index among ties */
        bd = d;
        bi = (int32_t)j;
      }
    }
    idx[i] = bi;
    if (d2) d2[i] = bd;
  }
  return B2ICP_OK;
}

void* b2o_kdtree_build(const float* tgt, size_t nt) {
  KdTree* t = new KdTree();
  t->build(tgt, nt);
  return t;
}
void b2o_kdtree_free(void* tree) { delete static_cast<KdTree*>(tree); }

int b2o_kdtree_nn(const void* tree, const float* q, size_t nq, int32_t* idx, float* d2) {
  if (!tree || !q || !idx) return B2ICP_ERR_INVALID_ARG;
  std::vector<float> tmp;
  if (!d2) {
    tmp.resize(nq);
    d2 = tmp.data();
  }
  nn_sweep(*static_cast<const KdTree*>(tree), q, nq, idx, d2);
  return B2ICP_OK;
}

int b2o_kdtree_knn(const void* tree, const float* q, size_t nq, int k, int32_t* idx, float* d2) {
  if (!tree || !q || !idx || !d2 || k < 1) return B2ICP_ERR_INVALID_ARG;
  const KdTree& t = *static_cast<const KdTree*>(tree);
  if ((size_t)k > t.n) return B2ICP_ERR_TOO_FEW_POINTS;
#pragma omp parallel for schedule(dynamic, 256) num_threads(g_threads)
  for (long i = 0; i < (long)nq; ++i) t.knn(q + 4 * i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
  return B2ICP_OK;
}

int b2o_transform_cloud_d(const float* in, size_t n, const double* T, float* out) {
  if (!in || !T || !out) return B2ICP_ERR_INVALID_ARG;
  for (size_t i = 0; i < n; ++i) {
    float o[4];
    xform_d(T, in + 4 * i, o);
    std::memcpy(out + 4 * i, o, sizeof(o));
  }
  return B2ICP_OK;
}

int b2o_transform_cloud_f(const float* in, size_t n, const float* T, float* out) {
  if (!in || !T || !out) return B2ICP_ERR_INVALID_ARG;
  for (size_t i = 0; i < n; ++i) {
    float o[4];
    xform_f(T, in + 4 * i, o);
    std::memcpy(out + 4 * i, o, sizeof(o));
  }
  return B2ICP_OK;
}

int b2o_umeyama(const float* src, const float* dst, size_t n, double* T16) {
  if (!src || !dst || !T16 || n < 3) return B2ICP_ERR_INVALID_ARG;
  UmeyamaSums S;
  for (size_t i = 0; i < n; ++i) S.add(src + 4 * i, dst + 4 * i);
  umeyama_from_sums(S, T16);
  return B2ICP_OK;
}

int b2o_svd3(const double* A9, double* U9, double* s3, double* V9) {
  svd3(A9, U9, s3, V9);
  return B2ICP_OK;
}

int b2o_svd3_cov(const double* A9, double* U9, double* s3, double* V9) {
  svd3_skip(A9, U9, s3, V9, 1e-15);
  return B2ICP_OK;
}

int b2o_align(const b2icp_params* p, const float* src, size_t ns, const float* tgt, size_t nt,
              const float* guess16, b2icp_result* out, float* aligned, int record_iter, int32_t* corr_idx,
              float* corr_d2, b2o_stage_ms* stages) {
  if (!p || !out) return B2ICP_ERR_INVALID_ARG;
  set_identity_result(out);
  b2o_stage_ms local;
  std::memset(&local, 0, sizeof(local));
  b2o_stage_ms* st = stages ? stages : &local;
  std::memset(st, 0, sizeof(*st));
  if (!src || ns == 0) return B2ICP_ERR_EMPTY_CLOUD; /* GICP setInputSource / initCompute reject */
  if (!tgt || nt == 0) return B2ICP_ERR_NO_TARGET;
  if (!all_finite(src, ns) || !all_finite(tgt, nt)) return B2ICP_ERR_NONFINITE_INPUT;
  auto t0 = clk::now();
  int rc;
  if (p->mode == B2ICP_MODE_GICP_BFGS)
    rc = b2o_align_gicp_impl(p, src, ns, tgt, nt, guess16, out, aligned, record_iter, corr_idx, corr_d2, st,
                             g_threads);
  else
    rc = align_p2p(p, src, ns, tgt, nt, guess16, out, aligned, record_iter, corr_idx, corr_d2, st);
  st->total = ms_since(t0);
  return rc;
}

/* Registration::getFitnessScore(max_range): transform the source by the float final
 * transformation, exact 1-NN each, mean of d2 over the pairs with d2 <= max_range;
 * DBL_MAX when no pair qualifies. */
int b2o_fitness(const float* src, size_t ns, const float* tgt, size_t nt, const float* T16, double max_range,
                double* out) {
  if (!src || !tgt || !T16 || !out) return B2ICP_ERR_INVALID_ARG;
  KdTree tree;
  tree.build(tgt, nt);
  std::vector<float> cur(4 * ns), d2(ns);
  std::vector<int32_t> idx(ns);
  for (size_t i = 0; i < ns; ++i) xform_f(T16, src + 4 * i, &cur[4 * i]);
  nn_sweep(tree, cur.data(), ns, idx.data(), d2.data());
  double sum = 0;
  long nr = 0;
  for (size_t i = 0; i < ns; ++i)
    if (idx[i] >= 0 && (double)d2[i] <= max_range) {
      sum += (double)d2[i];
      ++nr;
    }
  *out = nr > 0 ? sum / (double)nr : std::numeric_limits<double>::max();
  return B2ICP_OK;
}

/* OctreeMapper::addPointsToMap (reference src/icpslam/octree_mapper.cpp:63-71): for each
 * point in order, `if (!isVoxelOccupiedAtPoint(p)) addPointToCloud(p)`.  Voxels are the cells of the global
 * lattice floor(p / resolution) (PCL's lattice is anchored on the octree's bounding box, which depends on the
 * insertion history: documented deviation).  Non-finite points are skipped. */
int b2o_map_insert(const float* map, size_t n_map, const float* in, size_t n, double resolution, float* out_added,
                   size_t* n_added) {
  if (!n_added) return B2ICP_ERR_INVALID_ARG;
  *n_added = 0;
  if (!(resolution > 0)) return B2ICP_ERR_INVALID_ARG;
  if (n == 0) return B2ICP_OK;
  if (!in || !out_added || (n_map && !map)) return B2ICP_ERR_INVALID_ARG;
  const double inv = 1.0 / resolution;
  auto key_of = [&](const float* p, uint64_t& key) {
    if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) return false;
    const int64_t ix = (int64_t)std::floor((double)p[0] * inv), iy = (int64_t)std::floor((double)p[1] * inv),
                  iz = (int64_t)std::floor((double)p[2] * inv);
    key = ((uint64_t)(ix & 0x1FFFFF) << 42) | ((uint64_t)(iy & 0x1FFFFF) << 21) | (uint64_t)(iz & 0x1FFFFF);
    return true;
  };
  std::unordered_set<uint64_t> occupied;
  occupied.reserve((n_map + n) * 2);
  uint64_t key;
  for (size_t i = 0; i < n_map; ++i)
    if (key_of(map + 4 * i, key)) occupied.insert(key);
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!key_of(in + 4 * i, key)) continue;
    if (occupied.insert(key).second) {
      out_added[4 * m] = in[4 * i];
      out_added[4 * m + 1] = in[4 * i + 1];
      out_added[4 * m + 2] = in[4 * i + 2];
      out_added[4 * m + 3] = 1.0f;
      ++m;
    }
  }
  *n_added = m;
  return B2ICP_OK;
}

/* pcl::VoxelGrid<PointXYZ>::applyFilter (Appendix A.8) as called by IcpOdometer::voxelFilterCloud
 * (reference src/icpslam/icp_odometer.cpp:96-101) with a cubic leaf.  PCL's std::sort is unstable, so the
 * order of the float additions inside a leaf is unspecified there; here it is input order (stable sort).
 * out must hold n points; returns the number written in *n_out. */
int b2o_voxel_filter(const float* in, size_t n, float leaf, float* out, size_t* n_out) {
  if (!n_out) return B2ICP_ERR_INVALID_ARG;
  *n_out = 0;
  if (n == 0) return B2ICP_OK;
  if (!in || !out || !(leaf > 0)) return B2ICP_ERR_INVALID_ARG;
  if (!all_finite(in, n)) return B2ICP_ERR_NONFINITE_INPUT;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], in[4 * i + d]);
      mx[d] = std::max(mx[d], in[4 * i + d]);
    }
  int min_b[3], div[3];
  long long cells = 1;
  for (int d = 0; d < 3; ++d) {
    min_b[d] = (int)std::floor(mn[d] * inv);
    div[d] = (int)std::floor(mx[d] * inv) - min_b[d] + 1;
    cells *= div[d];
    if (cells > (long long)INT32_MAX) { /* "Leaf size is too small": output = input */
      std::memcpy(out, in, n * 4 * sizeof(float));
      *n_out = n;
      return B2ICP_OK;
    }
  }
  std::vector<std::pair<int, uint32_t>> keys(n);
  for (size_t i = 0; i < n; ++i) {
    int i0 = (int)(std::floor(in[4 * i] * inv) - (float)min_b[0]);
    int i1 = (int)(std::floor(in[4 * i + 1] * inv) - (float)min_b[1]);
    int i2 = (int)(std::floor(in[4 * i + 2] * inv) - (float)min_b[2]);
    keys[i] = {i0 + i1 * div[0] + i2 * div[0] * div[1], (uint32_t)i};
  }
  std::stable_sort(keys.begin(), keys.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t w = 0;
  for (size_t a = 0; a < n;) {
    size_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    while (b < n && keys[b].first == keys[a].first) {
      const float* p = in + 4 * (size_t)keys[b].second;
      sx += p[0];
      sy += p[1];
      sz += p[2];
      ++b;
    }
    const float cnt = (float)(b - a);
    out[4 * w] = sx / cnt;
    out[4 * w + 1] = sy / cnt;
    out[4 * w + 2] = sz / cnt;
    out[4 * w + 3] = 1.0f;
    ++w;
    a = b;
  }
  *n_out = w;
  return B2ICP_OK;
}

/* Pose6DOF algebra, reference src/utils/pose6DOF.cpp. pose7 = {px,py,pz,qw,qx,qy,qz}. */
static void quat_mul(const double* a, const double* b, double* o) { /* w,x,y,z */
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w;
  o[1] = x;
  o[2] = y;
  o[3] = z;
}
static void quat_rot(const double* q, const double* v, double* o) {
  /* v' = v + 2 w (u x v) + 2 u x (u x v) */
  const double* u = q + 1;
  double uv[3], uuv[3];
  cross3(u, v, uv);
  cross3(u, uv, uuv);
  for (int i = 0; i < 3; ++i) o[i] = v[i] + 2.0 * (q[0] * uv[i] + uuv[i]);
}
static void quat_normalize(double* q) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n > 0)
    for (int i = 0; i < 4; ++i) q[i] /= n;
}

/* pose6DOF.cpp:98-105: pos = p1.pos + R(p1) * p2.pos; rot = q1 * q2; normalize */
void b2o_pose_compose(const double* a, const double* b, double* o) {
  double r[3], q[4];
  quat_rot(a + 3, b, r);
  quat_mul(a + 3, b + 3, q);
  quat_normalize(q);
  for (int i = 0; i < 3; ++i) o[i] = a[i] + r[i];
  for (int i = 0; i < 4; ++i) o[3 + i] = q[i];
}

/* pose6DOF.cpp:117-122: pos = -(q^-1 * pos); rot = q^-1 */
void b2o_pose_inverse(const double* a, double* o) {
  double qi[4] = {a[3], -a[4], -a[5], -a[6]};
  double n2 = a[3] * a[3] + a[4] * a[4] + a[5] * a[5] + a[6] * a[6];
  for (int i = 0; i < 4; ++i) qi[i] /= n2;
  double r[3];
  quat_rot(qi, a, r);
  for (int i = 0; i < 3; ++i) o[i] = -r[i];
  for (int i = 0; i < 4; ++i) o[3 + i] = qi[i];
}

/* pose6DOF.cpp:8-13,185-190: pos = T[0:3,3]; rot = Quaterniond(T[0:3,0:3]); normalize.
 * Eigen's rotation-matrix -> quaternion conversion (Shepperd's branches). */
void b2o_pose_from_matrix(const double* T, double* o) {
  o[0] = T[3];
  o[1] = T[7];
  o[2] = T[11];
  double m[3][3] = {{T[0], T[1], T[2]}, {T[4], T[5], T[6]}, {T[8], T[9], T[10]}};
  double q[4];
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[2][1] - m[1][2]) * t;
    q[2] = (m[0][2] - m[2][0]) * t;
    q[3] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[k][j] - m[j][k]) * t;
    q[1 + j] = (m[j][i] + m[i][j]) * t;
    q[1 + k] = (m[k][i] + m[i][k]) * t;
  }
  quat_normalize(q);
  for (int i = 0; i < 4; ++i) o[3 + i] = q[i];
}

} /* extern "C" */
