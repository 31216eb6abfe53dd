/*
 * gicp_oracle.cpp — CPU oracle of pcl::GeneralizedIterativeClosestPoint, the class the reference
 * actually instantiates (reference src/icpslam/icp_odometer.cpp:188, src/icpslam/octree_mapper.cpp:104).
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see b2icp_oracle.h).
 *
 * Restates, following SURVEY.md Appendix A.2 / A.4:
 *   computeCovariances            k = 20 neighbourhood covariance, SVD, eigenvalues -> (1, 1, gicp_epsilon)
 *   computeTransformation         outer loop: float query transform, 1-NN, strict d2 < max^2 gate,
 *                                 Mahalanobis M = (R C1 R^T + C2)^-1, BFGS, delta test
 *   estimateRigidTransformationBFGS + OptimizationFunctorWithIndices (f, df, fdf), applyState,
 *                                 computeRDerivative
 *   BFGS<Functor>                 PCL's port of GSL vector_bfgs2 + linear_minimize.c (Fletcher line search)
 * PCL's BFGS default step_size is taken as 1 (bfgs.h Parameters(); the surveyor flagged it [verify]).
 */
#include "b2icp_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace {

using clk = std::chrono::steady_clock;
inline double ms_since(clk::time_point t0) {
  return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

constexpr double kStepSize = 1.0;       /* BFGS Parameters::step_size */
constexpr double kGradientTol = 1e-2;   /* gicp.hpp: const double gradient_tol = 1e-2 */
constexpr double kDblEps = 2.220446049250313e-16;

enum Status { kRunning = -1, kSuccess = 0, kNoProgress = 1 };

/* ---- small linear algebra ---------------------------------------------------------------------- */
inline void mat3_mul(const double* A, const double* B, double* C) {
  double t[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  std::memcpy(C, t, sizeof(t));
}
inline void mat3_mul_bt(const double* A, const double* B, double* C) { /* A * B^T */
  double t[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
  std::memcpy(C, t, sizeof(t));
}
/* Eigen's closed-form 3x3 inverse: cofactors / determinant */
inline void mat3_inv(const double* M, double* I) {
  double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
  double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
  double id = 1.0 / det;
  I[0] = c00 * id;
  I[1] = (M[2] * M[7] - M[1] * M[8]) * id;
  I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
  I[3] = c01 * id;
  I[4] = (M[0] * M[8] - M[2] * M[6]) * id;
  I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
  I[6] = c02 * id;
  I[7] = (M[1] * M[6] - M[0] * M[7]) * id;
  I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

/* Eigen float 4x4 * (x,y,z,1): ((m0 x + m1 y) + m2 z) + m3 */
inline void xform_f(const float* T, const float* p, float* o) {
  float x = p[0], y = p[1], z = p[2];
  o[0] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
  o[1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
  o[2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
}

/* GICP::applyState: R = AngleAxisf(x5,Z) * AngleAxisf(x4,Y) * AngleAxisf(x3,X) (float quaternions),
 * t.topLeft3x3 = R * t.topLeft3x3; t.col(3) += (x0,x1,x2,0). */
void apply_state(float* t, const double* x) {
  float hz = 0.5f * (float)x[5], hy = 0.5f * (float)x[4], hx = 0.5f * (float)x[3];
  /* quaternions (w, x, y, z) */
  float qz[4] = {std::cos(hz), 0.f, 0.f, std::sin(hz)};
  float qy[4] = {std::cos(hy), 0.f, std::sin(hy), 0.f};
  float qx[4] = {std::cos(hx), std::sin(hx), 0.f, 0.f};
  auto qmul = [](const float* a, const float* b, float* o) {
    float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    float xx = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    float yy = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    float zz = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    o[0] = w;
    o[1] = xx;
    o[2] = yy;
    o[3] = zz;
  };
  float qzy[4], q[4];
  qmul(qz, qy, qzy);
  qmul(qzy, qx, q);
  /* Eigen QuaternionBase::toRotationMatrix */
  const float tx = 2.f * q[1], ty = 2.f * q[2], tz = 2.f * q[3];
  const float twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
  const float txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
  const float tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
  float R[9] = {1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx,
                txz - twy,         tyz + twx, 1.f - (txx + tyy)};
  float n[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s += R[3 * r + k] * t[4 * k + c];
      n[3 * r + c] = s;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) t[4 * r + c] = n[3 * r + c];
  t[3] += (float)x[0];
  t[7] += (float)x[1];
  t[11] += (float)x[2];
}

/* GICP::computeRDerivative: g[3..5] = tr(dR/dphi,theta,psi * R) with tr(A*B) = sum_ij A(j,i) B(i,j). */
void compute_r_derivative(const double* x, const double* R, double* g) {
  double phi = x[3], theta = x[4], psi = x[5];
  double cphi = std::cos(phi), sphi = std::sin(phi), ctheta = std::cos(theta), stheta = std::sin(theta),
         cpsi = std::cos(psi), spsi = std::sin(psi);
  double dphi[9] = {0., sphi * spsi + cphi * cpsi * stheta,  cphi * spsi - cpsi * sphi * stheta,
                    0., -cpsi * sphi + cphi * spsi * stheta, -cphi * cpsi - sphi * spsi * stheta,
                    0., cphi * ctheta,                       -ctheta * sphi};
  double dtheta[9] = {-cpsi * stheta, cpsi * ctheta * sphi, cphi * cpsi * ctheta,
                      -spsi * stheta, ctheta * sphi * spsi, cphi * ctheta * spsi,
                      -ctheta,        -sphi * stheta,       -cphi * stheta};
  double dpsi[9] = {-ctheta * spsi, -cphi * cpsi - sphi * spsi * stheta, cpsi * sphi - cphi * spsi * stheta,
                    cpsi * ctheta,  -cphi * spsi + cpsi * sphi * stheta, sphi * spsi + cphi * cpsi * stheta,
                    0.,             0.,                                  0.};
  auto inner = [&](const double* A) {
    double r = 0.;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r += A[3 * j + i] * R[3 * i + j];
    return r;
  };
  g[3] = inner(dphi);
  g[4] = inner(dtheta);
  g[5] = inner(dpsi);
}

/* The fixed reduction tree of the device kernels (icpslam_b200/csrc/gicp.cuh), restated so that both
 * sides see bit-identical sums: values are taken in source-index order (0.0 for points without a
 * correspondence) in blocks of 256; inside a block every 32-value group is reduced by the xor butterfly
 * 16, 8, 4, 2, 1, the eight group sums are added in order, and the block sums are added in order.
 * PCL adds serially; that differs in the last bits only, which PCL's own result does not survive either
 * (tests/test_oracle.py::test_gicp_is_roundoff_sensitive). */
double tree_sum(const double* v, size_t n) {
  double total = 0.0;
  for (size_t b = 0; b < n; b += 256) {
    double bs = 0.0;
    const size_t bend = std::min(n, b + 256);
    for (size_t w = b; w < bend; w += 32) {
      double t[32], u[32];
      for (int l = 0; l < 32; ++l) t[l] = (w + l < n) ? v[w + l] : 0.0;
      for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ o];
        std::memcpy(t, u, sizeof(t));
      }
      bs += t[0];
    }
    total += bs;
  }
  return total;
}

/* ---- evaluation trace (tests only): every cost-functor evaluation appends x[6], f, |g| (NaN when the value was
 * not asked for) to a caller's buffer, so that an independent restatement can be compared step by step ------ */
double* g_trace = nullptr;
size_t g_trace_cap = 0, g_trace_n = 0;

/* ---- the cost functor (OptimizationFunctorWithIndices) ------------------------------------------ */
struct Functor {
  const float* src;                  /* `output` = untransformed source, xyzw */
  const float* tgt;
  size_t ns = 0;
  const int32_t* corr = nullptr;     /* per source point: target index or -1 */
  long m = 0;                        /* number of correspondences */
  const std::vector<double>* mahalanobis; /* 9 per source point */
  float base[16];                    /* base_transformation_ */
  int threads;
  bool exact_double = false;
  long evals = 0;
  std::vector<double> terms;         /* 13 x ns: f, g0..g2, R00..R22 */

  void fdf(const double* x, double* f, double* g) {
    ++evals;
    float Tx[16];
    std::memcpy(Tx, base, sizeof(Tx));
    apply_state(Tx, x);
    terms.assign(13 * ns, 0.0);
    for (size_t i = 0; i < ns; ++i) {
      if (corr[i] < 0) continue;
      const float* ps = src + 4 * i;
      const float* pt = tgt + 4 * (size_t)corr[i];
      double res[3];
      if (!exact_double) {
        float pp[3];
        xform_f(Tx, ps, pp);
        res[0] = (double)(pp[0] - pt[0]);
        res[1] = (double)(pp[1] - pt[1]);
        res[2] = (double)(pp[2] - pt[2]);
      } else {
        /* sensitivity experiment only (params.reserved[0] = 1): the same residual with a double
         * rotation and double arithmetic — what a closed-form normal-equation evaluation would see */
        const double cr = std::cos(x[3]), sr = std::sin(x[3]), cp = std::cos(x[4]), sp = std::sin(x[4]),
                     cy = std::cos(x[5]), sy = std::sin(x[5]);
        const double Rd[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                              sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                              -sp,     cp * sr,                cp * cr};
        for (int r = 0; r < 3; ++r)
          res[r] = Rd[3 * r] * ps[0] + Rd[3 * r + 1] * ps[1] + Rd[3 * r + 2] * ps[2] + x[r] - (double)pt[r];
      }
      const double* M = &(*mahalanobis)[9 * i];
      double tmp[3] = {M[0] * res[0] + M[1] * res[1] + M[2] * res[2], M[3] * res[0] + M[4] * res[1] + M[5] * res[2],
                       M[6] * res[0] + M[7] * res[1] + M[8] * res[2]};
      terms[0 * ns + i] = res[0] * tmp[0] + res[1] * tmp[1] + res[2] * tmp[2];
      terms[1 * ns + i] = tmp[0];
      terms[2 * ns + i] = tmp[1];
      terms[3 * ns + i] = tmp[2];
      float pb[3];
      xform_f(base, ps, pb);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) terms[(4 + 3 * r + c) * ns + i] = (double)pb[r] * tmp[c];
    }
    double S[13];
    for (int k = 0; k < 13; ++k) S[k] = tree_sum(&terms[(size_t)k * ns], ns);
    if (f) *f = S[0] / (double)m;
    if (g) {
      g[0] = S[1] * 2.0 / (double)m;
      g[1] = S[2] * 2.0 / (double)m;
      g[2] = S[3] * 2.0 / (double)m;
      double R[9];
      for (int k = 0; k < 9; ++k) R[k] = S[4 + k] * (2.0 / (double)m);
      compute_r_derivative(x, R, g);
    }
    if (g_trace && g_trace_n < g_trace_cap) {
      double* r = g_trace + 8 * g_trace_n++;
      for (int k = 0; k < 6; ++k) r[k] = x[k];
      r[6] = f ? *f : std::numeric_limits<double>::quiet_NaN();
      r[7] = g ? std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + g[3] * g[3] + g[4] * g[4] + g[5] * g[5])
               : std::numeric_limits<double>::quiet_NaN();
    }
  }
};

/* ---- BFGS (GSL vector_bfgs2 as ported by PCL) ------------------------------------------------------ */
struct BFGS {
  static constexpr int N = 6;
  Functor* fn;
  double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5;
  int order = 3;
  /* state */
  double f = 0, g0norm = 0, pnorm = 0, fp0 = 0, delta_f = 0;
  double gradient[N], x0[N], g0[N], p[N], dx[N];
  /* wrapper caches */
  double x_alpha[N], g_alpha[N], f_alpha = 0, df_alpha = 0;
  double f_cache_key = 0, df_cache_key = 0, x_cache_key = 0, g_cache_key = 0;

  static double dot(const double* a, const double* b) {
    double s = 0;
    for (int i = 0; i < N; ++i) s += a[i] * b[i];
    return s;
  }
  static double norm(const double* a) { return std::sqrt(dot(a, a)); }

  double slope() const { return dot(g_alpha, p); }
  void moveto(double alpha) {
    if (alpha == x_cache_key) return;
    for (int i = 0; i < N; ++i) x_alpha[i] = x0[i] + alpha * p[i];
    x_cache_key = alpha;
  }
  double wrap_f(double alpha) {
    if (alpha == f_cache_key) return f_alpha;
    moveto(alpha);
    fn->fdf(x_alpha, &f_alpha, nullptr);
    f_cache_key = alpha;
    return f_alpha;
  }
  double wrap_df(double alpha) {
    if (alpha == df_cache_key) return df_alpha;
    moveto(alpha);
    if (alpha != g_cache_key) {
      double dummy;
      fn->fdf(x_alpha, &dummy, g_alpha); /* PCL's df(): a pass that also forms the residuals */
      g_cache_key = alpha;
    }
    df_alpha = slope();
    df_cache_key = alpha;
    return df_alpha;
  }
  void wrap_fdf(double alpha, double* fo, double* dfo) {
    if (alpha == f_cache_key && alpha == df_cache_key) {
      *fo = f_alpha;
      *dfo = df_alpha;
      return;
    }
    if (alpha == f_cache_key || alpha == df_cache_key) {
      *fo = wrap_f(alpha);
      *dfo = wrap_df(alpha);
      return;
    }
    moveto(alpha);
    fn->fdf(x_alpha, &f_alpha, g_alpha);
    f_cache_key = alpha;
    g_cache_key = alpha;
    df_alpha = slope();
    df_cache_key = alpha;
    *fo = f_alpha;
    *dfo = df_alpha;
  }
  void change_direction() {
    std::memcpy(x_alpha, x0, sizeof(x_alpha));
    x_cache_key = 0.0;
    f_cache_key = 0.0;
    std::memcpy(g_alpha, g0, sizeof(g_alpha));
    g_cache_key = 0.0;
    df_alpha = slope();
    df_cache_key = 0.0;
  }

  void init(const double* x) {
    delta_f = 0;
    std::memset(dx, 0, sizeof(dx));
    fn->fdf(x, &f, gradient);
    std::memcpy(x0, x, sizeof(x0));
    std::memcpy(g0, gradient, sizeof(g0));
    g0norm = norm(g0);
    for (int i = 0; i < N; ++i) p[i] = gradient[i] * (-1.0 / g0norm);
    pnorm = norm(p);
    fp0 = -g0norm;
    /* prepare wrapper */
    std::memcpy(x_alpha, x0, sizeof(x_alpha));
    x_cache_key = 0;
    f_alpha = f;
    f_cache_key = 0;
    std::memcpy(g_alpha, g0, sizeof(g_alpha));
    g_cache_key = 0;
    df_alpha = slope();
    df_cache_key = 0;
  }

  /* gsl_poly_solve_quadratic */
  static int solve_quadratic(double a, double b, double c, double* r0, double* r1) {
    if (a == 0) {
      if (b == 0) return 0;
      *r0 = -c / b;
      return 1;
    }
    double disc = b * b - 4 * a * c;
    if (disc > 0) {
      if (b == 0) {
        double r = std::sqrt(-c / a);
        *r0 = -r;
        *r1 = r;
      } else {
        double sgnb = (b > 0 ? 1 : -1);
        double temp = -0.5 * (b + sgnb * std::sqrt(disc));
        double ra = temp / a, rb = c / temp;
        if (ra < rb) {
          *r0 = ra;
          *r1 = rb;
        } else {
          *r0 = rb;
          *r1 = ra;
        }
      }
      return 2;
    } else if (disc == 0) {
      *r0 = -0.5 * b / a;
      *r1 = -0.5 * b / a;
      return 2;
    }
    return 0;
  }
  static double cubic(double c0, double c1, double c2, double c3, double z) { return c0 + z * (c1 + z * (c2 + z * c3)); }
  static void check_extremum(double c0, double c1, double c2, double c3, double z, double* zmin, double* fmin) {
    double y = cubic(c0, c1, c2, c3, z);
    if (y < *fmin) {
      *zmin = z;
      *fmin = y;
    }
  }
  static double interp_cubic(double f0, double fp0, double f1, double fp1, double zl, double zh) {
    double eta = 3 * (f1 - f0) - 2 * fp0 - fp1;
    double xi = fp0 + fp1 - 2 * (f1 - f0);
    double c0 = f0, c1 = fp0, c2 = eta, c3 = xi;
    double zmin = zl, fmin = cubic(c0, c1, c2, c3, zl);
    check_extremum(c0, c1, c2, c3, zh, &zmin, &fmin);
    double z0 = 0, z1 = 0;
    int n = solve_quadratic(3 * c3, 2 * c2, c1, &z0, &z1);
    if (n == 2) {
      if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
      if (z1 > zl && z1 < zh) check_extremum(c0, c1, c2, c3, z1, &zmin, &fmin);
    } else if (n == 1) {
      if (z0 > zl && z0 < zh) check_extremum(c0, c1, c2, c3, z0, &zmin, &fmin);
    }
    return zmin;
  }
  static double interp_quad(double f0, double fp0, double f1, double zl, double zh) {
    double fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0));
    double fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0));
    double c = 2 * (f1 - f0 - fp0);
    double zmin = zl, fmin = fl;
    if (fh < fmin) {
      zmin = zh;
      fmin = fh;
    }
    if (c > 0) {
      double z = -fp0 / c;
      if (z > zl && z < zh) {
        double fz = f0 + z * (fp0 + z * (f1 - f0 - fp0));
        if (fz < fmin) {
          zmin = z;
          fmin = fz;
        }
      }
    }
    return zmin;
  }
  double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax) const {
    double zmin = (xmin - a) / (b - a), zmax = (xmax - a) / (b - a);
    if (zmin > zmax) std::swap(zmin, zmax);
    double z;
    if (order > 2 && std::isfinite(fpb))
      z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zmin, zmax);
    else
      z = interp_quad(fa, fpa * (b - a), fb, zmin, zmax);
    return a + z * (b - a);
  }

  /* linear_minimize.c: minimize() */
  int line_search(double alpha1, double* alpha_new) {
    double f0_, fp0_, falpha, falpha_prev, fpalpha = 0, fpalpha_prev, delta, alpha_next;
    double alpha = alpha1, alpha_prev = 0.0;
    double a, b, fa, fb, fpa, fpb;
    const int bracket_iters = 100, section_iters = 100;
    int i = 0;
    wrap_fdf(0.0, &f0_, &fp0_);
    falpha_prev = f0_;
    fpalpha_prev = fp0_;
    a = 0.0;
    b = alpha;
    fa = f0_;
    fb = 0.0;
    fpa = fp0_;
    fpb = 0.0;
    while (i++ < bracket_iters) {
      falpha = wrap_f(alpha);
      if (falpha > f0_ + alpha * rho * fp0_ || falpha >= falpha_prev) {
        a = alpha_prev;
        fa = falpha_prev;
        fpa = fpalpha_prev;
        b = alpha;
        fb = falpha;
        fpb = std::numeric_limits<double>::quiet_NaN();
        break;
      }
      fpalpha = wrap_df(alpha);
      if (std::fabs(fpalpha) <= -sigma * fp0_) {
        *alpha_new = alpha;
        return kSuccess;
      }
      if (fpalpha >= 0) {
        a = alpha;
        fa = falpha;
        fpa = fpalpha;
        b = alpha_prev;
        fb = falpha_prev;
        fpb = fpalpha_prev;
        break;
      }
      delta = alpha - alpha_prev;
      {
        double lower = alpha + delta, upper = alpha + tau1 * delta;
        alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, lower, upper);
      }
      alpha_prev = alpha;
      falpha_prev = falpha;
      fpalpha_prev = fpalpha;
      alpha = alpha_next;
    }
    while (i++ < section_iters) {
      delta = b - a;
      {
        double lower = a + tau2 * delta, upper = b - tau3 * delta;
        alpha = interpolate(a, fa, fpa, b, fb, fpb, lower, upper);
      }
      falpha = wrap_f(alpha);
      if ((a - alpha) * fpa <= kDblEps) return kNoProgress;
      if (falpha > f0_ + rho * alpha * fp0_ || falpha >= fa) {
        b = alpha;
        fb = falpha;
        fpb = std::numeric_limits<double>::quiet_NaN();
      } else {
        fpalpha = wrap_df(alpha);
        if (std::fabs(fpalpha) <= -sigma * fp0_) {
          *alpha_new = alpha;
          return kSuccess;
        }
        if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
          b = a;
          fb = fa;
          fpb = fpa;
          a = alpha;
          fa = falpha;
          fpa = fpalpha;
        } else {
          a = alpha;
          fa = falpha;
          fpa = fpalpha;
        }
      }
    }
    return kSuccess;
  }

  int one_step(double* x) {
    double alpha = 0.0, alpha1;
    double f0_ = f;
    if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) {
      std::memset(dx, 0, sizeof(dx));
      return kNoProgress;
    }
    if (delta_f < 0) {
      double del = std::max(-delta_f, 10 * kDblEps * std::fabs(f0_));
      alpha1 = std::min(1.0, 2.0 * del / (-fp0));
    } else {
      alpha1 = std::fabs(kStepSize);
    }
    int status = line_search(alpha1, &alpha);
    if (status != kSuccess) return status;
    /* updatePosition */
    {
      double fa_, dfa_;
      wrap_fdf(alpha, &fa_, &dfa_);
      f = f_alpha;
      std::memcpy(x, x_alpha, sizeof(x_alpha));
      std::memcpy(gradient, g_alpha, sizeof(g_alpha));
    }
    delta_f = f - f0_;
    {
      double dx0[N], dg0[N];
      for (int i = 0; i < N; ++i) {
        dx0[i] = x[i] - x0[i];
        dx[i] = dx0[i];
        dg0[i] = gradient[i] - g0[i];
      }
      double dxg = dot(dx0, gradient), dgg = dot(dg0, gradient), dxdg = dot(dx0, dg0), dgnorm = norm(dg0), A, B;
      if (dxdg != 0) {
        B = dxg / dxdg;
        A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
      } else {
        B = 0;
        A = 0;
      }
      for (int i = 0; i < N; ++i) p[i] = gradient[i] - A * dx0[i] - B * dg0[i];
    }
    std::memcpy(g0, gradient, sizeof(g0));
    std::memcpy(x0, x, sizeof(x0));
    g0norm = norm(g0);
    pnorm = norm(p);
    double dir = (dot(p, gradient) >= 0) ? -1.0 : 1.0;  /* GSL vector_bfgs2: dir = (pg >= 0.0) ? -1.0 : +1.0 */
    for (int i = 0; i < N; ++i) p[i] *= dir / pnorm;
    pnorm = norm(p);
    fp0 = dot(p, g0);
    change_direction();
    return kSuccess;
  }
  int test_gradient(double epsabs) const { return norm(gradient) < epsabs ? kSuccess : kRunning; }
};

/* GICP::computeCovariances (Appendix A.2) */
int compute_covariances(const float* cloud, size_t n, const void* tree, int k, double eps, double* cov9) {
  if ((size_t)k > n) return B2ICP_ERR_TOO_FEW_POINTS;
  std::vector<int32_t> idx((size_t)k * n);
  std::vector<float> d2((size_t)k * n);
  int rc = b2o_kdtree_knn(tree, cloud, n, k, idx.data(), d2.data());
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i) {
    double mean[3] = {0, 0, 0};
    double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
    for (int j = 0; j < k; ++j) {
      const float* pt = cloud + 4 * (size_t)idx[(size_t)k * i + j];
      mean[0] += pt[0];
      mean[1] += pt[1];
      mean[2] += pt[2];
      c00 += pt[0] * pt[0]; /* float product, double accumulation (PCL: cov(0,0) += pt.x*pt.x) */
      c10 += pt[1] * pt[0];
      c11 += pt[1] * pt[1];
      c20 += pt[2] * pt[0];
      c21 += pt[2] * pt[1];
      c22 += pt[2] * pt[2];
    }
    for (int d = 0; d < 3; ++d) mean[d] /= (double)k;
    double C[9];
    C[0] = c00 / k - mean[0] * mean[0];
    C[3] = c10 / k - mean[1] * mean[0];
    C[4] = c11 / k - mean[1] * mean[1];
    C[6] = c20 / k - mean[2] * mean[0];
    C[7] = c21 / k - mean[2] * mean[1];
    C[8] = c22 / k - mean[2] * mean[2];
    C[1] = C[3];
    C[2] = C[6];
    C[5] = C[7];
    double U[9], s[3], V[9];
    b2o_svd3_cov(C, U, s, V);
    double* out = cov9 + 9 * i;
    for (int e = 0; e < 9; ++e) out[e] = 0;
    for (int col = 0; col < 3; ++col) {
      double v = col == 2 ? eps : 1.0;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] += v * U[3 * r + col] * U[3 * c + col];
    }
  }
  return B2ICP_OK;
}

}  // namespace

extern "C" void b2o_gicp_trace(double* buf, size_t cap_records) {
  g_trace = buf;
  g_trace_cap = buf ? cap_records : 0;
  g_trace_n = 0;
}
extern "C" size_t b2o_gicp_trace_count(void) { return g_trace_n; }

extern "C" int b2o_covariances(const float* xyzw, size_t n, int k, double gicp_epsilon, double* cov9) {
  if (!xyzw || !cov9 || k < 1) return B2ICP_ERR_INVALID_ARG;
  void* tree = b2o_kdtree_build(xyzw, n);
  int rc = compute_covariances(xyzw, n, tree, k, gicp_epsilon, cov9);
  b2o_kdtree_free(tree);
  return rc;
}

/* GICP::computeTransformation (Appendix A.2) behind Registration::align (A.1). */
int b2o_align_gicp_impl(const b2icp_params* p, const float* src, size_t ns, const float* tgt, size_t nt,
                        const float* guess16, b2icp_result* out, float* aligned, int record_iter, int32_t* corr_idx,
                        float* corr_d2, b2o_stage_ms* st, int threads) {
  const int k = p->k_correspondences;
  if ((size_t)k > ns || (size_t)k > nt) return B2ICP_ERR_TOO_FEW_POINTS;
  auto t0 = clk::now();
  void* tree_t = b2o_kdtree_build(tgt, nt);
  void* tree_s = b2o_kdtree_build(src, ns);
  st->build += ms_since(t0);

  auto tc = clk::now();
  std::vector<double> cov_t(9 * nt), cov_s(9 * ns);
  int rc = compute_covariances(tgt, nt, tree_t, k, p->gicp_epsilon, cov_t.data());
  if (!rc) rc = compute_covariances(src, ns, tree_s, k, p->gicp_epsilon, cov_s.data());
  st->covariances += ms_since(tc);
  b2o_kdtree_free(tree_s);
  if (rc) {
    b2o_kdtree_free(tree_t);
    return rc;
  }

  float guess[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (guess16) std::memcpy(guess, guess16, sizeof(guess));
  float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; /* transformation_ */
  float prevT[16];
  std::memcpy(prevT, T, sizeof(T));
  std::vector<double> mahal(9 * ns);
  for (size_t i = 0; i < ns; ++i)
    for (int e = 0; e < 9; ++e) mahal[9 * i + e] = (e % 4 == 0) ? 1.0 : 0.0;
  std::vector<float> query(4 * ns), d2(ns);
  std::vector<int32_t> nn(ns);
  std::vector<int32_t> corr_now(ns);
  const double dist_threshold = p->max_correspondence_distance * p->max_correspondence_distance;
  int iters = 0, converged = 0, status = B2ICP_OK, n_corr = 0;
  double mse = std::numeric_limits<double>::quiet_NaN();
  (void)threads;

  while (!converged) {
    /* transform_R = double(transformation_) * double(guess), explicit triple loop */
    double TR[16];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int kk = 0; kk < 4; ++kk) s += (double)T[4 * i + kk] * (double)guess[4 * kk + j];
        TR[4 * i + j] = s;
      }
    double R[9] = {TR[0], TR[1], TR[2], TR[4], TR[5], TR[6], TR[8], TR[9], TR[10]};
    auto tt = clk::now();
    for (size_t i = 0; i < ns; ++i) {
      float g1[4] = {0, 0, 0, 1};
      xform_f(guess, src + 4 * i, g1);
      xform_f(T, g1, &query[4 * i]);
      query[4 * i + 3] = 1.0f;
    }
    st->transform += ms_since(tt);
    auto tn = clk::now();
    b2o_kdtree_nn(tree_t, query.data(), ns, nn.data(), d2.data());
    st->nn += ms_since(tn);

    auto ts = clk::now();
    double dsum = 0;
    n_corr = 0;
    for (size_t i = 0; i < ns; ++i) {
      const bool keep = nn[i] >= 0 && (double)d2[i] < dist_threshold; /* strict < */
      if (keep) {
        const double* C1 = &cov_s[9 * i];
        const double* C2 = &cov_t[9 * (size_t)nn[i]];
        double M[9], temp[9];
        mat3_mul(R, C1, M);
        mat3_mul_bt(M, R, temp);
        for (int e = 0; e < 9; ++e) temp[e] += C2[e];
        mat3_inv(temp, &mahal[9 * i]);
        ++n_corr;
        dsum += (double)d2[i];
      }
      corr_now[i] = keep ? nn[i] : -1;
      if ((record_iter == iters || record_iter < 0) && corr_idx) corr_idx[i] = keep ? nn[i] : -1;
      if ((record_iter == iters || record_iter < 0) && corr_d2) corr_d2[i] = d2[i];
    }
    mse = n_corr ? dsum / n_corr : std::numeric_limits<double>::quiet_NaN();
    std::memcpy(prevT, T, sizeof(T));
    if (n_corr < 4) { /* NotEnoughPointsException -> break, converged_ stays false */
      status = B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES;
      st->solve += ms_since(ts);
      break;
    }
    /* estimateRigidTransformationBFGS */
    double x[6] = {(double)T[3], (double)T[7], (double)T[11], std::atan2((double)T[9], (double)T[10]),
                   std::asin(-(double)T[8]), std::atan2((double)T[4], (double)T[0])};
    Functor fn;
    fn.src = src;
    fn.tgt = tgt;
    fn.ns = ns;
    fn.corr = corr_now.data();
    fn.m = n_corr;
    fn.mahalanobis = &mahal;
    std::memcpy(fn.base, guess, sizeof(guess));
    fn.threads = threads;
    fn.exact_double = p->reserved[0] == 1;
    BFGS bfgs;
    bfgs.fn = &fn;
    int inner = 0, result;
    bfgs.init(x);
    do {
      ++inner;
      result = bfgs.one_step(x);
      if (result) break;
      result = bfgs.test_gradient(kGradientTol);
    } while (result == kRunning && inner < p->max_inner_iterations);
    if (!(result == kNoProgress || result == kSuccess || inner == p->max_inner_iterations)) {
      status = B2ICP_ERR_SOLVER_FAILED; /* SolverDidntConvergeException */
      st->solve += ms_since(ts);
      break;
    }
    for (int e = 0; e < 16; ++e) T[e] = (e % 5 == 0) ? 1.f : 0.f;
    apply_state(T, x);
    st->solve += ms_since(ts);

    /* delta test */
    double delta = 0.;
    for (int kk = 0; kk < 4; ++kk)
      for (int l = 0; l < 4; ++l) {
        double ratio = (kk < 3 && l < 3) ? 1. / p->rotation_epsilon : 1. / p->transformation_epsilon;
        double c_delta = ratio * std::fabs((double)(prevT[4 * kk + l] - T[4 * kk + l]));
        if (c_delta > delta) delta = c_delta;
      }
    ++iters;
    if (iters >= p->max_iterations || delta < 1) {
      converged = 1;
      std::memcpy(prevT, T, sizeof(T));
    }
  }
  b2o_kdtree_free(tree_t);

  /* final_transformation_ = previous_transformation_ * guess (Matrix4f product) */
  float fin[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int kk = 0; kk < 4; ++kk) s += prevT[4 * r + kk] * guess[4 * kk + c];
      fin[4 * r + c] = s;
    }
  for (int e = 0; e < 16; ++e) out->T[e] = (double)fin[e];
  out->converged = converged;
  out->iterations = iters;
  out->n_corr_last = n_corr;
  out->status_detail = status;
  out->mse_last = mse;
  if (aligned) {
    auto tt = clk::now();
    for (size_t i = 0; i < ns; ++i) {
      xform_f(fin, src + 4 * i, aligned + 4 * i);
      aligned[4 * i + 3] = 1.0f;
    }
    st->transform += ms_since(tt);
  }
  return status;
}
