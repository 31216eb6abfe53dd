// gicp_host.inl — host half of the GICP mode (included by b2icp.cu inside its anonymous namespace).
//
// pcl::GeneralizedIterativeClosestPoint::computeTransformation's outer loop and PCL's BFGS (a port of GSL
// vector_bfgs2 + linear_minimize.c) are O(1) scalar work per step and run here
// on the host in double precision with the C library's sin / cos / atan2 — the only per-point work, the
// cost functor, is a kernel.  The scalar recursion is written expression by expression in a fixed order because
// PCL's line search ends on a round-off test: the result is only reproducible when f, df and every scalar step
// are bit-identical.
//
// Batching: the recursion of ONE scan is a chain of ~100 dependent evaluations, each a few microseconds of device
// work behind a launch and a read-back — latency, not throughput.  A batch therefore runs every scan's recursion
// as its own FIBER (a private stack and a register switch): the unmodified blocking code of one scan runs until it
// needs an evaluation, posts its request (new correspondences and / or one cost-functor evaluation) and yields.  The
// scans of a batch are dealt to up to kGicpGroups GROUPS, each with its own stream: when every live fiber of a group
// has yielded, the coordinator serves the group's round — the cost functor of all its scans is ONE launch (tasks as a
// kernel parameter, blockIdx.y = scan, the 14 sums of a scan written into mapped pinned memory by its last CTA) — and
// goes on to the next group's fibers while the device works; it comes back, waits for the round's stream and resumes
// the group's fibers.  A scan's arithmetic does not depend on which other scans share its rounds or its group:
// results are bit-identical to the one-scan-at-a-time path of round 1.
struct GicpJob;
void gicp_yield(GicpJob* job);

struct GicpHostFunctor {
  b2icp_handle* h;
  ScanSlot* s;
  GridSlot* g;
  GicpJob* job = nullptr;
  float base[16];
  long evals = 0;
  int rc = 0;
  long m = 0;

  // GICP::applyState: R = AngleAxisf(x5,Z) * AngleAxisf(x4,Y) * AngleAxisf(x3,X) (float quaternions)
  static void apply_state(float* t, const double* x) {
    float hz = 0.5f * (float)x[5], hy = 0.5f * (float)x[4], hx = 0.5f * (float)x[3];
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
    const float tx = 2.f * q[1], ty = 2.f * q[2], tz = 2.f * q[3];
    const float twx = tx * q[0], twy = ty * q[0], twz = tz * q[0];
    const float txx = tx * q[1], txy = ty * q[1], txz = tz * q[1];
    const float tyy = ty * q[2], tyz = tz * q[2], tzz = tz * q[3];
    float R[9] = {1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx,
                  txz - twy,         tyz + twx, 1.f - (txx + tyy)};
    float n[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        float sum = 0.f;
        for (int k = 0; k < 3; ++k) sum += R[3 * r + k] * t[4 * k + c];
        n[3 * r + c] = sum;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) t[4 * r + c] = n[3 * r + c];
    t[3] += (float)x[0];
    t[7] += (float)x[1];
    t[11] += (float)x[2];
  }

  // GICP::computeRDerivative
  static void compute_r_derivative(const double* x, const double* R, double* g) {
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

  // one cost-functor evaluation: posted to the batch's round (see GicpJob), answered with the 14 sums
  void fdf(const double* x, double* f, double* g);
  const float4* g_tgt() const { return g->pts; }
};

// PCL's BFGS<FunctorType> (GSL vector_bfgs2), N = 6
struct GicpBFGS {
  static constexpr int N = 6;
  static constexpr double kStepSize = 1.0, kDblEps = 2.220446049250313e-16;
  enum { kRunning = -1, kSuccess = 0, kNoProgress = 1 };
  GicpHostFunctor* fn;
  double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5;
  int order = 3;
  double f = 0, g0norm = 0, pnorm = 0, fp0 = 0, delta_f = 0;
  double gradient[N], x0[N], g0[N], p[N], dx[N];
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
      fn->fdf(x_alpha, &dummy, g_alpha);
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
    std::memcpy(x_alpha, x0, sizeof(x_alpha));
    x_cache_key = 0;
    f_alpha = f;
    f_cache_key = 0;
    std::memcpy(g_alpha, g0, sizeof(g_alpha));
    g_cache_key = 0;
    df_alpha = slope();
    df_cache_key = 0;
  }
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
  static double interp_cubic(double f0, double fp0_, double f1, double fp1, double zl, double zh) {
    double eta = 3 * (f1 - f0) - 2 * fp0_ - fp1;
    double xi = fp0_ + fp1 - 2 * (f1 - f0);
    double c0 = f0, c1 = fp0_, c2 = eta, c3 = xi;
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
  static double interp_quad(double f0, double fp0_, double f1, double zl, double zh) {
    double fl = f0 + zl * (fp0_ + zl * (f1 - f0 - fp0_));
    double fh = f0 + zh * (fp0_ + zh * (f1 - f0 - fp0_));
    double c = 2 * (f1 - f0 - fp0_);
    double zmin = zl, fmin = fl;
    if (fh < fmin) {
      zmin = zh;
      fmin = fh;
    }
    if (c > 0) {
      double z = -fp0_ / c;
      if (z > zl && z < zh) {
        double fz = f0 + z * (fp0_ + z * (f1 - f0 - fp0_));
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

// K5 driver: covariances of `count` clouds (cloud i: grid gs[i], n[i] points) -> covs[i] (9 doubles per point, original
// order).  Three launches for the whole batch: every point on its cloud's fine grid; the points that grid could not
// settle on a coarse grid over the same cloud (cell x 4); what is still open, exhaustively.
int compute_covariances_batch(b2icp_handle* h, GridSlot* const* gs, const size_t* n, DeviceBuf* const* covs, int count) {
  const int k = h->params.k_correspondences;
  if (k < 1 || k > kMaxK) return fail(h, B2ICP_ERR_INVALID_ARG, "k_correspondences must be in [1, 32]");
  if (count < 1) return B2ICP_OK;
  if (count > kMaxBatch) return fail(h, B2ICP_ERR_INVALID_ARG, "too many clouds");
  size_t max_n = 0, total = 0;
  for (int i = 0; i < count; ++i) {
    if ((size_t)k > n[i]) return fail(h, B2ICP_ERR_TOO_FEW_POINTS, "fewer points than k_correspondences");
    CK(covs[i]->ensure(n[i] * 9 * sizeof(double)));
    max_n = std::max(max_n, n[i]);
    total += n[i];
  }
  CK(h->knn_tasks.ensure((size_t)2 * count * sizeof(KnnTask)));
  CK(h->unres_list.ensure(total * sizeof(int2)));
  CK(h->knn_list2.ensure(total * sizeof(int2)));
  CK(h->knn_counts.ensure((2 + (size_t)kMaxBatch) * sizeof(unsigned int)));  // two list lengths + one point counter per cloud
  std::vector<KnnTask> tasks((size_t)2 * count);
  for (int i = 0; i < count; ++i) {
    KnnTask& f = tasks[(size_t)i];
    f.g = gs[i]->view;
    f.cloud = gs[i]->pts;
    f.cov = covs[i]->as<double>();
    f.n = (int)n[i];
    f.pad = 0;
  }
  // the coarse grids of the second pass: same clouds, same boxes, four times the cell edge.  Box and cell are
  // given, so their builds need no read-back and stay asynchronous in front of the first pass.
  std::vector<GridSlot*> cg((size_t)count);
  std::vector<size_t> cn(n, n + count);
  for (int i = 0; i < count; ++i) {
    GridSlot& c = gslot(h, (size_t)(2 * kMaxBatch + 16 + i));
    c.pts = gs[i]->pts;
    c.force_cell = 4.0 * gs[i]->cell;
    for (int d = 0; d < 3; ++d) {
      c.mn[d] = gs[i]->mn[d];
      c.mx[d] = gs[i]->mx[d];
    }
    c.bbox_known = true;
    cg[(size_t)i] = &c;
  }
  int rc = build_grids(h, cg.data(), cn.data(), count);
  if (rc) return rc;
  for (int i = 0; i < count; ++i) {
    tasks[(size_t)(count + i)] = tasks[(size_t)i];
    tasks[(size_t)(count + i)].g = cg[(size_t)i]->view;
  }
  const KnnTask* fine = h->knn_tasks.as<KnnTask>();
  const KnnTask* coarse = fine + count;
  unsigned int* c1 = h->knn_counts.as<unsigned int>();
  unsigned int* c2 = c1 + 1;
  // (pageable source: the copy is staged by the driver before the call returns)
  CK(cudaMemcpyAsync(h->knn_tasks.p, tasks.data(), tasks.size() * sizeof(KnnTask), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->knn_counts.p, 0, (2 + (size_t)kMaxBatch) * sizeof(unsigned int), h->stream));
  // k = 20 (PCL's default, what the reference runs) has its list in registers; any other k takes the general path
  const dim3 kgrid((unsigned)((max_n + 127) / 128), (unsigned)count, 1);
  if (k == 20) {
    knn_cov_kernel<20><<<kgrid, 128, 0, h->stream>>>(fine, k, h->params.gicp_epsilon, 6, h->unres_list.as<int2>(), c1);
    knn_cov_list_kernel<20><<<148 * 8, 128, 0, h->stream>>>(coarse, k, h->params.gicp_epsilon, 8, h->unres_list.as<int2>(), c1,
                                                            h->knn_list2.as<int2>(), c2);
  } else {
    knn_cov_kernel<0><<<kgrid, 128, 0, h->stream>>>(fine, k, h->params.gicp_epsilon, 6, h->unres_list.as<int2>(), c1);
    knn_cov_list_kernel<0><<<148 * 8, 128, 0, h->stream>>>(coarse, k, h->params.gicp_epsilon, 8, h->unres_list.as<int2>(), c1,
                                                           h->knn_list2.as<int2>(), c2);
  }
  knn_cov_fallback<<<148 * 4, kKnnFbThreads, 0, h->stream>>>(fine, k, h->params.gicp_epsilon, h->knn_list2.as<int2>(), c2);
  // raw covariances -> U diag(1, 1, eps) U^T, lanes load-balanced over the points of each cloud
  const unsigned int sx = (unsigned int)std::max<size_t>(1, std::min<size_t>((max_n + 127) / 128, (size_t)(148 * 16 + count - 1) / (size_t)count));
  cov_svd_kernel<<<dim3(sx, (unsigned)count, 1), 128, 0, h->stream>>>(fine, h->params.gicp_epsilon, c1 + 2);
  h->launches += 1;
  h->launches += 3;
  return B2ICP_OK;
}

int compute_covariances(b2icp_handle* h, GridSlot& g, size_t n, DeviceBuf& cov) {
  GridSlot* gp = &g;
  DeviceBuf* cp = &cov;
  return compute_covariances_batch(h, &gp, &n, &cp, 1);
}

// One scan of a GICP batch: its slot, its fiber and the request it is blocked on.
struct GicpJob {
  b2icp_handle* h = nullptr;
  int slot_index = 0;
  float guess[16];
  // request of the current round
  bool want_corr = false, want_fdf = false, finished = false;
  GicpIterArgs iter;
  GicpEvalArgs eval;
  double S[kGicpSums];  // answer to want_fdf
  // outcome
  IcpState out;
  int rc = B2ICP_OK;
  long evals = 0;
  // fiber
  FiberCtx ctx, *back = nullptr;
  std::vector<char> stack;
};

void gicp_yield(GicpJob* job) { fiber_switch(job->ctx, *job->back); }

void GicpHostFunctor::fdf(const double* x, double* f, double* g) {
  ++evals;
  if (rc) {
    if (f) *f = 0;
    if (g) for (int i = 0; i < 6; ++i) g[i] = 0;
    return;
  }
  std::memcpy(job->eval.Tx, base, sizeof(base));
  std::memcpy(job->eval.base, base, sizeof(base));
  apply_state(job->eval.Tx, x);
  job->want_fdf = true;
  gicp_yield(job);  // back when the round has been served
  if (job->rc) {
    rc = job->rc;
    if (f) *f = 0;
    if (g) for (int i = 0; i < 6; ++i) g[i] = 0;
    return;
  }
  const double* S = job->S;
  m = (long)S[13];
  const double md = (double)m;
  if (f) *f = S[0] / md;
  if (g) {
    g[0] = S[1] * 2.0 / md;
    g[1] = S[2] * 2.0 / md;
    g[2] = S[3] * 2.0 / md;
    double R[9];
    for (int k = 0; k < 9; ++k) R[k] = S[4 + k] * (2.0 / md);
    compute_r_derivative(x, R, g);
  }
}

// GICP::computeTransformation of one scan (Registration::align around it), on the scan's fiber.
void gicp_job_body(GicpJob& job) {
  b2icp_handle* h = job.h;
  ScanSlot& s = slot(h, (size_t)job.slot_index);
  GridSlot& g = gslot(h, s.grid);
  float* guess = job.guess;
  float T[16], prevT[16];
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
  std::memcpy(prevT, T, sizeof(T));
  const b2icp_params& p = h->params;
  const double dist_threshold = p.max_correspondence_distance * p.max_correspondence_distance;
  int iters = 0, converged = 0, status = B2ICP_OK;
  long n_corr = 0;
  GicpHostFunctor fn;
  fn.h = h;
  fn.s = &s;
  fn.g = &g;
  fn.job = &job;
  std::memcpy(fn.base, guess, 16 * sizeof(float));
  const int max_rings = rings_for_bound(h, (double)g.view.cell);

  while (!converged) {
    GicpIterArgs& a = job.iter;
    std::memcpy(a.guess, guess, 16 * sizeof(float));
    std::memcpy(a.T, T, sizeof(T));
    double TR[16];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double sum = 0;
        for (int kk = 0; kk < 4; ++kk) sum += (double)T[4 * i + kk] * (double)guess[4 * kk + j];
        TR[4 * i + j] = sum;
      }
    const double R[9] = {TR[0], TR[1], TR[2], TR[4], TR[5], TR[6], TR[8], TR[9], TR[10]};
    std::memcpy(a.R, R, sizeof(R));
    a.max2 = dist_threshold;
    a.bound2 = h->cfg.bound2;
    a.max_rings = max_rings;
    a.use_seed = iters > 0;
    job.want_corr = true;  // served in the same round as the first evaluation below, before it
    std::memcpy(prevT, T, sizeof(T));
    // estimateRigidTransformationBFGS
    double x[6] = {(double)T[3], (double)T[7], (double)T[11], std::atan2((double)T[9], (double)T[10]),
                   std::asin(-(double)T[8]), std::atan2((double)T[4], (double)T[0])};
    GicpBFGS bfgs;
    bfgs.fn = &fn;
    bfgs.init(x);  // the first evaluation also counts the correspondences
    if (fn.rc) {
      job.rc = fn.rc;
      break;
    }
    n_corr = fn.m;
    if (n_corr < 4) {  // NotEnoughPointsException -> break, converged_ stays false
      status = B2ICP_ERR_NOT_ENOUGH_CORRESPONDENCES;
      break;
    }
    int inner = 0, result;
    do {
      ++inner;
      result = bfgs.one_step(x);
      if (result) break;
      result = bfgs.test_gradient(1e-2);
    } while (result == GicpBFGS::kRunning && inner < p.max_inner_iterations);
    if (fn.rc) {
      job.rc = fn.rc;
      break;
    }
    if (!(result == GicpBFGS::kNoProgress || result == GicpBFGS::kSuccess || inner == p.max_inner_iterations)) {
      status = B2ICP_ERR_SOLVER_FAILED;
      break;
    }
    for (int e = 0; e < 16; ++e) T[e] = (e % 5 == 0) ? 1.f : 0.f;
    GicpHostFunctor::apply_state(T, x);
    double delta = 0.;
    for (int kk = 0; kk < 4; ++kk)
      for (int l = 0; l < 4; ++l) {
        double ratio = (kk < 3 && l < 3) ? 1. / p.rotation_epsilon : 1. / p.transformation_epsilon;
        double c_delta = ratio * std::fabs((double)(prevT[4 * kk + l] - T[4 * kk + l]));
        if (c_delta > delta) delta = c_delta;
      }
    ++iters;
    if (iters >= p.max_iterations || delta < 1) {
      converged = 1;
      std::memcpy(prevT, T, sizeof(T));
    }
  }
  // final_transformation_ = previous_transformation_ * guess (Matrix4f product)
  IcpState& st = job.out;
  std::memset(&st, 0, sizeof(st));
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float sum = 0.f;
      for (int kk = 0; kk < 4; ++kk) sum += prevT[4 * r + kk] * guess[4 * kk + c];
      st.final_T[4 * r + c] = sum;
    }
  st.converged = converged;
  st.iter = iters;
  st.n_corr = (int)n_corr;
  st.status = status;
  st.mse = std::nan("");
  job.evals = fn.evals;
}

thread_local GicpJob* g_entry_job = nullptr;
void gicp_fiber_entry() {
  GicpJob* job = g_entry_job;
  gicp_job_body(*job);
  job->finished = true;
  gicp_yield(job);
}

// GICP::computeTransformation for the scans in slots [0, B): sources uploaded, target grids built.  guesses: B x 16
// floats or NULL.  Fills h->h_states[0 .. B) and mirrors them (and the tasks getFitnessScore reads) to the device.
int run_gicp_batch(b2icp_handle* h, int B, const float* guesses) {
  if (B < 1 || B > kMaxBatch) return fail(h, B2ICP_ERR_INVALID_ARG, "batch too large");
  const bool dbg = getenv("B2ICP_GICP_DEBUG") != nullptr;  // tuning only: host wall time of the phases
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t_start = now();
  // ---- setup: work buffers; target covariances (cached with their grids); then the grids and the covariances of all
  // the sources of the batch, each one launch sequence for the whole batch
  size_t max_n = 0;
  int setup_rc[kMaxBatch];
  const int k = h->params.k_correspondences;
  {
    std::vector<GridSlot*> tg;
    std::vector<size_t> tn;
    std::vector<DeviceBuf*> tc;
    for (int i = 0; i < B; ++i) {
      ScanSlot& s = slot(h, (size_t)i);
      GridSlot& g = gslot(h, s.grid);
      {
        int rc = ensure_grid(h, g);
        if (rc) return rc;
      }
      const size_t n = s.src.n;
      max_n = std::max(max_n, n);
      setup_rc[i] = ((size_t)k > n || (size_t)k > (size_t)g.view.n) ? B2ICP_ERR_TOO_FEW_POINTS : B2ICP_OK;
      int rc = ensure_slot_work(h, s);
      if (rc) return rc;
      CK(s.corr_pos.ensure(n * sizeof(int)));  // sorted position of the last match: the seed of the next search
      const int nblk = (int)((n + kGicpThreads - 1) / kGicpThreads);
      CK(s.mahal.ensure(n * 9 * sizeof(double)));
      CK(s.gicp_partials.ensure((size_t)nblk * kGicpSums * sizeof(double)));
      if (!setup_rc[i] && !g.cov_valid && std::find(tg.begin(), tg.end(), &g) == tg.end()) {
        tg.push_back(&g);
        tn.push_back((size_t)g.view.n);
        tc.push_back(&g.cov);
      }
    }
    if (!tg.empty()) {
      int rc = compute_covariances_batch(h, tg.data(), tn.data(), tc.data(), (int)tg.size());
      if (rc) return rc;
      for (GridSlot* g : tg) g->cov_valid = true;
    }
    std::vector<GridSlot*> sg;
    std::vector<size_t> sn;
    std::vector<DeviceBuf*> sc;
    for (int i = 0; i < B; ++i) {
      if (setup_rc[i]) continue;
      ScanSlot& s = slot(h, (size_t)i);
      GridSlot& g = gslot(h, (size_t)(kMaxBatch + 8 + i));  // the source's own grid: only its neighbour lists are used
      g.pts = s.src.dev();
      sg.push_back(&g);
      sn.push_back(s.src.n);
      sc.push_back(&s.cov);
    }
    if (dbg) {
      cudaStreamSynchronize(h->stream);
      fprintf(stderr, "[b2icp]   setup: slots + target covariances %.2f ms\n", ms(t_start, now()));
    }
    const auto t_sg = now();
    if (!sg.empty()) {
      int rc = build_grids(h, sg.data(), sn.data(), (int)sg.size());
      if (dbg) {
        cudaStreamSynchronize(h->stream);
        fprintf(stderr, "[b2icp]   setup: %d source grids %.2f ms\n", (int)sg.size(), ms(t_sg, now()));
      }
      if (rc == B2ICP_ERR_NONFINITE_INPUT) {  // find the offender: build them one by one
        rc = B2ICP_OK;
        for (size_t q = 0, i = 0; i < (size_t)B; ++i) {
          if (setup_rc[i]) continue;
          GridSlot* one = sg[q];
          size_t nn = sn[q];
          const int r1 = build_grids(h, &one, &nn, 1);
          if (r1 == B2ICP_ERR_CUDA) return r1;
          if (r1) setup_rc[i] = r1;
          ++q;
        }
        std::vector<GridSlot*> sg2;
        std::vector<size_t> sn2;
        std::vector<DeviceBuf*> sc2;
        for (size_t q = 0, i = 0; i < (size_t)B; ++i) {
          if (setup_rc[i] == B2ICP_ERR_TOO_FEW_POINTS) continue;
          if (!setup_rc[i]) {
            sg2.push_back(sg[q]);
            sn2.push_back(sn[q]);
            sc2.push_back(sc[q]);
          }
          ++q;
        }
        sg.swap(sg2);
        sn.swap(sn2);
        sc.swap(sc2);
      }
      if (rc) return rc;
      if (!sg.empty()) {
        rc = compute_covariances_batch(h, sg.data(), sn.data(), sc.data(), (int)sg.size());
        if (rc) return rc;
      }
    }
  }
  if (dbg) cudaStreamSynchronize(h->stream);
  const auto t_setup = now();
  // ---- round buffers: task arrays (pinned + device) and the sums that come back, one region per group
  constexpr size_t kTaskSlots = (size_t)kGicpGroups * kMaxBatch;
  const size_t task_bytes = kTaskSlots * sizeof(GicpCorrTask);
  if (h->h_gicp_tasks_cap < task_bytes) {
    if (h->h_gicp_tasks) cudaFreeHost(h->h_gicp_tasks);
    h->h_gicp_tasks = nullptr;
    h->h_gicp_tasks_cap = 0;
    CK(cudaMallocHost((void**)&h->h_gicp_tasks, task_bytes));
    h->h_gicp_tasks_cap = task_bytes;
  }
  if (h->h_gicp_partials_cap < kTaskSlots * kGicpSums) {
    if (h->h_gicp_partials) cudaFreeHost(h->h_gicp_partials);
    h->h_gicp_partials = nullptr;
    h->h_gicp_partials_cap = 0;
    CK(cudaMallocHost((void**)&h->h_gicp_partials, kTaskSlots * kGicpSums * sizeof(double)));
    h->h_gicp_partials_cap = kTaskSlots * kGicpSums;
  }
  CK(h->gicp_tasks.ensure(task_bytes));
  if (!h->gicp_tickets.p) {  // per-scan tickets of the cost-functor kernel's last-CTA sum: zero between launches
    CK(h->gicp_tickets.ensure((size_t)kMaxBatch * sizeof(unsigned int)));
    CK(cudaMemsetAsync(h->gicp_tickets.p, 0, (size_t)kMaxBatch * sizeof(unsigned int), h->stream));
  }
  GicpCorrTask* h_corr = reinterpret_cast<GicpCorrTask*>(h->h_gicp_tasks);
  GicpCorrTask* d_corr = h->gicp_tasks.as<GicpCorrTask>();

  // ---- one fiber per scan
  std::vector<std::unique_ptr<GicpJob>> jobs;
  FiberCtx main_ctx;
  for (int i = 0; i < B; ++i) {
    jobs.emplace_back(new GicpJob());
    GicpJob& j = *jobs.back();
    j.h = h;
    j.slot_index = i;
    for (int k = 0; k < 16; ++k) j.guess[k] = guesses ? guesses[16 * i + k] : ((k % 5 == 0) ? 1.f : 0.f);
    if (setup_rc[i]) {  // no fiber: the scan reports its setup failure (PCL: PCL_ERROR and converged_ = false)
      j.finished = true;
      j.rc = setup_rc[i];
      std::memset(&j.out, 0, sizeof(j.out));
      for (int k = 0; k < 16; ++k) j.out.final_T[k] = j.guess[k];
      j.out.status = setup_rc[i];
      j.out.mse = std::nan("");
      continue;
    }
    j.stack.resize(256 * 1024);
    j.back = &main_ctx;
    fiber_make(j.ctx, main_ctx, j.stack.data(), j.stack.size(), gicp_fiber_entry);
  }
  // ---- groups: the live scans are dealt to up to kGicpGroups groups, each with its own stream and its own region
  // of the task / sum buffers.  While the device serves one group's round the host runs the other groups' fibers
  // (BFGS updates, the next requests), so neither side waits for the other; a scan's arithmetic does not depend on
  // which group it is in.
  struct Group {
    std::vector<int> members;
    cudaStream_t st = nullptr;
    bool pending = false;
    int live = 0, nf = 0;
    int fdf_of[kMaxBatch];
  };
  int live = 0;
  for (auto& jp : jobs) live += jp->finished ? 0 : 1;
  const int ngroups = std::max(1, std::min(std::min(h->gicp_groups, kGicpGroups), live));
  Group groups[kGicpGroups];
  for (int i = 0, q = 0; i < B; ++i) {
    if (jobs[i]->finished) continue;
    Group& G = groups[q++ % ngroups];
    G.members.push_back(i);
    ++G.live;
  }
  CK(cudaStreamSynchronize(h->stream));  // the groups' streams start after the setup
  for (int gi = 0; gi < ngroups; ++gi) {
    if (!h->gicp_streams[gi]) CK(cudaStreamCreateWithFlags(&h->gicp_streams[gi], cudaStreamNonBlocking));
    groups[gi].st = h->gicp_streams[gi];
  }
  long rounds = 0;
  int rc_round = B2ICP_OK;
  double t_host = 0.0, t_wait = 0.0;
  auto cuda_failed = [&](cudaError_t e) {
    rc_round = B2ICP_ERR_CUDA;
    h->err = std::string("GICP round: ") + cudaGetErrorString(e);
    for (auto& jp : jobs) jp->rc = B2ICP_ERR_CUDA;  // every fiber unwinds at its next evaluation
  };
  // run the group's live fibers up to their next request, then put the round on the group's stream:
  // correspondences first, then the evaluations, in stream order
  auto advance_and_launch = [&](int gi) {
    Group& G = groups[gi];
    const auto t0 = now();
    for (int i : G.members) {
      GicpJob& j = *jobs[i];
      if (j.finished) continue;
      j.want_corr = j.want_fdf = false;
      g_entry_job = &j;
      fiber_switch(main_ctx, j.ctx);
      if (j.finished) {
        --G.live;
        --live;
      }
    }
    const size_t off = (size_t)gi * kMaxBatch;
    int nc = 0, nf = 0;
    for (int i : G.members) {
      GicpJob& j = *jobs[i];
      if (j.finished) continue;
      ScanSlot& s = slot(h, (size_t)i);
      GridSlot& g = gslot(h, s.grid);
      const int n = (int)s.src.n;
      if (j.want_corr) {
        GicpCorrTask& t = h_corr[off + nc++];
        t.g = g.view;
        t.src = s.src.dev();
        t.cov_src = s.cov.as<double>();
        t.cov_tgt = g.cov.as<double>();
        t.mahal = s.mahal.as<double>();
        t.corr_idx = s.corr_idx.as<int>();
        t.corr_d2 = s.corr_d2.as<float>();
        t.corr_pos = s.corr_pos.as<int>();
        t.n = n;
        t.pad = 0;
        t.a = j.iter;
      }
      if (j.want_fdf) G.fdf_of[nf++] = i;
    }
    G.nf = nf;
    G.pending = false;
    if (nc == 0 && nf == 0) {
      t_host += ms(t0, now());
      return;
    }
    ++rounds;
    cudaError_t e = cudaSuccess;
    if (nc) {
      e = cudaMemcpyAsync(d_corr + off, h_corr + off, (size_t)nc * sizeof(GicpCorrTask), cudaMemcpyHostToDevice, G.st);
      gicp_corr_kernel<<<dim3((unsigned)((max_n + kSweepThreads - 1) / kSweepThreads), (unsigned)nc, 1), kSweepThreads, 0, G.st>>>(d_corr + off);
      h->launches += 1;
    }
    // evaluations: the tasks are kernel parameters and the 14 sums of every scan land in pinned host memory
    // (mapped, written by the scan's last CTA): one launch per kGicpParamTasks scans, nothing else on the stream
    for (int k0 = 0; k0 < nf && e == cudaSuccess; k0 += kGicpParamTasks) {
      const int cnt = std::min(kGicpParamTasks, nf - k0);
      GicpFdfBatch fb;
      for (int k = 0; k < cnt; ++k) {
        const int i = G.fdf_of[k0 + k];
        ScanSlot& s = slot(h, (size_t)i);
        GicpFdfTask& t = fb.t[k];
        t.src = s.src.dev();
        t.tgt = gslot(h, s.grid).pts;
        t.corr_idx = s.corr_idx.as<int>();
        t.mahal = s.mahal.as<double>();
        t.partials = s.gicp_partials.as<double>();
        t.sums = h->h_gicp_partials + (off + (size_t)(k0 + k)) * kGicpSums;
        t.ticket = h->gicp_tickets.as<unsigned int>() + i;
        t.n = (int)s.src.n;
        t.pad = 0;
        t.a = jobs[i]->eval;
      }
      gicp_fdf_kernel<<<dim3((unsigned)((max_n + kGicpThreads - 1) / kGicpThreads), (unsigned)cnt, 1), kGicpThreads, 0, G.st>>>(fb);
      h->launches += 1;
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) cuda_failed(e);
    G.pending = true;
    t_host += ms(t0, now());
  };
  auto collect = [&](int gi) {
    Group& G = groups[gi];
    const auto t0 = now();
    cudaError_t e = cudaStreamSynchronize(G.st);
    if (e == cudaSuccess) e = cudaGetLastError();
    t_wait += ms(t0, now());
    if (e != cudaSuccess) cuda_failed(e);
    const size_t off = (size_t)gi * kMaxBatch;
    for (int k = 0; k < G.nf; ++k)
      std::memcpy(jobs[G.fdf_of[k]]->S, h->h_gicp_partials + (off + k) * kGicpSums, kGicpSums * sizeof(double));
    G.pending = false;
  };
  for (int gi = 0; gi < ngroups; ++gi) advance_and_launch(gi);
  for (bool any = true; any;) {
    any = false;
    for (int gi = 0; gi < ngroups; ++gi) {
      Group& G = groups[gi];
      if (!G.pending) continue;
      collect(gi);
      if (G.live > 0) advance_and_launch(gi);
      any = any || G.pending;
    }
    for (int gi = 0; gi < ngroups; ++gi) any = any || groups[gi].pending;
  }
  if (dbg)
    fprintf(stderr, "[b2icp] GICP batch of %d in %d group(s): setup (grids + covariances) %.2f ms, %ld group rounds %.2f ms (host %.2f ms, waiting for the device %.2f ms)\n",
            B, ngroups, ms(t_start, t_setup), rounds, ms(t_setup, now()), t_host, t_wait);
  h->gicp_rounds = rounds;
  h->gicp_evals = 0;
  int worst = rc_round;
  for (int i = 0; i < B; ++i) {
    GicpJob& j = *jobs[i];
    h->gicp_evals += j.evals;
    h->h_states[i] = j.out;
    if (j.rc && worst == B2ICP_OK) worst = j.rc;
    // the device copy of the state feeds b2icp_fitness / the aligned-cloud transform
    ScanSlot& s = slot(h, (size_t)i);
    GridSlot& g = gslot(h, s.grid);
    ScanTask& t = h->h_tasks[i];
    t.grid = g.view;
    t.src = s.src.dev();
    t.cur = s.cur.as<float4>();
    t.corr_idx = s.corr_idx.as<int>();
    t.corr_d2 = s.corr_d2.as<float>();
    t.c0 = s.c0.as<float4>();
    t.c1 = s.c1.as<float4>();
    t.c2 = s.c2.as<float4>();
    t.partials = s.partials.as<double>();
    t.state = h->states.as<IcpState>() + i;
    t.n = (int)s.src.n;
    t.pad = 0;
  }
  h->timing.kernel_launches = h->launches;
  if (worst == B2ICP_ERR_CUDA) return worst;
  CK(cudaMemcpyAsync(h->tasks.p, h->h_tasks, sizeof(ScanTask) * B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->states.p, h->h_states, sizeof(IcpState) * B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return worst == B2ICP_OK ? B2ICP_OK : worst;
}

// GICP::computeTransformation for slot 0 (Registration::align around it); fills h->h_states[0].
int run_gicp(b2icp_handle* h, const float* guess16) { return run_gicp_batch(h, 1, guess16); }
