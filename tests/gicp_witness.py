"""An independent restatement of pcl::GeneralizedIterativeClosestPoint::computeTransformation and of PCL's BFGS
(GSL multimin/vector_bfgs2.c + linear_minimize.c), written from SURVEY.md Appendix A.2 / A.4 and the published GSL
algorithm — NOT from oracle/gicp_oracle.cpp — in numpy and plain Python.  It is the second witness of the GICP
oracle: tests/test_oracle.py asserts that both produce the same sequence of cost-functor evaluations (x, f, |g|).

The arithmetic definition is the one DESIGN.md states (it is what makes a step-by-step comparison possible at
all, PCL's line search ends on a round-off test): float32 point transforms in Eigen's operation order, double
Mahalanobis algebra expression by expression, cross-point sums over the fixed tree (256-point blocks, xor butterfly
inside every 32 values, 8 group sums in order, block sums in order), libm's sin / cos / atan2 / asin.
Test infrastructure only.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
from scipy.spatial import cKDTree

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
f32 = np.float32
DBL_EPS = 2.220446049250313e-16
SUCCESS, NO_PROGRESS, RUNNING = 0, 1, -1


# ---------------------------------------------------------------------------------------------------------------
# float32 / double building blocks
# ---------------------------------------------------------------------------------------------------------------
def xform_f(T, P):
    """Eigen Matrix4f * (x, y, z, 1): ((m0 x + m1 y) + m2 z) + m3 per row, every operation rounded to float32."""
    T = np.asarray(T, f32).reshape(4, 4)
    x, y, z = P[:, 0].astype(f32), P[:, 1].astype(f32), P[:, 2].astype(f32)
    rows = [((T[r, 0] * x + T[r, 1] * y) + T[r, 2] * z) + T[r, 3] for r in range(3)]
    return np.stack(rows, axis=1).astype(f32)


def apply_state(T, x):
    """GICP::applyState: R = AngleAxisf(x5, Z) * AngleAxisf(x4, Y) * AngleAxisf(x3, X) through float quaternions
    (Eigen: AngleAxis -> Quaternion (cos(a/2), axis sin(a/2)); Quaternion product; toRotationMatrix), then
    t.topLeft3x3 = R * t.topLeft3x3 and t.col(3) += (x0, x1, x2, 0)."""
    t = np.array(T, f32).reshape(4, 4).copy()

    def quat(angle, axis):
        h = f32(0.5) * f32(angle)
        q = [f32(_libm.cosf(h)), f32(0), f32(0), f32(0)]
        q[1 + axis] = f32(_libm.sinf(h))
        return q

    def qmul(a, b):  # (w, x, y, z)
        return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
                a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]]

    q = qmul(qmul(quat(x[5], 2), quat(x[4], 1)), quat(x[3], 0))
    w, qx, qy, qz = q
    two = f32(2)
    tx, ty, tz = two * qx, two * qy, two * qz
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    one = f32(1)
    R = [[one - (tyy + tzz), txy - twz, txz + twy],
         [txy + twz, one - (txx + tzz), tyz - twx],
         [txz - twy, tyz + twx, one - (txx + tyy)]]
    n = np.zeros((3, 3), f32)
    for r in range(3):
        for c in range(3):
            s = f32(0)
            for k in range(3):
                s = f32(s + R[r][k] * t[k, c])
            n[r, c] = s
    t[:3, :3] = n
    for r in range(3):
        t[r, 3] = f32(t[r, 3] + f32(x[r]))
    return t


def tree_sum(v):
    """The fixed reduction tree: blocks of 256 values in index order; inside a block every 32 values by the xor
    butterfly 16, 8, 4, 2, 1 (value of lane 0), the 8 group sums added in order, the block sums added in order."""
    v = np.asarray(v, np.float64)
    n = len(v)
    pad = (-n) % 256
    a = np.concatenate([v, np.zeros(pad)]).reshape(-1, 8, 32)
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        a = a + a[:, :, lanes ^ o]
    g = a[:, :, 0]  # [blocks, 8]
    bs = np.zeros(g.shape[0])
    for w in range(8):
        bs = bs + g[:, w]
    total = 0.0
    for b in bs:
        total = total + float(b)
    return total


def mahalanobis(Rm, C1, C2):
    """M_i = (R C1_i R^T + C2_i)^-1 in double: the triple products term by term, Eigen's cofactor inverse."""
    R = np.asarray(Rm, np.float64).reshape(3, 3)
    n = len(C1)
    M = np.zeros((n, 3, 3))
    for r in range(3):
        for c in range(3):
            M[:, r, c] = (R[r, 0] * C1[:, 0, c] + R[r, 1] * C1[:, 1, c]) + R[r, 2] * C1[:, 2, c]
    t = np.zeros((n, 3, 3))
    for r in range(3):
        for c in range(3):
            t[:, r, c] = ((M[:, r, 0] * R[c, 0] + M[:, r, 1] * R[c, 1]) + M[:, r, 2] * R[c, 2]) + C2[:, r, c]
    m = t.reshape(n, 9)
    c00 = m[:, 4] * m[:, 8] - m[:, 5] * m[:, 7]
    c01 = m[:, 5] * m[:, 6] - m[:, 3] * m[:, 8]
    c02 = m[:, 3] * m[:, 7] - m[:, 4] * m[:, 6]
    det = (m[:, 0] * c00 + m[:, 1] * c01) + m[:, 2] * c02
    idet = 1.0 / det
    inv = np.stack([c00 * idet, (m[:, 2] * m[:, 7] - m[:, 1] * m[:, 8]) * idet, (m[:, 1] * m[:, 5] - m[:, 2] * m[:, 4]) * idet,
                    c01 * idet, (m[:, 0] * m[:, 8] - m[:, 2] * m[:, 6]) * idet, (m[:, 2] * m[:, 3] - m[:, 0] * m[:, 5]) * idet,
                    c02 * idet, (m[:, 1] * m[:, 6] - m[:, 0] * m[:, 7]) * idet, (m[:, 0] * m[:, 4] - m[:, 1] * m[:, 3]) * idet],
                   axis=1)
    return inv.reshape(n, 3, 3)


# ---------------------------------------------------------------------------------------------------------------
# the cost functor (OptimizationFunctorWithIndices)
# ---------------------------------------------------------------------------------------------------------------
class Functor:
    def __init__(self, src, tgt, corr, M, base):
        self.src, self.tgt, self.corr, self.M, self.base = src, tgt, corr, M, np.asarray(base, f32).reshape(4, 4)
        self.keep = corr >= 0
        self.m = int(self.keep.sum())
        self.trace = []

    def fdf(self, x, want_f=True, want_g=True):
        n = len(self.src)
        Tx = apply_state(self.base, x)
        pp = xform_f(Tx, self.src)
        pt = self.tgt[np.where(self.keep, self.corr, 0), :3].astype(f32)
        res = (pp - pt).astype(np.float64)            # float subtraction, then promoted
        M = self.M
        tmp = np.zeros((n, 3))
        for r in range(3):
            tmp[:, r] = (M[:, r, 0] * res[:, 0] + M[:, r, 1] * res[:, 1]) + M[:, r, 2] * res[:, 2]
        k = self.keep
        fterm = np.where(k, (res[:, 0] * tmp[:, 0] + res[:, 1] * tmp[:, 1]) + res[:, 2] * tmp[:, 2], 0.0)
        f = tree_sum(fterm) / float(self.m)
        g = None
        if want_g:
            g = np.zeros(6)
            for c in range(3):
                g[c] = tree_sum(np.where(k, tmp[:, c], 0.0)) * 2.0 / float(self.m)
            pb = xform_f(self.base, self.src).astype(np.float64)
            Rs = np.zeros(9)
            for r in range(3):
                for c in range(3):
                    Rs[3 * r + c] = tree_sum(np.where(k, pb[:, r] * tmp[:, c], 0.0)) * (2.0 / float(self.m))
            g[3:] = r_derivative(x, Rs)
        # (f is a by-product of every evaluation; the oracle's gradient-only call records it too)
        self.trace.append(list(x) + [f if (want_f or want_g) else math.nan, norm(list(g)) if want_g else math.nan])
        return f, g


def r_derivative(x, R):
    """GICP::computeRDerivative: g[3..5] = tr(dR/d(phi, theta, psi) * Rsum), tr(A B) = sum_ij A(j, i) B(i, j)."""
    phi, theta, psi = x[3], x[4], x[5]
    cphi, sphi, cth, sth, cpsi, spsi = math.cos(phi), math.sin(phi), math.cos(theta), math.sin(theta), math.cos(psi), math.sin(psi)
    dphi = [0., sphi * spsi + cphi * cpsi * sth, cphi * spsi - cpsi * sphi * sth,
            0., -cpsi * sphi + cphi * spsi * sth, -cphi * cpsi - sphi * spsi * sth,
            0., cphi * cth, -cth * sphi]
    dth = [-cpsi * sth, cpsi * cth * sphi, cphi * cpsi * cth,
           -spsi * sth, cth * sphi * spsi, cphi * cth * spsi,
           -cth, -sphi * sth, -cphi * sth]
    dpsi = [-cth * spsi, -cphi * cpsi - sphi * spsi * sth, cpsi * sphi - cphi * spsi * sth,
            cpsi * cth, -cphi * spsi + cpsi * sphi * sth, sphi * spsi + cphi * cpsi * sth,
            0., 0., 0.]

    def inner(A):
        r = 0.0
        for i in range(3):
            for j in range(3):
                r += A[3 * j + i] * R[3 * i + j]
        return r
    return np.array([inner(dphi), inner(dth), inner(dpsi)])


# ---------------------------------------------------------------------------------------------------------------
# GSL multimin/vector_bfgs2.c + linear_minimize.c (as PCL's BFGS<FunctorType> ports them), N = 6
# ---------------------------------------------------------------------------------------------------------------
def dot(a, b):
    s = 0.0
    for i in range(len(a)):
        s += a[i] * b[i]
    return s


def norm(a):
    return math.sqrt(dot(a, a))


def solve_quadratic(a, b, c):
    """gsl_poly_solve_quadratic: real roots of a x^2 + b x + c in ascending order."""
    if a == 0:
        return [] if b == 0 else [-c / b]
    disc = b * b - 4 * a * c
    if disc > 0:
        if b == 0:
            r = math.sqrt(-c / a)
            return [-r, r]
        sgnb = 1 if b > 0 else -1
        temp = -0.5 * (b + sgnb * math.sqrt(disc))
        r1, r2 = temp / a, c / temp
        return [r1, r2] if r1 < r2 else [r2, r1]
    if disc == 0:
        return [-0.5 * b / a, -0.5 * b / a]
    return []


def interp_quad(f0, fp0, f1, zl, zh):
    fl = f0 + zl * (fp0 + zl * (f1 - f0 - fp0))
    fh = f0 + zh * (fp0 + zh * (f1 - f0 - fp0))
    c = 2 * (f1 - f0 - fp0)
    zmin, fmin = zl, fl
    if fh < fmin:
        zmin, fmin = zh, fh
    if c > 0:
        z = -fp0 / c
        if zl < z < zh:
            f = f0 + z * (fp0 + z * (f1 - f0 - fp0))
            if f < fmin:
                zmin, fmin = z, f
    return zmin


def interp_cubic(f0, fp0, f1, fp1, zl, zh):
    eta = 3 * (f1 - f0) - 2 * fp0 - fp1
    xi = fp0 + fp1 - 2 * (f1 - f0)
    c0, c1, c2, c3 = f0, fp0, eta, xi

    def cubic(z):
        return c0 + z * (c1 + z * (c2 + z * c3))
    zmin, fmin = zl, cubic(zl)
    for z in [zh] + [r for r in solve_quadratic(3 * c3, 2 * c2, c1) if zl < r < zh]:
        y = cubic(z)
        if y < fmin:
            zmin, fmin = z, y
    return zmin


def interpolate(a, fa, fpa, b, fb, fpb, xmin, xmax, order):
    zmin, zmax = (xmin - a) / (b - a), (xmax - a) / (b - a)
    if zmin > zmax:
        zmin, zmax = zmax, zmin
    if order > 2 and not (math.isnan(fpb) or math.isinf(fpb)):
        z = interp_cubic(fa, fpa * (b - a), fb, fpb * (b - a), zmin, zmax)
    else:
        z = interp_quad(fa, fpa * (b - a), fb, zmin, zmax)
    return a + z * (b - a)


class BFGS:
    """vector_bfgs2 state + its function wrapper (values cached on the step length alpha)."""
    rho, sigma, tau1, tau2, tau3, order, step_size = 0.01, 0.01, 9.0, 0.05, 0.5, 3, 1.0
    bracket_iters = section_iters = 100

    def __init__(self, fn: Functor):
        self.fn = fn

    # -- wrapper
    def moveto(self, alpha):
        if alpha == self.x_key:
            return
        self.x_alpha = [self.x0[i] + alpha * self.p[i] for i in range(6)]
        self.x_key = alpha

    def slope(self):
        return dot(self.g_alpha, self.p)

    def wrap_f(self, alpha):
        if alpha == self.f_key:
            return self.f_alpha
        self.moveto(alpha)
        self.f_alpha, _ = self.fn.fdf(self.x_alpha, True, False)
        self.f_key = alpha
        return self.f_alpha

    def wrap_df(self, alpha):
        if alpha == self.df_key:
            return self.df_alpha
        self.moveto(alpha)
        if alpha != self.g_key:
            _, g = self.fn.fdf(self.x_alpha, False, True)
            self.g_alpha = list(g)
            self.g_key = alpha
        self.df_alpha = self.slope()
        self.df_key = alpha
        return self.df_alpha

    def wrap_fdf(self, alpha):
        if alpha == self.f_key and alpha == self.df_key:
            return self.f_alpha, self.df_alpha
        if alpha == self.f_key or alpha == self.df_key:
            return self.wrap_f(alpha), self.wrap_df(alpha)
        self.moveto(alpha)
        f, g = self.fn.fdf(self.x_alpha, True, True)
        self.f_alpha, self.g_alpha = f, list(g)
        self.f_key = self.g_key = alpha
        self.df_alpha = self.slope()
        self.df_key = alpha
        return self.f_alpha, self.df_alpha

    def change_direction(self):
        self.x_alpha, self.x_key = list(self.x0), 0.0
        self.f_key = 0.0
        self.g_alpha, self.g_key = list(self.g0), 0.0
        self.df_alpha, self.df_key = self.slope(), 0.0

    # -- vector_bfgs2_set
    def init(self, x):
        self.delta_f = 0.0
        self.f, g = self.fn.fdf(x, True, True)
        self.gradient = list(g)
        self.x0, self.g0 = list(x), list(g)
        self.g0norm = norm(self.g0)
        self.p = [gi * (-1.0 / self.g0norm) for gi in self.gradient]
        self.pnorm = norm(self.p)
        self.fp0 = -self.g0norm
        # prepare_wrapper
        self.x_alpha, self.x_key = list(self.x0), 0.0
        self.f_alpha, self.f_key = self.f, 0.0
        self.g_alpha, self.g_key = list(self.g0), 0.0
        self.df_alpha, self.df_key = self.slope(), 0.0

    # -- linear_minimize.c: minimize()
    def line_search(self, alpha1):
        f0, fp0 = self.wrap_fdf(0.0)
        falpha_prev, fpalpha_prev = f0, fp0
        alpha, alpha_prev = alpha1, 0.0
        a, b, fa, fb, fpa, fpb = 0.0, alpha, f0, 0.0, fp0, 0.0
        i = 0
        while i < self.bracket_iters:
            i += 1
            falpha = self.wrap_f(alpha)
            if falpha > f0 + alpha * self.rho * fp0 or falpha >= falpha_prev:
                a, fa, fpa = alpha_prev, falpha_prev, fpalpha_prev
                b, fb, fpb = alpha, falpha, math.nan
                break
            fpalpha = self.wrap_df(alpha)
            if abs(fpalpha) <= -self.sigma * fp0:
                return SUCCESS, alpha
            if fpalpha >= 0:
                a, fa, fpa = alpha, falpha, fpalpha
                b, fb, fpb = alpha_prev, falpha_prev, fpalpha_prev
                break
            delta = alpha - alpha_prev
            alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, alpha + delta,
                                     alpha + self.tau1 * delta, self.order)
            alpha_prev, falpha_prev, fpalpha_prev = alpha, falpha, fpalpha
            alpha = alpha_next
        else:
            i += 1  # `while (i++ < bracket_iters)` leaves i one past the bound when it runs out
        while i < self.section_iters:
            i += 1
            delta = b - a
            alpha = interpolate(a, fa, fpa, b, fb, fpb, a + self.tau2 * delta, b - self.tau3 * delta, self.order)
            falpha = self.wrap_f(alpha)
            if (a - alpha) * fpa <= DBL_EPS:
                return NO_PROGRESS, alpha  # roundoff prevents progress
            if falpha > f0 + self.rho * alpha * fp0 or falpha >= fa:
                b, fb, fpb = alpha, falpha, math.nan
            else:
                fpalpha = self.wrap_df(alpha)
                if abs(fpalpha) <= -self.sigma * fp0:
                    return SUCCESS, alpha
                if ((b - a) >= 0 and fpalpha >= 0) or ((b - a) <= 0 and fpalpha <= 0):
                    b, fb, fpb = a, fa, fpa
                    a, fa, fpa = alpha, falpha, fpalpha
                else:
                    a, fa, fpa = alpha, falpha, fpalpha
        return SUCCESS, 0.0  # section_iters exhausted: GSL returns success with the caller's alpha (0) untouched

    # -- vector_bfgs2_iterate
    def one_step(self, x):
        f0 = self.f
        if self.pnorm == 0.0 or self.g0norm == 0.0 or self.fp0 == 0.0:
            return NO_PROGRESS
        if self.delta_f < 0:
            d = max(-self.delta_f, 10 * DBL_EPS * abs(f0))
            alpha1 = min(1.0, 2.0 * d / (-self.fp0))
        else:
            alpha1 = abs(self.step_size)
        status, alpha = self.line_search(alpha1)
        if status != SUCCESS:
            return status
        # update_position
        self.wrap_fdf(alpha)
        self.f = self.f_alpha
        x[:] = self.x_alpha
        self.gradient = list(self.g_alpha)
        self.delta_f = self.f - f0
        # the (memoryless) BFGS update of the direction
        dx0 = [x[i] - self.x0[i] for i in range(6)]
        dg0 = [self.gradient[i] - self.g0[i] for i in range(6)]
        dxg, dgg, dxdg, dgnorm = dot(dx0, self.gradient), dot(dg0, self.gradient), dot(dx0, dg0), norm(dg0)
        if dxdg != 0:
            B = dxg / dxdg
            A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg
        else:
            A = B = 0.0
        self.p = [self.gradient[i] - A * dx0[i] - B * dg0[i] for i in range(6)]
        self.g0, self.x0 = list(self.gradient), list(x)
        self.g0norm, self.pnorm = norm(self.g0), norm(self.p)
        pg = dot(self.p, self.gradient)
        direction = -1.0 if pg >= 0.0 else 1.0
        self.p = [pi * (direction / self.pnorm) for pi in self.p]
        self.pnorm = norm(self.p)
        self.fp0 = dot(self.p, self.g0)
        self.change_direction()
        return SUCCESS

    def test_gradient(self, eps):
        return SUCCESS if norm(self.gradient) < eps else RUNNING


# ---------------------------------------------------------------------------------------------------------------
# GICP::computeTransformation (guess = identity, as the reference calls it)
# ---------------------------------------------------------------------------------------------------------------
def gicp_align(src, tgt, cov_src, cov_tgt, max_iterations, max_corr_dist=1.0, rotation_epsilon=2e-3,
               transformation_epsilon=1e-6, max_inner=20):
    """Returns dict(T float32 4x4, iterations, converged, n_corr, trace[n, 8])."""
    src, tgt = np.asarray(src, f32), np.asarray(tgt, f32)
    tree = cKDTree(tgt[:, :3].astype(np.float64))
    guess = np.eye(4, dtype=f32)
    T = np.eye(4, dtype=f32)
    prev = T.copy()
    thr = max_corr_dist * max_corr_dist
    iters, converged, n_corr, trace = 0, False, 0, []
    while not converged:
        TR = np.zeros((4, 4))
        for i in range(4):
            for j in range(4):
                s = 0.0
                for k in range(4):
                    s += float(T[i, k]) * float(guess[k, j])
                TR[i, j] = s
        q = xform_f(T, xform_f(guess, src))
        _, nn = tree.query(q.astype(np.float64), k=1)
        d = (q - tgt[nn, :3])
        d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(f32)   # FLANN L2_Simple in float
        keep = d2.astype(np.float64) < thr                                               # strict <
        corr = np.where(keep, nn, -1)
        M = np.tile(np.eye(3), (len(src), 1, 1))
        M[keep] = mahalanobis(TR[:3, :3], cov_src[keep], cov_tgt[nn[keep]])
        n_corr = int(keep.sum())
        prev = T.copy()
        if n_corr < 4:
            break
        x = [float(T[0, 3]), float(T[1, 3]), float(T[2, 3]), math.atan2(float(T[2, 1]), float(T[2, 2])),
             math.asin(-float(T[2, 0])), math.atan2(float(T[1, 0]), float(T[0, 0]))]
        fn = Functor(src, tgt, corr, M, guess)
        bfgs = BFGS(fn)
        bfgs.init(x)
        inner = 0
        while True:
            inner += 1
            result = bfgs.one_step(x)
            if result:
                break
            result = bfgs.test_gradient(1e-2)
            if not (result == RUNNING and inner < max_inner):
                break
        trace += fn.trace
        if not (result in (NO_PROGRESS, SUCCESS) or inner == max_inner):
            break
        T = apply_state(np.eye(4, dtype=f32), x)
        delta = 0.0
        for k in range(4):
            for l in range(4):
                ratio = 1.0 / rotation_epsilon if (k < 3 and l < 3) else 1.0 / transformation_epsilon
                delta = max(delta, ratio * abs(float(f32(prev[k, l] - T[k, l]))))
        iters += 1
        if iters >= max_iterations or delta < 1:
            converged = True
            prev = T.copy()
    return dict(T=prev, iterations=iters, converged=converged, n_corr=n_corr, trace=np.array(trace))


def finite_difference_gradient(fn: Functor, x, h=1e-6):
    g = np.zeros(6)
    for i in range(6):
        xp, xm = list(x), list(x)
        xp[i] += h
        xm[i] -= h
        g[i] = (fn.fdf(xp, True, False)[0] - fn.fdf(xm, True, False)[0]) / (2 * h)
    return g
