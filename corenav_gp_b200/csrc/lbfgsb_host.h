// L-BFGS-B (no bounds) as a resumable per-window state machine; see gp_slip.cu for what it restates.
// Plain C++ with fixed-size state and no library containers, usable on the host AND in device code: gp_slip.cu runs one
// state machine per window in a device kernel (opt_feed_kernel - no host round trip per iteration), the CPU test shim
// tests/host/lbfgsb_shim.cpp drives the very same code with a caller-supplied objective.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define CNGP_HD __host__ __device__
#else
#define CNGP_HD
#endif

namespace cngp_host {

constexpr int kMaxP = 25;                         // CNGP_MAX_PARAMS + 1 (kernel hyper-parameters + noise)
CNGP_HD inline double hd_max(double a, double b) { return a > b ? a : b; }
CNGP_HD inline double hd_min(double a, double b) { return a < b ? a : b; }

constexpr double kEps = 2.220446049250313e-16;   // dpmeps
constexpr double kFactr = 1e7, kPgtol = 1e-5;    // scipy fmin_l_bfgs_b defaults (paramz passes none)
constexpr int kHist = 10;                         // m
constexpr double kFtol = 1e-3, kGtol = 0.9, kXtol = 0.1, kStpMax = 1e10;
constexpr double kLim = 36.0;                     // paramz transformations._lim_val

CNGP_HD inline double softplus(double z) { return z > kLim ? z : log1p(exp(z)); }
CNGP_HD inline double softplus_inv(double t) { return t > kLim ? t : log(expm1(t)); }
CNGP_HD inline double softplus_gradfactor(double t) { return t > kLim ? 1.0 : -expm1(-t); }

// MINPACK-2 dcstep: safeguarded cubic/quadratic step of the More-Thuente search.
CNGP_HD inline void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp, double fp, double dp,
            bool& brackt, double stpmin, double stpmax) {
  const double sgnd = dp * (dx / fabs(dx));
  double stpf, stpc, stpq, theta, s, gamma, p, q, r;
  if (fp > fx) {                       // case 1: higher function value - the minimum is bracketed
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = hd_max(fabs(theta), hd_max(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp < stx) gamma = -gamma;
    p = (gamma - dx) + theta;
    q = ((gamma - dx) + gamma) + dp;
    r = p / q;
    stpc = stx + r * (stp - stx);
    stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
    stpf = fabs(stpc - stx) < fabs(stpq - stx) ? stpc : stpc + (stpq - stpc) / 2.0;
    brackt = true;
  } else if (sgnd < 0.0) {             // case 2: lower value, derivatives of opposite sign - bracketed
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = hd_max(fabs(theta), hd_max(fabs(dx), fabs(dp)));
    gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + dx;
    r = p / q;
    stpc = stp + r * (stx - stp);
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
    brackt = true;
  } else if (fabs(dp) < fabs(dx)) {    // case 3: lower value, same sign, derivative magnitude decreases
    theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
    s = hd_max(fabs(theta), hd_max(fabs(dx), fabs(dp)));
    gamma = s * sqrt(hd_max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
    if (stp > stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
    else stpc = stp > stx ? stpmax : stpmin;
    stpq = stp + (dp / (dp - dx)) * (stx - stp);
    if (brackt) {
      stpf = fabs(stpc - stp) < fabs(stpq - stp) ? stpc : stpq;
      stpf = stp > stx ? hd_min(stp + 0.66 * (sty - stp), stpf) : hd_max(stp + 0.66 * (sty - stp), stpf);
    } else {
      stpf = fabs(stpc - stp) > fabs(stpq - stp) ? stpc : stpq;
      stpf = hd_min(stpmax, stpf);
      stpf = hd_max(stpmin, stpf);
    }
  } else {                             // case 4: lower value, same sign, derivative does not decrease
    if (brackt) {
      theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
      s = hd_max(fabs(theta), hd_max(fabs(dy), fabs(dp)));
      gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
      if (stp > sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + dy;
      r = p / q;
      stpf = stp + r * (sty - stp);
    } else {
      stpf = stp > stx ? stpmax : stpmin;
    }
  }
  if (fp > fx) {
    sty = stp; fy = fp; dy = dp;
  } else {
    if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
    stx = stp; fx = fp; dx = dp;
  }
  stp = stpf;
}

// MINPACK-2 dcsrch as a resumable object: start() sets the first trial step, step(f, g) consumes the value and the
// directional derivative at the current trial step and returns true when the search has ended (converged or warned).
struct LineSearch {
  bool brackt;
  int stage;
  double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1, stp;
  CNGP_HD void start(double f0, double g0, double stp0) {
    brackt = false; stage = 1; finit = f0; ginit = g0; gtest = kFtol * ginit;
    width = kStpMax; width1 = 2.0 * width;
    stx = 0.0; fx = finit; gx = ginit; sty = 0.0; fy = finit; gy = ginit;
    stmin = 0.0; stp = stp0; stmax = stp + 4.0 * stp;
  }
  CNGP_HD bool step(double f, double g) {
    const double ftest = finit + stp * gtest;
    if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
    if (brackt && (stp <= stmin || stp >= stmax)) return true;          // rounding errors prevent progress
    if (brackt && stmax - stmin <= kXtol * stmax) return true;          // xtol test satisfied
    if (stp == kStpMax && f <= ftest && g <= gtest) return true;
    if (stp == 0.0 && (f > ftest || g >= gtest)) return true;
    if (f <= ftest && fabs(g) <= kGtol * (-ginit)) return true;         // strong Wolfe conditions hold
    if (stage == 1 && f <= fx && f > ftest) {
      double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
      double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
      dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, stmin, stmax);
      fx = fxm + stx * gtest; fy = fym + sty * gtest; gx = gxm + gtest; gy = gym + gtest;
    } else {
      dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, brackt, stmin, stmax);
    }
    if (brackt) {
      if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
      width1 = width;
      width = fabs(sty - stx);
    }
    if (brackt) {
      stmin = hd_min(stx, sty); stmax = hd_max(stx, sty);
    } else {
      stmin = stp + 1.1 * (stp - stx); stmax = stp + 4.0 * (stp - stx);
    }
    stp = hd_max(stp, 0.0);
    stp = hd_min(stp, kStpMax);
    if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= kXtol * stmax)) stp = stx;
    return false;
  }
};

// One window's L-BFGS-B (no bounds) run, driven from outside: trial() is the point to evaluate next, feed(f, g)
// consumes the objective and gradient there.
struct Optimizer {
  int P, max_iters;
  double z[kMaxP], g[kMaxP], d[kMaxP], zt[kMaxP], s_hist[kHist * kMaxP], y_hist[kHist * kMaxP], rho[kHist];
  double f, gd0, dnorm;
  int hist_n, hist_head, iters, nfev, ls_evals;
  bool done, first;
  LineSearch ls;

  CNGP_HD void init(const double* theta0, int P_, int max_iters_) {
    P = P_; max_iters = max_iters_;
    f = gd0 = dnorm = 0.0;
    hist_n = hist_head = iters = nfev = ls_evals = 0;
    done = false; first = true;
    for (int i = 0; i < kHist * kMaxP; ++i) s_hist[i] = y_hist[i] = 0.0;
    for (int i = 0; i < kHist; ++i) rho[i] = 0.0;
    for (int i = 0; i < kMaxP; ++i) z[i] = g[i] = d[i] = zt[i] = 0.0;
    for (int i = 0; i < P; ++i) zt[i] = z[i] = softplus_inv(theta0[i]);
  }
  CNGP_HD double gmax(const double* v) const {
    double m = 0.0;
    for (int i = 0; i < P; ++i) m = hd_max(m, fabs(v[i]));
    return m;
  }
  // d = -H g by the two-loop recursion (H0 = s'y / y'y of the newest pair)
  CNGP_HD void direction() {
    double qv[kMaxP], al[kHist];
    for (int i = 0; i < P; ++i) qv[i] = g[i];
    for (int t = 0; t < hist_n; ++t) {
      const int k = (hist_head - 1 - t + 2 * kHist) % kHist;
      double a = 0.0;
      for (int i = 0; i < P; ++i) a += s_hist[k * P + i] * qv[i];
      a *= rho[k];
      al[k] = a;
      for (int i = 0; i < P; ++i) qv[i] -= a * y_hist[k * P + i];
    }
    if (hist_n > 0) {
      const int k = (hist_head - 1 + kHist) % kHist;
      double yy = 0.0;
      for (int i = 0; i < P; ++i) yy += y_hist[k * P + i] * y_hist[k * P + i];
      const double gam = 1.0 / (rho[k] * yy);
      for (int i = 0; i < P; ++i) qv[i] *= gam;
    }
    for (int t = hist_n - 1; t >= 0; --t) {
      const int k = (hist_head - 1 - t + 2 * kHist) % kHist;
      double b = 0.0;
      for (int i = 0; i < P; ++i) b += y_hist[k * P + i] * qv[i];
      b *= rho[k];
      for (int i = 0; i < P; ++i) qv[i] += (al[k] - b) * s_hist[k * P + i];
    }
    for (int i = 0; i < P; ++i) d[i] = -qv[i];
  }
  // set up the line search from the current iterate; returns false when no descent direction can be found
  CNGP_HD bool begin_search() {
    for (int attempt = 0; attempt < 2; ++attempt) {
      direction();
      gd0 = 0.0; dnorm = 0.0;
      for (int i = 0; i < P; ++i) { gd0 += g[i] * d[i]; dnorm += d[i] * d[i]; }
      dnorm = sqrt(dnorm);
      if (gd0 < 0.0) break;
      if (hist_n == 0) return false;          // steepest descent is not a descent direction: g == 0
      hist_n = 0;                             // refresh the memory and retry (lnsrlb info = -4)
    }
    const double stp0 = (iters == 0) ? hd_min(1.0 / dnorm, kStpMax) : 1.0;
    ls.start(f, gd0, stp0);
    ls_evals = 0;
    for (int i = 0; i < P; ++i) zt[i] = z[i] + ls.stp * d[i];
    return true;
  }
  CNGP_HD void feed(double ft, const double* gt) {
    ++nfev;
    if (first) {
      first = false;
      f = ft;
      for (int i = 0; i < P; ++i) g[i] = gt[i];
      if (gmax(g) <= kPgtol || !begin_search()) done = true;
      return;
    }
    ++ls_evals;
    double gdt = 0.0;
    for (int i = 0; i < P; ++i) gdt += gt[i] * d[i];
    const double stp_eval = ls.stp;
    const bool finished = ls.step(ft, gdt);
    if (!finished) {
      if (ls_evals >= 20 || nfev >= max_iters) {
        // line search gave up: restart from the last iterate without memory, or stop (lnsrlb info != 0)
        if (nfev >= max_iters || hist_n == 0) { done = true; return; }
        hist_n = 0;
        if (!begin_search()) done = true;
        return;
      }
      for (int i = 0; i < P; ++i) zt[i] = z[i] + ls.stp * d[i];
      return;
    }
    // accept z + stp d: update the memory with s = stp d, y = g_new - g
    const double fold = f;
    const double dr = (gdt - gd0) * stp_eval, ddum = -gd0 * stp_eval;
    if (dr > kEps * ddum) {
      const int k = hist_head;
      for (int i = 0; i < P; ++i) {
        s_hist[k * P + i] = stp_eval * d[i];
        y_hist[k * P + i] = gt[i] - g[i];
      }
      rho[k] = 1.0 / dr;
      hist_head = (hist_head + 1) % kHist;
      hist_n = hd_min(hist_n + 1, kHist);
    }
    for (int i = 0; i < P; ++i) { z[i] += stp_eval * d[i]; g[i] = gt[i]; }
    f = ft;
    ++iters;
    if (gmax(g) <= kPgtol) { done = true; return; }
    if (fold - f <= kEps * kFactr * hd_max(hd_max(fabs(fold), fabs(f)), 1.0)) { done = true; return; }
    if (iters >= max_iters || nfev >= max_iters) { done = true; return; }
    if (!begin_search()) done = true;
  }
};

}  // namespace cngp_host
