// Device evaluation of composite covariance functions (GPy forms).
//
// Replaces GPy kern.K / kern.Kdiag / kern.update_gradients_full as reached from
// core_navigation/script/gp_slip_node.py:31,35-36,48 (kernel = RBF(1) * Brownian(1) and the candidate families of
// gp_slip_node.py:32-34 and "Kernel Selection/README.md":18-20).
//
// Fidelity notes (restated from GPy 1.9.x, not copied):
//  * stationary kernels use GPy's EXPANDED squared distance  r2 = -2 x x' + (x^2 + x'^2), evaluated with the same
//    individually rounded operations (no FMA contraction), forced to 0 on the diagonal of K(X,X) and clipped at 0;
//    at large |x| this differs from (x-x')^2 by ~1e-16 x^2, which matters for parity at N = 32768.
//  * StdPeriodic uses the direct difference pi (x - x') / period.
//  * Brownian: variance * min(|x|,|x'|) when signs agree else 0;  Kdiag = variance |x|.
//  * White contributes only to K(X,X) (diagonal) and to Kdiag.
#pragma once
#include "cngp_common.cuh"

namespace cngp {

struct LeafConst {
  double c0, c1, c2;
};

// Derived per-window constants of one leaf from its hyper-parameters.
__device__ __forceinline__ LeafConst leaf_prepare(int type, const double* th) {
  LeafConst c{th[0], 0.0, 0.0};
  switch (type) {
    case CNGP_K_RBF: c.c1 = -0.5 / (th[1] * th[1]); break;
    case CNGP_K_MAT32: c.c1 = 1.7320508075688772 / th[1]; break;
    case CNGP_K_MAT52: c.c1 = 2.2360679774997897 / th[1]; break;
    case CNGP_K_RATQUAD: c.c1 = 0.5 / (th[1] * th[1]); c.c2 = th[2]; break;
    case CNGP_K_STDPERIODIC: c.c1 = 3.14159265358979323846 / th[1]; c.c2 = -0.5 / (th[2] * th[2]); break;
    default: break;
  }
  return c;
}

// GPy Stationary._unscaled_dist squared (expanded form, individually rounded).
__device__ __forceinline__ double r2_expanded(double xa, double xb) {
  const double m2t = __dmul_rn(-2.0, __dmul_rn(xa, xb));
  const double s = __dadd_rn(__dmul_rn(xa, xa), __dmul_rn(xb, xb));
  return fmax(__dadd_rn(m2t, s), 0.0);
}

template <bool SYM>
__device__ __forceinline__ double leaf_value(int type, const LeafConst& c, double xa, double xb, double r2, bool same) {
  switch (type) {
    case CNGP_K_RBF: return c.c0 * exp(r2 * c.c1);
    case CNGP_K_MAT32: {
      const double r = sqrt(r2) * c.c1;
      return c.c0 * (1.0 + r) * exp(-r);
    }
    case CNGP_K_MAT52: {
      const double r = sqrt(r2) * c.c1;
      return c.c0 * (1.0 + r + r * r * (1.0 / 3.0)) * exp(-r);
    }
    case CNGP_K_RATQUAD: return c.c0 * exp(-c.c2 * log1p(r2 * c.c1));
    case CNGP_K_STDPERIODIC: {
      const double s = sin((xa - xb) * c.c1);
      return c.c0 * exp(s * s * c.c2);
    }
    case CNGP_K_BROWNIAN: {
      const bool agree = (xa > 0.0 && xb > 0.0) || (xa < 0.0 && xb < 0.0) || (xa == 0.0 && xb == 0.0);
      return agree ? c.c0 * fmin(fabs(xa), fabs(xb)) : 0.0;
    }
    case CNGP_K_LINEAR: return c.c0 * (xa * xb);
    case CNGP_K_BIAS: return c.c0;
    case CNGP_K_WHITE: return (SYM && same) ? c.c0 : 0.0;
  }
  return 0.0;
}

// k(xa, xb).  SYM = true: entry (ia, ib) of K(X,X) (diagonal r2 forced to 0, White on the diagonal);
// SYM = false: entry of the cross-covariance K(X, X*).
template <bool SYM>
__device__ __forceinline__ double keval(const KProg& kp, const LeafConst* hc, double xa, double xb, bool same) {
  double r2 = r2_expanded(xa, xb);
  if (SYM && same) r2 = 0.0;
  double acc = 0.0;
  for (int t = 0; t < kp.n_terms; ++t) {
    double prod = 1.0;
    for (int u = kp.term_start[t]; u < kp.term_start[t + 1]; ++u)
      prod *= leaf_value<SYM>(kp.leaf_type[u], hc[u], xa, xb, r2, same);
    acc += prod;
  }
  return acc;
}

// Kdiag(x*) of the composite (GPy Kdiag: stationary/periodic/bias/white -> variance, Brownian -> variance |x|,
// Linear -> variance x^2).
__device__ __forceinline__ double kdiag_eval(const KProg& kp, const LeafConst* hc, double x) {
  double acc = 0.0;
  for (int t = 0; t < kp.n_terms; ++t) {
    double prod = 1.0;
    for (int u = kp.term_start[t]; u < kp.term_start[t + 1]; ++u) {
      const int type = kp.leaf_type[u];
      double v = hc[u].c0;
      if (type == CNGP_K_BROWNIAN) v *= fabs(x);
      else if (type == CNGP_K_LINEAR) v *= x * x;
      prod *= v;
    }
    acc += prod;
  }
  return acc;
}

// Leaf value and its derivatives with respect to the leaf's own hyper-parameters (GPy update_gradients_full
// integrands).  Returns the value; dv[j] = d leaf / d theta_j (j < number of leaf parameters).
template <bool SYM>
__device__ __forceinline__ double leaf_value_grad(int type, const double* th, double xa, double xb, double r2,
                                                  bool same, double dv[3]) {
  dv[0] = dv[1] = dv[2] = 0.0;
  switch (type) {
    case CNGP_K_RBF: {
      const double il2 = 1.0 / (th[1] * th[1]);
      const double e = exp(-0.5 * r2 * il2);
      dv[0] = e;
      dv[1] = th[0] * e * r2 * il2 / th[1];
      return th[0] * e;
    }
    case CNGP_K_MAT32: {
      const double r = sqrt(r2) * (1.7320508075688772 / th[1]);  // sqrt(3) r
      const double e = exp(-r);
      dv[0] = (1.0 + r) * e;
      dv[1] = th[0] * r * r * e / th[1];  // -dk/dr * r / l with dk/dr = -3 r e^{-sqrt3 r}
      return th[0] * dv[0];
    }
    case CNGP_K_MAT52: {
      const double r = sqrt(r2) * (2.2360679774997897 / th[1]);  // sqrt(5) r
      const double e = exp(-r);
      dv[0] = (1.0 + r + r * r * (1.0 / 3.0)) * e;
      dv[1] = th[0] * (r * r * (1.0 + r) * (1.0 / 3.0)) * e / th[1];
      return th[0] * dv[0];
    }
    case CNGP_K_RATQUAD: {
      const double il2 = 1.0 / (th[1] * th[1]);
      const double h = 0.5 * r2 * il2;          // r^2/2
      const double l1p = log1p(h);
      const double kr = exp(-th[2] * l1p);
      dv[0] = kr;
      dv[1] = th[0] * kr * th[2] * (2.0 * h) / ((1.0 + h) * th[1]);
      dv[2] = -th[0] * kr * l1p;
      return th[0] * kr;
    }
    case CNGP_K_STDPERIODIC: {
      const double base = 3.14159265358979323846 * (xa - xb) / th[1];
      double s, co;
      sincos(base, &s, &co);
      const double il2 = 1.0 / (th[2] * th[2]);
      const double e = exp(-0.5 * s * s * il2);
      dv[0] = e;
      dv[1] = th[0] * e * il2 * s * co * (base / th[1]);
      dv[2] = th[0] * e * s * s * il2 / th[2];
      return th[0] * e;
    }
    case CNGP_K_BROWNIAN: {
      const bool agree = (xa > 0.0 && xb > 0.0) || (xa < 0.0 && xb < 0.0) || (xa == 0.0 && xb == 0.0);
      dv[0] = agree ? fmin(fabs(xa), fabs(xb)) : 0.0;
      return th[0] * dv[0];
    }
    case CNGP_K_LINEAR: dv[0] = xa * xb; return th[0] * dv[0];
    case CNGP_K_BIAS: dv[0] = 1.0; return th[0];
    case CNGP_K_WHITE: dv[0] = (SYM && same) ? 1.0 : 0.0; return th[0] * dv[0];
  }
  return 0.0;
}

__host__ __device__ __forceinline__ int leaf_nparams(int type) {
  return (type == CNGP_K_RATQUAD || type == CNGP_K_STDPERIODIC) ? 3 : (type >= CNGP_K_BROWNIAN ? 1 : 2);
}

}  // namespace cngp
