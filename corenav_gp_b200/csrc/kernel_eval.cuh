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

__device__ __forceinline__ double fast_exp(double x);   // branch-free 16-operation exp for x <= ~1, defined below

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
    case CNGP_K_RBF: return c.c0 * fast_exp(r2 * c.c1);
    case CNGP_K_MAT32: {
      const double r = sqrt(r2) * c.c1;
      return c.c0 * (1.0 + r) * fast_exp(-r);
    }
    case CNGP_K_MAT52: {
      const double r = sqrt(r2) * c.c1;
      return c.c0 * (1.0 + r + r * r * (1.0 / 3.0)) * fast_exp(-r);
    }
    case CNGP_K_RATQUAD: return c.c0 * fast_exp(-c.c2 * log1p(r2 * c.c1));
    case CNGP_K_STDPERIODIC: {
      const double s = sin((xa - xb) * c.c1);
      return c.c0 * fast_exp(s * s * c.c2);
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

// The same with the reciprocals of the length scales taken once per problem (GradConst) instead of once per entry,
// and the branch-free 16-operation exp: gp_grad_kernel evaluates N(N+1)/2 entries per problem.
struct GradConst {
  double i1, i1sq, i2;   // 1/th[1], 1/th[1]^2, 1/th[2] (or 1/th[2]^2 for the periodic leaf's i1sq)
};
__device__ __forceinline__ GradConst grad_prepare(int type, const double* th) {
  GradConst c{0.0, 0.0, 0.0};
  switch (type) {
    case CNGP_K_RBF: case CNGP_K_MAT32: case CNGP_K_MAT52: case CNGP_K_RATQUAD:
      c.i1 = 1.0 / th[1]; c.i1sq = 1.0 / (th[1] * th[1]); break;
    case CNGP_K_STDPERIODIC: c.i1 = 1.0 / th[1]; c.i1sq = 1.0 / (th[2] * th[2]); c.i2 = 1.0 / th[2]; break;
    default: break;
  }
  return c;
}
template <bool SYM>
__device__ __forceinline__ double leaf_value_grad_c(int type, const double* th, const GradConst& c, double xa, double xb,
                                                    double r2, bool same, double dv[3]) {
  dv[0] = dv[1] = dv[2] = 0.0;
  switch (type) {
    case CNGP_K_RBF: {
      const double e = fast_exp(-0.5 * r2 * c.i1sq);
      dv[0] = e;
      dv[1] = th[0] * e * r2 * c.i1sq * c.i1;
      return th[0] * e;
    }
    case CNGP_K_MAT32: {
      const double r = sqrt(r2) * (1.7320508075688772 * c.i1);
      const double e = fast_exp(-r);
      dv[0] = (1.0 + r) * e;
      dv[1] = th[0] * r * r * e * c.i1;
      return th[0] * dv[0];
    }
    case CNGP_K_MAT52: {
      const double r = sqrt(r2) * (2.2360679774997897 * c.i1);
      const double e = fast_exp(-r);
      dv[0] = (1.0 + r + r * r * (1.0 / 3.0)) * e;
      dv[1] = th[0] * (r * r * (1.0 + r) * (1.0 / 3.0)) * e * c.i1;
      return th[0] * dv[0];
    }
    case CNGP_K_RATQUAD: {
      const double h = 0.5 * r2 * c.i1sq;
      const double l1p = log1p(h);
      const double kr = fast_exp(-th[2] * l1p);
      dv[0] = kr;
      dv[1] = th[0] * kr * th[2] * (2.0 * h) * c.i1 / (1.0 + h);
      dv[2] = -th[0] * kr * l1p;
      return th[0] * kr;
    }
    case CNGP_K_STDPERIODIC: {
      const double base = 3.14159265358979323846 * (xa - xb) * c.i1;
      double s, co;
      sincos(base, &s, &co);
      const double e = fast_exp(-0.5 * s * s * c.i1sq);
      dv[0] = e;
      dv[1] = th[0] * e * c.i1sq * s * co * (base * c.i1);
      dv[2] = th[0] * e * s * s * c.i1sq * c.i2;
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

// ------------------------------------------------------------------------------------------------------------
// Specialised evaluators.  The composite kernels the reference actually deploys or benchmarks get a straight-line
// evaluator built from per-point features computed once per point (x, x^2 and, for the periodic leaf, cos/sin of the
// phase 2 pi x / p so that  sin^2(pi (x-x')/p) = (1 - cos(phi - phi'))/2  needs two fma instead of a sin), and an
// exp that is 16 FP64 operations (Cody-Waite reduction + degree-12 polynomial, ~1 ulp) - the kernel matrix is
// (N^2/2 + N M) evaluations per window, the same order of work as the factorisation.
//   KID 0  generic sum-of-products interpreter (any expression)
//   KID 1  rbf                       (BASELINE.json configs[0])
//   KID 2  rbf + stdperiodic         (configs[1..2], "SE + periodic")
//   KID 3  rbf * brownian            (the deployed kernel, gp_slip_node.py:31)
// ------------------------------------------------------------------------------------------------------------
constexpr int KID_GENERIC = 0, KID_RBF = 1, KID_RBF_PER = 2, KID_RBF_BROWN = 3;
constexpr int KID_TILES = 4;   // gp_fit_kernel only: factor a block held as tiles in global memory (chol_large.cu)

// exp(x) for x <= ~1 (kernel exponents are never positive); flushes to 0 below -708.
__device__ __forceinline__ double fast_exp(double x) {
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest integer in the low word
  const double t = fma(x, 1.4426950408889634, magic);
  const int n = __double2loint(t);
  const double nd = t - magic;
  double f = fma(nd, -6.93147180369123816490e-01, x);
  f = fma(nd, -1.90821492927058770002e-10, f);
  double p = 2.08767569878680989792e-09;           // 1/12!
  p = fma(p, f, 2.50521083854417187751e-08);       // 1/11!
  p = fma(p, f, 2.75573192239858906526e-07);       // 1/10!
  p = fma(p, f, 2.75573192239858906526e-06);       // 1/9!
  p = fma(p, f, 2.48015873015873015873e-05);       // 1/8!
  p = fma(p, f, 1.98412698412698412698e-04);       // 1/7!
  p = fma(p, f, 1.38888888888888888889e-03);       // 1/6!
  p = fma(p, f, 8.33333333333333333333e-03);       // 1/5!
  p = fma(p, f, 4.16666666666666666667e-02);       // 1/4!
  p = fma(p, f, 1.66666666666666666667e-01);       // 1/3!
  p = fma(p, f, 0.5);
  p = fma(p, f, 1.0);
  p = fma(p, f, 1.0);
  // scale by 2^n through the exponent field; below 2^-1021 flush to zero (integer compare: keeps the FP64 pipe free)
  const int hi = (n < -1021) ? 0 : __double2hiint(p) + (n << 20);
  const int lo = (n < -1021) ? 0 : __double2loint(p);
  return __hiloint2double(hi, lo);
}

// Table-driven exp for the assembly hot loops: exp(x) = 2^e * T[j] * (1 + r h(r)),  64 e + j = round(64 x / ln 2),
// |r| <= ln2/128, h = degree-4 Taylor of (e^r - 1)/r (truncation 3.5e-17).  Nine FP64 operations instead of sixteen;
// the table (2^(j/64), correctly rounded) is read from shared memory where the caller has pre-multiplied it by the
// leaf variance, which also removes the final multiply.  The reduction uses a single constant: its absolute error is
// <= 1.1e-16 |x| e^x <= 4e-17, below the rounding of the near-diagonal entries.  Valid for x <= ~1; x < -600 -> 0.
__device__ const double EXP2_TAB64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

__device__ __forceinline__ double exp_tab(double x, const double* tab) {
  const double magic = 6755399441055744.0;                     // 1.5 * 2^52
  const double t = fma(x, 92.33248261689366, magic);           // 64 / ln 2
  const int n = __double2loint(t);
  const double nd = t - magic;
  const double r = fma(nd, -0.010830424696249145, x);          // ln 2 / 64
  double h = fma(r, 8.33333333333333333e-03, 4.16666666666666667e-02);
  h = fma(h, r, 1.66666666666666667e-01);
  h = fma(h, r, 0.5);
  h = fma(h, r, 1.0);
  const double T = tab[n & 63];
  const double res = fma(T * r, h, T);
  // scale by 2^(n >> 6) through the exponent field, flush to zero for x < -600.  The selects are opaque inline asm
  // on purpose: written as C++ conditionals the compiler turns them into a BRANCH around the polynomial, which makes
  // every exp its own basic block and stops the scheduler from interleaving the evaluations with the DMMA stream.
  int hi = __double2hiint(res) + ((n >> 6) << 20), lo = __double2loint(res);
  asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %2, 0xC082C000;\n\tselp.b32 %0, 0, %0, p;\n\tselp.b32 %1, 0, %1, p;\n\t}"
      : "+r"(hi), "+r"(lo)
      : "r"(__double2hiint(x)));
  return __hiloint2double(hi, lo);
}

struct PointFeat {
  double x, xx, c, s;  // x, x*x, cos(2 pi x / p), sin(2 pi x / p)
};

template <int KID>
struct FastK {
  double c0, c1, c2, c3, c4;
  // Phase origin of the periodic leaf's features.  The covariance only sees differences of phases, so any origin gives
  // the same value in exact arithmetic; taking it at the window's first stamp keeps the phases O(span) - with the raw
  // stamp at |x| ~ 1e6 the product x (2 pi / p) alone is rounded to 2e-11 rad, which costs 3e-8 on the predictive mean
  // (GPy forms the difference x - x' first and does not have this problem).  x - base is exact for nearby stamps.
  double base = 0.0;
  __device__ __forceinline__ void init(const double* th) {
    c0 = th[0];
    c1 = -0.5 / (th[1] * th[1]);
    if (KID == KID_RBF_PER) {
      c2 = 2.0 * 3.14159265358979323846 / th[3];   // phase scale 2 pi / period
      c3 = 0.25 / (th[4] * th[4]);                 // A = 1 / (4 l^2): -0.5 sin^2 / l^2 = A (cos(dphi) - 1)
      c4 = th[2];
    } else if (KID == KID_RBF_BROWN) {
      c4 = th[2];
    }
  }
  __device__ __forceinline__ PointFeat point(double x) const {
    PointFeat f{x, __dmul_rn(x, x), 0.0, 0.0};
    if (KID == KID_RBF_PER) sincos((x - base) * c2, &f.s, &f.c);
    return f;
  }
  // GPy expanded-form r^2 from the features (same roundings as r2_expanded; -2 m is exact so the fma is too)
  // ... and clipped at 0 as GPy does (np.clip(r2, 0, inf)): at |x| ~ 1e6 the expanded form is off by ~1e-4 in either
  // direction, and a NEGATIVE r2 under a short length scale would give exp(+1e-4 / l^2) instead of 1 - far outside the
  // 1e-9 parity (tests/test_gpu_parity_report.py::test_r2_clip_with_large_stamps).
  __device__ __forceinline__ double r2(const PointFeat& a, const PointFeat& b) const {
    return fmax(fma(-2.0, __dmul_rn(a.x, b.x), __dadd_rn(a.xx, b.xx)), 0.0);
  }
  __device__ __forceinline__ double eval(const PointFeat& a, const PointFeat& b, bool same_sym) const {
    double rr = r2(a, b);
    if (same_sym) rr = 0.0;
    const double e1 = c0 * fast_exp(rr * c1);
    if (KID == KID_RBF) return e1;
    if (KID == KID_RBF_PER) {
      const double cd = fma(a.c, b.c, a.s * b.s);          // cos(phi_a - phi_b)
      return fma(c4, fast_exp(fma(c3, cd, -c3)), e1);
    }
    // rbf * brownian
    const bool agree = (a.x > 0.0 && b.x > 0.0) || (a.x < 0.0 && b.x < 0.0) || (a.x == 0.0 && b.x == 0.0);
    return agree ? e1 * (c4 * fmin(fabs(a.x), fabs(b.x))) : 0.0;
  }
  // Table variant (exp_tab): tab[0..63] = scale1() 2^(j/64), tab[64..127] = scale2() 2^(j/64), built by the caller.
  __device__ __forceinline__ double scale1() const { return KID == KID_RBF_BROWN ? c0 * c4 : c0; }
  __device__ __forceinline__ double scale2() const { return KID == KID_RBF_PER ? c4 : 0.0; }
  __device__ __forceinline__ double eval_tab(const PointFeat& a, const PointFeat& b, bool same_sym, const double* tab) const {
    double rr = r2(a, b);
    if (same_sym) rr = 0.0;
    const double e1 = exp_tab(rr * c1, tab);
    if (KID == KID_RBF) return e1;
    if (KID == KID_RBF_PER) {
      const double cd = fma(a.c, b.c, a.s * b.s);          // cos(phi_a - phi_b)
      return e1 + exp_tab(fma(c3, cd, -c3), tab + 64);
    }
    const bool agree = (a.x > 0.0 && b.x > 0.0) || (a.x < 0.0 && b.x < 0.0) || (a.x == 0.0 && b.x == 0.0);
    return agree ? e1 * fmin(fabs(a.x), fabs(b.x)) : 0.0;
  }
  __device__ __forceinline__ double kdiag(double x) const {
    if (KID == KID_RBF) return c0;
    if (KID == KID_RBF_PER) return c0 + c4;
    return c0 * (c4 * fabs(x));
  }
};

// Uniform front end used by the kernels: KID 0 forwards to the interpreter (kept out of line so the unrolled
// callers stay small), KID > 0 to FastK.
static __device__ __noinline__ double keval_generic_sym(const KProg* kp, const LeafConst* hc, double xa, double xb, bool same) {
  return keval<true>(*kp, hc, xa, xb, same);
}
static __device__ __noinline__ double keval_generic_cross(const KProg* kp, const LeafConst* hc, double xa, double xb) {
  return keval<false>(*kp, hc, xa, xb, false);
}

// which specialised evaluator (if any) matches a sum-of-products program
__host__ inline int match_fast_kernel(const KProg& kp) {
  if (kp.n_terms == 1 && kp.n_leaves == 1 && kp.leaf_type[0] == CNGP_K_RBF) return KID_RBF;
  if (kp.n_terms == 2 && kp.n_leaves == 2 && kp.leaf_type[0] == CNGP_K_RBF && kp.leaf_type[1] == CNGP_K_STDPERIODIC &&
      kp.leaf_param[0] == 0 && kp.leaf_param[1] == 2)
    return KID_RBF_PER;
  if (kp.n_terms == 1 && kp.n_leaves == 2 && kp.leaf_type[0] == CNGP_K_RBF && kp.leaf_type[1] == CNGP_K_BROWNIAN &&
      kp.leaf_param[0] == 0 && kp.leaf_param[1] == 2)
    return KID_RBF_BROWN;
  return KID_GENERIC;
}

}  // namespace cngp
