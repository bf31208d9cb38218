// Phase A of the batched exact GP: assemble Ky = K(X,X) + (noise + 1e-8) I, factor it, solve z = L^-1 y, LML.
//
// Replaces GPy ExactGaussianInference.inference as reached from GPRegression(...) at
// core_navigation/script/gp_slip_node.py:35 (row a3): kern.K, pdinv -> jitchol -> dpotrf, dpotrs, logdet.
//
// One CTA per (candidate, window) problem.  Left-looking tile Cholesky over 8-wide tile columns; every update,
// the triangular solve below the diagonal (through the explicitly inverted 8x8 diagonal tile) and the forward solve
// for z (carried as an extra 1-row "tile row" under the matrix) are tile_mma = FP64 DMMA.  The kernel matrix is
// never materialised: each 8x8 tile of Ky is evaluated in registers at the moment its column becomes current.
// Output factor layout (consumed by gp_var.cuh / gp_grad.cuh): column-block-major tiles, tile (j,j) holds
// inv(L_jj) (lower triangular), tiles (i>j, j) hold L_ij.
#pragma once
#include "kernel_eval.cuh"

namespace cngp {

constexpr int FIT_WARPS = 8;
constexpr int FIT_THREADS = FIT_WARPS * 32;
constexpr int FIT_MAXT = 5;  // ceil((32 + 1) / 8) row tiles per warp in the first column
constexpr int NPAD = CNGP_MAX_N + 8;

struct FitArgs {
  KProg kp;
  const double* theta;   // hyper-parameters, noise last
  long long theta_stride;
  int theta_mode;        // 0 shared, 1 per window (problem % n_windows), 2 per candidate (problem / n_windows)
  const double* x;       // [n_windows][N]
  const double* y;       // [n_windows][N]
  int N, nt;             // nt = ceil(N / 8)
  int n_windows;
  long long problem0;    // first problem of this launch (chunking)
  double* L;             // [chunk][tiles_in_lower(nt)][64]
  double* z;             // [chunk][nt*8]   z = L^-1 y
  double* feat;          // [chunk][4][nt*8] per-point features (x, x^2, cos, sin) for phase B, or null
  double* lml;           // [n_problems] or null
  double* logdet;        // [n_problems] or null
  double* quad;          // [n_problems] or null   y' Ky^-1 y
  int* status;           // [n_problems] or null
  int jitter_retry;
};

// In-warp Cholesky of an 8x8 SPD tile held in the lane layout, followed by the inverse of the factor.
// dt / linv: 64-double shared scratch private to the calling warp.  Returns the failing pivot (1-based) or 0;
// adds sum(log L_kk) to *half_logdet.  On return linv holds inv(L) row-major (zeros above the diagonal).
__device__ __forceinline__ int chol8_inv8(tile2 c, int lane, double* dt, double* linv, double* rsd, double* half_logdet) {
  const int r = lane >> 2, q = lane & 3;
  tile_store(dt, lane, c);
  __syncwarp();
  int fail = 0;
  double hl = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double pk = dt[k * 8 + k];
    if (!(pk > 0.0) && fail == 0) fail = k + 1;
    const double d = sqrt(pk);
    const double rs = 1.0 / d;
    hl += log(d);
    const double lrk = dt[r * 8 + k] * rs;
    const double lc0 = dt[(2 * q) * 8 + k] * rs;
    const double lc1 = dt[(2 * q + 1) * 8 + k] * rs;
    __syncwarp();
    if (2 * q > k && r >= 2 * q) { c.a -= lrk * lc0; dt[r * 8 + 2 * q] = c.a; }
    if (2 * q + 1 > k && r >= 2 * q + 1) { c.b -= lrk * lc1; dt[r * 8 + 2 * q + 1] = c.b; }
    if (2 * q == k && r >= k) { c.a = (r == k) ? d : lrk; dt[r * 8 + k] = c.a; }
    if (2 * q + 1 == k && r >= k) { c.b = (r == k) ? d : lrk; dt[r * 8 + k] = c.b; }
    if (lane == 0) rsd[k] = rs;
    __syncwarp();
  }
  *half_logdet += hl;
  // inverse by columns: lane cc < 8 owns column cc of X = inv(L)
  if (lane < 8) {
    const int cc = lane;
    double xcol[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (m < rr) acc = fma(dt[rr * 8 + m], (m >= cc) ? xcol[m] : 0.0, acc);
      const double v = (rr == cc) ? rsd[rr] : -acc * rsd[rr];
      xcol[rr] = (rr >= cc) ? v : 0.0;
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) linv[rr * 8 + cc] = xcol[rr];
  }
  __syncwarp();
  return fail;
}

template <int KID>
__global__ void __launch_bounds__(FIT_THREADS) gp_fit_kernel(const FitArgs a) {
  __shared__ double fx[NPAD], fxx[NPAD], fc[NPAD], fs[NPAD];   // per-point features
  __shared__ double ys[NPAD];
  __shared__ double zs[NPAD];
  __shared__ LeafConst hc[CNGP_MAX_LEAVES];
  __shared__ KProg kps;
  __shared__ __align__(16) double dt[64];
  __shared__ __align__(16) double linv[64];
  __shared__ double rsd[8];
  __shared__ double s_red[FIT_THREADS / 32];
  __shared__ int s_fail;
  __shared__ double s_hl;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const long long lp = blockIdx.x;                 // problem within this launch
  const long long p = a.problem0 + lp;             // global problem
  const int win = (int)(p % a.n_windows);
  const long long ti = a.theta_mode == 0 ? 0 : (a.theta_mode == 1 ? (long long)win : p / a.n_windows);
  const double* th = a.theta + ti * a.theta_stride;
  const int N = a.N, nt = a.nt;
  const double noise = th[a.kp.n_params];
  FastK<KID> fk;
  if (KID != KID_GENERIC) fk.init(th);

  for (int i = tid; i < nt * 8; i += FIT_THREADS) {
    const double xv = i < N ? a.x[(long long)win * N + i] : 0.0;
    ys[i] = i < N ? a.y[(long long)win * N + i] : 0.0;
    PointFeat f{xv, __dmul_rn(xv, xv), 0.0, 0.0};
    if (KID != KID_GENERIC) f = fk.point(xv);
    fx[i] = f.x; fxx[i] = f.xx; fc[i] = f.c; fs[i] = f.s;
    if (a.feat) {
      double* fp = a.feat + lp * (long long)(4 * nt * 8);
      fp[i] = f.x; fp[nt * 8 + i] = f.xx; fp[2 * nt * 8 + i] = f.c; fp[3 * nt * 8 + i] = f.s;
    }
  }
  if (KID == KID_GENERIC) {
    if (tid < a.kp.n_leaves) hc[tid] = leaf_prepare(a.kp.leaf_type[tid], th + a.kp.leaf_param[tid]);
    if (tid == 0) kps = a.kp;
  }
  __syncthreads();

  // Ky(row, col) from the staged features; identity padding beyond N
  auto ky_entry = [&](int row, int col, const PointFeat& fa, const PointFeat& fb) -> double {
    if (row < N && col < N) {
      if (KID == KID_GENERIC) return keval_generic_sym(&kps, hc, fa.x, fb.x, row == col);
      return fk.eval(fa, fb, row == col);
    }
    return (row == col) ? 1.0 : 0.0;
  };
  auto feat_at = [&](int i) -> PointFeat {
    if (KID == KID_GENERIC) return PointFeat{fx[i], 0.0, 0.0, 0.0};
    if (KID == KID_RBF_PER) return PointFeat{fx[i], fxx[i], fc[i], fs[i]};
    return PointFeat{fx[i], fxx[i], 0.0, 0.0};
  };

  double* Lp = a.L + lp * (long long)tiles_in_lower(nt) * 64;
  const int max_attempts = a.jitter_retry ? 6 : 1;
  double extra = 0.0;
  int fail_pivot = 0, attempts_used = 0;

  for (int attempt = 0; attempt < max_attempts; ++attempt) {
    if (attempt == 1) {
      // GPy jitchol: jitter = mean(diag(Ky)) * 1e-6, then x10 per retry
      double s = 0.0;
      for (int i = tid; i < N; i += FIT_THREADS) {
        const PointFeat f = feat_at(i);
        s += ky_entry(i, i, f, f) + noise + CNGP_JITTER;
      }
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_red[w] = s;
      __syncthreads();
      s = 0.0;
      for (int i = 0; i < FIT_WARPS; ++i) s += s_red[i];
      extra = s / N * 1e-6;
    } else if (attempt > 1) {
      extra *= 10.0;
    }
    if (tid == 0) { s_fail = 0; s_hl = 0.0; }
    __syncthreads();
    const double dadd = noise + CNGP_JITTER + extra;

    for (int j = 0; j < nt; ++j) {
      // my row tiles of this column: i_t = j + w + 8 t  (t < nvalid), plus possibly the z row (i == nt)
      const int rem = nt - j - w;                       // i_t < nt  <=>  8 t < rem
      const int nvalid = rem > 0 ? min(FIT_MAXT, (rem + 7) >> 3) : 0;
      const bool zmine = rem >= 0 && (rem & 7) == 0 && (rem >> 3) < FIT_MAXT;
      // ---- left-looking update: S_t = sum_{k<j} L(i_t,k) L(j,k)^T;  column k starts at cb, tile (i,k) at cb+(i-k)*64
      tile2 S[FIT_MAXT];
#pragma unroll
      for (int t = 0; t < FIT_MAXT; ++t) S[t] = tile2{0.0, 0.0};
      tile2 Sz{0.0, 0.0};
      const double* cb = Lp + 2 * lane;
#pragma unroll 2
      for (int k = 0; k < j; ++k) {
        const double* py = cb + (j - k) * 64;
        const double* px = py + w * 64;
        const double2 yv = *reinterpret_cast<const double2*>(py);
        const tile2 Y{yv.x, yv.y};
#pragma unroll
        for (int t = 0; t < FIT_MAXT; ++t) {
          if (t < nvalid) {
            const double2 xv = *reinterpret_cast<const double2*>(px + t * (FIT_WARPS * 64));
            tile_mma(S[t], tile2{xv.x, xv.y}, Y);
          }
        }
        if (zmine) {
          tile2 X{0.0, 0.0};
          if (r == 0) { X.a = zs[8 * k + 2 * q]; X.b = zs[8 * k + 2 * q + 1]; }
          tile_mma(Sz, X, Y);
        }
        cb += (nt - k) * 64;
      }
      // ---- C_t = Ky tile - S_t (Ky evaluated here, never stored) ----
      {
        const int c0 = 8 * j + 2 * q, c1 = c0 + 1;
        const PointFeat fc0 = feat_at(c0), fc1 = feat_at(c1);
#pragma unroll
        for (int t = 0; t < FIT_MAXT; ++t) {
          if (t < nvalid) {
            const int row = 8 * (j + w + FIT_WARPS * t) + r;
            const PointFeat fr = feat_at(row);
            double v0 = ky_entry(row, c0, fr, fc0), v1 = ky_entry(row, c1, fr, fc1);
            if (row == c0 && row < N) v0 += dadd;
            if (row == c1 && row < N) v1 += dadd;
            S[t].a = v0 - S[t].a;
            S[t].b = v1 - S[t].b;
          }
        }
        if (zmine) {
          Sz.a = ((r == 0) ? ys[c0] : 0.0) - Sz.a;
          Sz.b = ((r == 0) ? ys[c1] : 0.0) - Sz.b;
        }
      }
      // ---- diagonal tile: factor + invert (warp 0 owns i == j at t == 0) ----
      double* colj = Lp + (long long)tile_index(j, j, nt) * 64;
      if (w == 0) {
        double hl = 0.0;
        const int f = chol8_inv8(S[0], lane, dt, linv, rsd, &hl);
        if (lane == 0) {
          if (f && s_fail == 0) s_fail = 8 * j + f;
          s_hl += hl;
        }
        tile_store(colj, lane, tile_load(linv, lane));
      }
      __syncthreads();
      // ---- rows below: L(i,j) = C_t inv(L_jj)^T ----
      const tile2 Yinv = tile_load(linv, lane);
#pragma unroll
      for (int t = 0; t < FIT_MAXT; ++t) {
        if (t < nvalid && (w + t) > 0) {
          tile2 Lt{0.0, 0.0};
          tile_mma(Lt, S[t], Yinv);
          tile_store(colj + (w + FIT_WARPS * t) * 64, lane, Lt);
        }
      }
      if (zmine) {
        tile2 Lt{0.0, 0.0};
        tile_mma(Lt, Sz, Yinv);
        if (r == 0) { zs[8 * j + 2 * q] = Lt.a; zs[8 * j + 2 * q + 1] = Lt.b; }
      }
      __syncthreads();
    }
    fail_pivot = s_fail;
    attempts_used = attempt;
    if (fail_pivot == 0) break;
    __syncthreads();
  }

  // ---- outputs: z, quad = z'z, logdet = 2 sum log L_kk, lml ----
  double* zp = a.z + lp * (long long)(nt * 8);
  double qs = 0.0;
  for (int i = tid; i < nt * 8; i += FIT_THREADS) {
    const double v = zs[i];
    zp[i] = v;
    qs += v * v;
  }
  for (int o = 16; o; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
  __syncthreads();
  if (lane == 0) s_red[w] = qs;
  __syncthreads();
  if (tid == 0) {
    double quad = 0.0;
    for (int i = 0; i < FIT_WARPS; ++i) quad += s_red[i];
    const double logdet = 2.0 * s_hl;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    const bool bad = fail_pivot != 0;
    if (a.quad) a.quad[p] = bad ? nanv : quad;
    if (a.logdet) a.logdet[p] = bad ? nanv : logdet;
    if (a.lml) a.lml[p] = bad ? nanv : 0.5 * (-(double)N * CNGP_LOG_2PI - logdet - quad);
    if (a.status) a.status[p] = bad ? -fail_pivot : attempts_used;
  }
}

}  // namespace cngp
