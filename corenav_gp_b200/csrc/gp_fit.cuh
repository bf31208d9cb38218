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

constexpr int FIT_MAXT = 3;  // row tiles per warp in the first column: ceil((nt + 1) / warps) <= 3
constexpr int NPAD = CNGP_MAX_N + 8;
// Shared-memory tile pool.  With h = ceil(nt/2): region A holds the packed lower triangle of the top-left h x h tile
// block while columns k < h are being factored, region B the (nt-h) x h block below it; rows < h are dead once column
// h starts, so columns k >= h (the packed lower triangle of the trailing (nt-h) x (nt-h) block) reuse region A.
// 392 tiles (196 KB) at nt = 32 instead of 528.
__host__ __device__ __forceinline__ int fit_pool_tiles(int nt) {
  const int h = (nt + 1) / 2;
  return h * (h + 1) / 2 + (nt - h) * h;
}
constexpr size_t fit_smem_bytes(int nt) {
  return (size_t)(((nt + 1) / 2) * ((nt + 1) / 2 + 1) / 2 + (nt - (nt + 1) / 2) * ((nt + 1) / 2)) * 512;
}

struct FitArgs {
  KProg kp;
  const double* theta;   // hyper-parameters, noise last
  long long theta_stride;
  int theta_mode;        // 0 shared, 1 per window (problem % n_windows), 2 per candidate (problem / n_windows),
                         // 3 per problem (theta row = problem; the window comes from win_map)
  const int* win_map;    // [n_problems] window of each problem (theta_mode 3), or null
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
  // KID_TILES only (large-N blocked Cholesky, chol_large.cu): the SPD block is read from tile storage instead of being
  // evaluated - tile (i, j) of the block at Asrc + j * a_col_stride + i * 64 (lower tiles only are read).
  const double* Asrc;
  long long a_col_stride;
};

// In-warp Cholesky of an 8x8 SPD tile held in the lane layout, followed by the inverse of the factor.  This sits on
// the critical path of every tile column, so it is built for latency: the tile stays in registers, column k is
// exchanged with four shuffles per step, the pivot uses rsqrt (sqrt + div cost 165 dependent cycles, rsqrt 80), and
// the logarithms of the pivots are NOT taken here - the pivots are parked in dpiv and logged in parallel at the end.
// dt / linv: 64-double shared scratch private to the calling warp.  Returns the failing pivot (1-based) or 0.
// On return linv holds inv(L) row-major (zeros above the diagonal) and dpiv[0..7] the diagonal of L.
__device__ __forceinline__ int chol8_inv8(tile2 c, int lane, double* dt, double* linv, double* dpiv) {
  const int r = lane >> 2, q = lane & 3;
  int fail = 0;
  double rsv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double v = (k & 1) ? c.b : c.a;                              // my entry of column pair k>>1
    const double pk = __shfl_sync(0xffffffffu, v, 4 * k + (k >> 1));     // T[k][k]
    const double trk = __shfl_sync(0xffffffffu, v, 4 * r + (k >> 1));    // T[r][k]
    const double tc0 = __shfl_sync(0xffffffffu, v, 8 * q + (k >> 1));    // T[2q][k]
    const double tc1 = __shfl_sync(0xffffffffu, v, 8 * q + 4 + (k >> 1));// T[2q+1][k]
    if (!(pk > 0.0) && fail == 0) fail = k + 1;
    const double rs = rsqrt(pk);
    rsv[k] = rs;
    const double lrk = trk * rs;
    if (2 * q > k && r >= 2 * q) c.a = fma(-lrk, tc0 * rs, c.a);
    if (2 * q + 1 > k && r >= 2 * q + 1) c.b = fma(-lrk, tc1 * rs, c.b);
    if (2 * q == k && r >= k) c.a = (r == k) ? pk * rs : lrk;
    if (2 * q + 1 == k && r >= k) c.b = (r == k) ? pk * rs : lrk;
  }
  tile_store(dt, lane, c);
  {
    const int kk = lane & 7, src = 4 * kk + (kk >> 1);   // L[kk][kk] lives in lane (kk, kk/2), element kk & 1
    const double da = __shfl_sync(0xffffffffu, c.a, src), db = __shfl_sync(0xffffffffu, c.b, src);
    if (lane < 8) dpiv[lane] = (lane & 1) ? db : da;
  }
  __syncwarp();
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
      const double v = (rr == cc) ? rsv[rr] : -acc * rsv[rr];
      xcol[rr] = (rr >= cc) ? v : 0.0;
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) linv[rr * 8 + cc] = xcol[rr];
  }
  __syncwarp();
  return fail;
}

template <int KID, int FIT_WARPS>
__global__ void __launch_bounds__(FIT_WARPS * 32) gp_fit_kernel(const FitArgs a) {
  constexpr int FIT_THREADS = FIT_WARPS * 32;
  extern __shared__ __align__(128) double pool[];               // tile pool, see fit_pool_tiles
  __shared__ double fx[NPAD], fxx[NPAD], fc[NPAD], fs[NPAD];   // per-point features
  __shared__ double ys[NPAD];
  __shared__ double zs[NPAD];
  __shared__ LeafConst hc[CNGP_MAX_LEAVES];
  __shared__ KProg kps;
  __shared__ __align__(16) double dt[64];
  __shared__ __align__(16) double linv[64];
  __shared__ double dpiv[NPAD];          // diagonal of L (pivots), logged in parallel at the end
  __shared__ double s_red[FIT_WARPS];
  __shared__ int s_fail;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const long long lp = blockIdx.x;                 // problem within this launch
  const long long p = a.problem0 + lp;             // global problem
  const int win = a.win_map ? a.win_map[p] : (int)(p % a.n_windows);
  const long long ti = a.theta_mode == 0 ? 0 : (a.theta_mode == 1 ? (long long)win : (a.theta_mode == 3 ? p : p / a.n_windows));
  const double* th = a.theta + ti * a.theta_stride;
  const int N = a.N, nt = a.nt;
  const int h = (nt + 1) / 2, nb = nt - h;          // tile rows of the top block / bottom block
  const double noise = (KID == KID_TILES) ? 0.0 : th[a.kp.n_params];
  FastK<KID> fk;
  if (KID != KID_GENERIC && KID != KID_TILES) fk.init(th);

  for (int i = tid; i < nt * 8; i += FIT_THREADS) {
    const double xv = (KID != KID_TILES && i < N) ? a.x[(long long)win * N + i] : 0.0;
    ys[i] = (KID != KID_TILES && i < N) ? a.y[(long long)win * N + i] : 0.0;
    PointFeat f{xv, __dmul_rn(xv, xv), 0.0, 0.0};
    if (KID != KID_GENERIC && KID != KID_TILES) f = fk.point(xv);
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
    if (KID == KID_TILES) return 0.0;   // never used: the tiles are loaded below
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
  double* poolA = pool + 2 * lane;                       // lane-offset views of the two pool regions
  double* poolB = pool + (h * (h + 1) / 2) * 64 + 2 * lane;
  const int max_attempts = (a.jitter_retry && KID != KID_TILES) ? 6 : 1;
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
    if (tid == 0) s_fail = 0;
    __syncthreads();
    const double dadd = noise + CNGP_JITTER + extra;

    for (int j = 0; j < nt; ++j) {
      // my row tiles of this column: i_t = j + w + FIT_WARPS t  (t < nvalid), plus possibly the z row (i == nt)
      const int rem = nt - j - w;                       // i_t < nt  <=>  FIT_WARPS t < rem
      const int nvalid = rem > 0 ? min(FIT_MAXT, (rem + FIT_WARPS - 1) / FIT_WARPS) : 0;
      const bool zmine = rem >= 0 && (rem % FIT_WARPS) == 0 && (rem / FIT_WARPS) < FIT_MAXT;
      // ---- left-looking update: S_t = sum_{k<j} L(i_t,k) L(j,k)^T, every tile read from the shared-memory pool
      tile2 S[FIT_MAXT];
      const double* xb[FIT_MAXT];    // per-tile base: region + i_t * 64 (the column part is added per k)
      bool inA[FIT_MAXT];
#pragma unroll
      for (int t = 0; t < FIT_MAXT; ++t) {
        S[t] = tile2{0.0, 0.0};
        const int i = j + w + FIT_WARPS * t;
        inA[t] = (j < h) && (i < h);
        xb[t] = (inA[t] ? poolA : poolB) + i * 64;
      }
      tile2 Sz{0.0, 0.0};
      const int kA_end = j < h ? j : h;
      {
        // columns k < h:  A-tile (i,k) at A[k h - k(k-1)/2 + i - k],  B-tile (i,k) at B[k nb + i - h]
        int offA = 0, offB = -h;
        const double* yb = (j < h ? poolA : poolB) + j * 64;
        const bool yA = j < h;
#pragma unroll 2
        for (int k = 0; k < kA_end; ++k) {
          const double2 yv = *reinterpret_cast<const double2*>(yb + (yA ? offA : offB) * 64);
          const tile2 Y{yv.x, yv.y};
#pragma unroll
          for (int t = 0; t < FIT_MAXT; ++t) {
            if (t < nvalid) {
              const double2 xv = *reinterpret_cast<const double2*>(xb[t] + (inA[t] ? offA : offB) * 64);
              tile_mma(S[t], tile2{xv.x, xv.y}, Y);
            }
          }
          if (zmine) {
            tile2 X{0.0, 0.0};
            if (r == 0) { X.a = zs[8 * k + 2 * q]; X.b = zs[8 * k + 2 * q + 1]; }
            tile_mma(Sz, X, Y);
          }
          offA += h - k - 1;
          offB += nb;
        }
      }
      if (j > h) {
        // columns h <= k < j live in region A again: tile (i,k) at A[(k-h) nb - (k-h)(k-h-1)/2 + i - k]
        int off2 = -h;
        const double* yb = poolA + j * 64;
#pragma unroll 2
        for (int k = h; k < j; ++k) {
          const double2 yv = *reinterpret_cast<const double2*>(yb + off2 * 64);
          const tile2 Y{yv.x, yv.y};
#pragma unroll
          for (int t = 0; t < FIT_MAXT; ++t) {
            if (t < nvalid) {
              const double2 xv = *reinterpret_cast<const double2*>(poolA + (off2 + j + w + FIT_WARPS * t) * 64);
              tile_mma(S[t], tile2{xv.x, xv.y}, Y);
            }
          }
          if (zmine) {
            tile2 X{0.0, 0.0};
            if (r == 0) { X.a = zs[8 * k + 2 * q]; X.b = zs[8 * k + 2 * q + 1]; }
            tile_mma(Sz, X, Y);
          }
          off2 += nb - (k - h) - 1;
        }
      }
      // ---- C_t = Ky tile - S_t (Ky evaluated here, never stored) ----
      {
        const int c0 = 8 * j + 2 * q, c1 = c0 + 1;
        const PointFeat fc0 = feat_at(c0), fc1 = feat_at(c1);
#pragma unroll
        for (int t = 0; t < FIT_MAXT; ++t) {
          if (t < nvalid) {
            const int row = 8 * (j + w + FIT_WARPS * t) + r;
            double v0, v1;
            if (KID == KID_TILES) {
              const double2 av = *reinterpret_cast<const double2*>(
                  a.Asrc + j * a.a_col_stride + (long long)(j + w + FIT_WARPS * t) * 64 + 2 * lane);
              v0 = av.x; v1 = av.y;
            } else {
              const PointFeat fr = feat_at(row);
              v0 = ky_entry(row, c0, fr, fc0); v1 = ky_entry(row, c1, fr, fc1);
              if (row == c0 && row < N) v0 += dadd;
              if (row == c1 && row < N) v1 += dadd;
            }
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
        const int f = chol8_inv8(S[0], lane, dt, linv, dpiv + 8 * j);
        if (lane == 0 && f && s_fail == 0) s_fail = 8 * j + f;
        tile_store(colj, lane, tile_load(linv, lane));
      }
      __syncthreads();
      // ---- rows below: L(i,j) = C_t inv(L_jj)^T  -> pool (for later columns) and global (for phase B) ----
      const tile2 Yinv = tile_load(linv, lane);
      // pool position of tile (i, j): same mapping as above with k = j
      const int pj = j < h ? (j * h - j * (j - 1) / 2 - j) : ((j - h) * nb - (j - h) * (j - h - 1) / 2 - j);
      const int pjB = j * nb - h;
#pragma unroll
      for (int t = 0; t < FIT_MAXT; ++t) {
        if (t < nvalid && (w + t) > 0) {
          const int i = j + w + FIT_WARPS * t;
          tile2 Lt{0.0, 0.0};
          tile_mma(Lt, S[t], Yinv);
          tile_store(colj + (i - j) * 64, lane, Lt);
          double* dst = (j < h && i >= h) ? (poolB + (pjB + i) * 64) : (poolA + (pj + i) * 64);
          *reinterpret_cast<double2*>(dst) = make_double2(Lt.a, Lt.b);
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
  double qs = 0.0, hl = 0.0;
  for (int i = tid; i < nt * 8; i += FIT_THREADS) {
    const double v = zs[i];
    zp[i] = v;
    qs += v * v;
    hl += log(dpiv[i]);      // padded rows have pivot 1
  }
  for (int o = 16; o; o >>= 1) {
    qs += __shfl_xor_sync(0xffffffffu, qs, o);
    hl += __shfl_xor_sync(0xffffffffu, hl, o);
  }
  __syncthreads();
  if (lane == 0) { s_red[w] = qs; fx[w] = hl; }   // fx is free by now
  __syncthreads();
  if (tid == 0) {
    double quad = 0.0, hls = 0.0;
    for (int i = 0; i < FIT_WARPS; ++i) { quad += s_red[i]; hls += fx[i]; }
    const double logdet = 2.0 * hls;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    const bool bad = fail_pivot != 0;
    if (a.quad) a.quad[p] = bad ? nanv : quad;
    if (a.logdet) a.logdet[p] = bad ? nanv : logdet;
    if (a.lml) a.lml[p] = bad ? nanv : 0.5 * (-(double)N * CNGP_LOG_2PI - logdet - quad);
    if (a.status) a.status[p] = bad ? -fail_pivot : attempts_used;
  }
}

}  // namespace cngp
