// Phase A of the batched exact GP: assemble Ky = K(X,X) + (noise + 1e-8) I, factor it, solve z = L^-1 y, LML.
//
// Replaces GPy ExactGaussianInference.inference as reached from GPRegression(...) at
// core_navigation/script/gp_slip_node.py:35 (row a3): kern.K, pdinv -> jitchol -> dpotrf, dpotrs, logdet.
//
// One CTA per (candidate, window) problem: NW worker warps + one DIAGONAL warp.  Left-looking tile Cholesky over
// 8-wide tile columns with a one-column look-ahead, so that the only serial chain of the algorithm - the in-warp
// factorisation + inversion of the 8x8 diagonal tile (chol8_inv8, ~1.3k cycles of dependent shuffles / rsqrt / fma
// per column) - runs on the diagonal warp WHILE the workers accumulate the bulk of the next column:
//
//   workers, column j:  (b) S(i, j+1) = sum_{k<j} L(i,k) L(j+1,k)^T  for their rows of column j+1, evaluate Ky(i, j+1)
//                           (the kernel matrix is never materialised), park the diagonal row's partial sum in smem
//                       --- barrier 1: inv(L_jj) published by the diagonal warp ---
//                       (d) L(i,j) = C(i,j) inv(L_jj)^T  -> tile pool (for later columns) and global (for phase B)
//                       --- barrier 2 ---
//                       (a) C(i, j+1) = Ky(i, j+1) - S(i, j+1) - L(i,j) L(j+1,j)^T
//   diagonal warp:      chol8_inv8(C(j,j)) -> inv(L_jj); arrive on barrier 1; wait on barrier 2;
//                       C(j+1,j+1) = parked partial - L(j+1,j) L(j+1,j)^T.
//
// Every update, the triangular solve below the diagonal (through the explicitly inverted 8x8 diagonal tile) and the
// forward solve for z (carried as an extra 1-row "tile row" under the matrix) are tile_mma = FP64 DMMA.
// Output factor layout (consumed by gp_var.cuh / gp_grad.cuh): column-block-major tiles, tile (j,j) holds
// inv(L_jj) (lower triangular), tiles (i>j, j) hold L_ij.
#pragma once
#include "kernel_eval.cuh"

namespace cngp {

constexpr int FIT_MAXT = 2;  // row tiles per worker warp and column: nt <= 2 NW
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// In-warp Cholesky of an 8x8 SPD tile held in the lane layout, fused with the inverse of the factor.  This is the
// critical path of every tile column, so it is built for latency: the tile stays in registers, column k is exchanged
// with independent shuffles, the pivot uses rsqrt (sqrt + div cost 165 dependent cycles, rsqrt 80), and the inverse
// rides along for free: the row operations of the elimination are applied to an identity tile in the same step
// (forward substitution L X = I by columns: X[k] = x[k] / L[k][k], x[r] -= L[r][k] X[k]), so there is no second pass.
// The logarithms of the pivots are NOT taken here - the pivots are parked in dpiv and logged in parallel at the end.
// On return linv holds inv(L) row-major (zeros above the diagonal) and dpiv[0..7] the diagonal of L.
// Returns the failing pivot (1-based) or 0.
__device__ __forceinline__ int chol8_inv8(tile2 c, int lane, double* linv, double* dpiv) {
  const int r = lane >> 2, q = lane & 3;
  int fail = 0;
  tile2 x{r == 2 * q ? 1.0 : 0.0, r == 2 * q + 1 ? 1.0 : 0.0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double v = (k & 1) ? c.b : c.a;                              // my entry of column pair k>>1
    const double pk = __shfl_sync(0xffffffffu, v, 4 * k + (k >> 1));     // T[k][k]
    const double trk = __shfl_sync(0xffffffffu, v, 4 * r + (k >> 1));    // T[r][k]
    const double tc0 = __shfl_sync(0xffffffffu, v, 8 * q + (k >> 1));    // T[2q][k]
    const double tc1 = __shfl_sync(0xffffffffu, v, 8 * q + 4 + (k >> 1));// T[2q+1][k]
    const double xka = __shfl_sync(0xffffffffu, x.a, 4 * k + q);         // x[k][2q]
    const double xkb = __shfl_sync(0xffffffffu, x.b, 4 * k + q);         // x[k][2q+1]
    if (!(pk > 0.0) && fail == 0) fail = k + 1;
    const double rs = rsqrt(pk);
    const double lrk = trk * rs;
    if (2 * q > k && r >= 2 * q) c.a = fma(-lrk, tc0 * rs, c.a);
    if (2 * q + 1 > k && r >= 2 * q + 1) c.b = fma(-lrk, tc1 * rs, c.b);
    const double xsa = xka * rs, xsb = xkb * rs;
    if (r == k) { x.a = xsa; x.b = xsb; }
    if (r > k) { x.a = fma(-lrk, xsa, x.a); x.b = fma(-lrk, xsb, x.b); }
    if (lane == 0) dpiv[k] = pk * rs;
  }
  tile_store(linv, lane, x);
  __syncwarp();
  return fail;
}

template <int KID, int NW>
__global__ void __launch_bounds__((NW + 1) * 32) gp_fit_kernel(const FitArgs a) {
  constexpr int FIT_THREADS = (NW + 1) * 32;
  extern __shared__ __align__(128) double pool[];               // tile pool, see fit_pool_tiles
  __shared__ double fx[NPAD], fxx[NPAD], fc[NPAD], fs[NPAD];   // per-point features
  __shared__ double ys[NPAD];
  __shared__ __align__(16) double zs[NPAD];
  __shared__ LeafConst hc[CNGP_MAX_LEAVES];
  __shared__ KProg kps;
  __shared__ __align__(16) double linv[64];      // inv(L_jj) of the current column (diagonal warp -> workers)
  __shared__ __align__(16) double dpart[64];     // Ky(j+1,j+1) - sum_{k<j} ... (worker 0 -> diagonal warp)
  __shared__ __align__(16) double etab[128];     // exp_tab tables scaled by the leaf variances
  __shared__ double dpiv[NPAD];          // diagonal of L (pivots), logged in parallel at the end
  __shared__ double s_red[NW + 1], s_red2[NW + 1];
  __shared__ int s_fail;
  __shared__ __align__(16) double zero2[2];

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const long long lp = blockIdx.x;                 // problem within this launch
  const long long p = a.problem0 + lp;             // global problem
  const int win = a.win_map ? a.win_map[p] : (int)(p % a.n_windows);
  const long long ti = a.theta_mode == 0 ? 0 : (a.theta_mode == 1 ? (long long)win : (a.theta_mode == 3 ? p : p / a.n_windows));
  const double* th = a.theta + ti * a.theta_stride;
  const int N = a.N, nt = a.nt;
  const int h = (nt + 1) / 2, nb = nt - h;          // tile rows of the top block / bottom block
  const int hh = h * (h + 1) / 2;                   // tiles in region A
  const double noise = (KID == KID_TILES) ? 0.0 : th[a.kp.n_params];
  FastK<KID> fk;
  if (KID != KID_GENERIC && KID != KID_TILES) {
    fk.init(th);
    if (tid < 64) {
      const double t = EXP2_TAB64[tid];
      etab[tid] = fk.scale1() * t;
      etab[64 + tid] = fk.scale2() * t;
    }
  }

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

  auto feat_at = [&](int i) -> PointFeat {
    if (KID == KID_GENERIC) return PointFeat{fx[i], 0.0, 0.0, 0.0};
    if (KID == KID_RBF_PER) return PointFeat{fx[i], fxx[i], fc[i], fs[i]};
    return PointFeat{fx[i], fxx[i], 0.0, 0.0};
  };
  // Ky(row, col) without the diagonal term; identity padding beyond N
  auto ky_entry = [&](int row, int col, const PointFeat& fa, const PointFeat& fb) -> double {
    if (KID == KID_TILES) return 0.0;   // never used: the tiles are loaded
    if (row < N && col < N) {
      if (KID == KID_GENERIC) return keval_generic_sym(&kps, hc, fa.x, fb.x, row == col);
      return fk.eval_tab(fa, fb, row == col, etab);
    }
    return (row == col) ? 1.0 : 0.0;
  };
  // tile (i, c) of Ky in the lane layout (dadd on the diagonal entries), evaluated or - KID_TILES - loaded
  auto ky_tile = [&](int i, int c, double dadd) -> tile2 {
    if (KID == KID_TILES) {
      const double2 av = *reinterpret_cast<const double2*>(a.Asrc + c * a.a_col_stride + (long long)i * 64 + 2 * lane);
      return tile2{av.x, av.y};
    }
    const int c0 = 8 * c + 2 * q, c1 = c0 + 1, row = 8 * i + r;
    const PointFeat fr = feat_at(row), f0 = feat_at(c0), f1 = feat_at(c1);
    double v0 = ky_entry(row, c0, fr, f0), v1 = ky_entry(row, c1, fr, f1);
    if (row == c0 && row < N) v0 += dadd;
    if (row == c1 && row < N) v1 += dadd;
    return tile2{v0, v1};
  };
  // pool slot of factor tile (i, k)
  auto pool_idx = [&](int i, int k) -> int {
    if (k < h) return (i < h) ? (k * h - k * (k - 1) / 2 + i - k) : (hh + k * nb + i - h);
    return (k - h) * nb - (k - h) * (k - h - 1) / 2 + i - k;
  };

  double* Lp = a.L + lp * (long long)tiles_in_lower(nt) * 64;
  double* poolL = pool + 2 * lane;                       // lane-offset view of the pool
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
      for (int i = 0; i <= NW; ++i) s += s_red[i];
      extra = s / N * 1e-6;
    } else if (attempt > 1) {
      extra *= 10.0;
    }
    if (tid == 0) { s_fail = 0; zero2[0] = 0.0; zero2[1] = 0.0; }
    __syncthreads();
    const double dadd = noise + CNGP_JITTER + extra;

    if (w == NW) {
      // ================= diagonal warp =================
      tile2 Cd = ky_tile(0, 0, dadd);
      for (int j = 0; j < nt; ++j) {
        const int f = chol8_inv8(Cd, lane, linv, dpiv + 8 * j);
        if (lane == 0 && f && s_fail == 0) s_fail = 8 * j + f;
        named_bar_arrive(1, FIT_THREADS);                           // inv(L_jj) is in linv
        tile_store(Lp + (long long)tile_index(j, j, nt) * 64, lane, tile_load(linv, lane));
        named_bar_sync(2, FIT_THREADS);                             // column j stored, dpart of column j+1 parked
        if (j + 1 < nt) {
          const tile2 Y = tile_load(pool + pool_idx(j + 1, j) * 64, lane);
          tile2 S = tile_load(dpart, lane);
          tile_mma(S, tile2{-Y.a, -Y.b}, Y);
          Cd = S;
        }
      }
    } else {
      // ================= worker warps =================
      // Slot t of this warp in column c is row tile i = c + w + NW t:  i == c the diagonal tile (worker 0 only
      // pre-accumulates it for the diagonal warp), c < i < nt a regular tile, i == nt the z row (y carried as an extra
      // 1-row tile row under the matrix).  Column 0 has nt + 1 rows for 2 NW slots, so its z row sits in the slot of
      // the diagonal tile (0,0), which the diagonal warp evaluates itself.
      auto slot_kind = [&](int c, int t) -> int {   // 0 none, 1 regular, 2 z row, 3 diagonal
        const int i = c + w + NW * t;
        if (c == 0 && w == 0 && t == 0) return 2;
        if (i == c) return 3;
        if (i < nt) return 1;
        return (i == nt && c > 0) ? 2 : 0;
      };
      auto y_tile = [&](int c) -> tile2 {
        return tile2{r == 0 ? ys[8 * c + 2 * q] : 0.0, r == 0 ? ys[8 * c + 2 * q + 1] : 0.0};
      };
      // z-row operand as a tile: row 0 = z[8k .. 8k+7], rows 1..7 = 0  (lanes r > 0 read a zero word with stride 0)
      const double* zbase = (r == 0) ? zs + 2 * q : zero2;
      const int zstride = (r == 0) ? 8 : 0;

      tile2 Ccur[FIT_MAXT];
#pragma unroll
      for (int t = 0; t < FIT_MAXT; ++t) {
        const int kd = slot_kind(0, t);
        Ccur[t] = kd == 1 ? ky_tile(w + NW * t, 0, dadd) : (kd == 2 ? y_tile(0) : tile2{0.0, 0.0});
      }

      for (int j = 0; j < nt; ++j) {
        const int c = j + 1;
        // ---- (b) look-ahead: partial sums of column c over k < j, Ky tiles of column c ----
        tile2 S0[FIT_MAXT], S1[FIT_MAXT], Kn[FIT_MAXT];
        int kind[FIT_MAXT];
#pragma unroll
        for (int t = 0; t < FIT_MAXT; ++t) {
          kind[t] = c < nt ? slot_kind(c, t) : 0;
          S0[t] = tile2{0.0, 0.0};
          S1[t] = tile2{0.0, 0.0};
          Kn[t] = tile2{0.0, 0.0};
        }
        if (kind[0] != 0) {
          // One loop serves both pool regions and the z row: every operand is (pointer, stride, stride decrement) -
          // region A columns shrink by one tile per k, region B columns have a fixed pitch, the z operand strides 8.
          auto accumulate = [&](int k0, int k1, const double* pY, int sY, int dY, const double* pX0, int sX0, int dX0,
                                const double* pX1, int sX1, int dX1, const bool two) {
            if (k0 >= k1) return;
            tile2 Y = tile_load(pY, 0), X0 = tile_load(pX0, 0), X1{0.0, 0.0};
            if (two) X1 = tile_load(pX1, 0);
            for (int k = k0; k < k1; k += 2) {
              pY += sY; sY -= dY; pX0 += sX0; sX0 -= dX0; pX1 += sX1; sX1 -= dX1;
              tile2 Yn{0.0, 0.0}, X0n{0.0, 0.0}, X1n{0.0, 0.0};
              if (k + 1 < k1) {
                Yn = tile_load(pY, 0); X0n = tile_load(pX0, 0);
                if (two) X1n = tile_load(pX1, 0);
              }
              tile_mma(S0[0], X0, Y);
              if (two) tile_mma(S0[1], X1, Y);
              if (k + 1 >= k1) break;
              pY += sY; sY -= dY; pX0 += sX0; sX0 -= dX0; pX1 += sX1; sX1 -= dX1;
              if (k + 2 < k1) {
                Y = tile_load(pY, 0); X0 = tile_load(pX0, 0);
                if (two) X1 = tile_load(pX1, 0);
              }
              tile_mma(S1[0], X0n, Yn);
              if (two) tile_mma(S1[1], X1n, Yn);
            }
          };
          const bool two = kind[1] != 0;
          const int k1 = j < h ? j : h;
          // columns k < min(j, h): tile (i,k) at A[k(h-1) - k(k-1)/2 + i] for i < h, at B[k nb + i - h] otherwise
          {
            const double* pA = poolL;
            const double* pB = poolL + (hh - h) * 64;
            const bool yA = c < h;
            const double* pX[FIT_MAXT]; int sX[FIT_MAXT], dX[FIT_MAXT];
#pragma unroll
            for (int t = 0; t < FIT_MAXT; ++t) {
              const int i = c + w + NW * t;
              const bool xa = i < h;
              pX[t] = kind[t] == 2 ? zbase : (xa ? pA : pB) + i * 64;
              sX[t] = kind[t] == 2 ? zstride : (xa ? (h - 1) * 64 : nb * 64);
              dX[t] = (kind[t] != 2 && xa) ? 64 : 0;
            }
            accumulate(0, k1, (yA ? pA : pB) + c * 64, yA ? (h - 1) * 64 : nb * 64, yA ? 64 : 0, pX[0], sX[0], dX[0],
                       pX[1], sX[1], dX[1], two);
          }
          // columns h <= k < j live in region A again: tile (i,k) at A[(k-h) nb - (k-h)(k-h-1)/2 + i - k]
          if (j > h) {
            const double* p2 = poolL - h * 64;
            const double* pX[FIT_MAXT]; int sX[FIT_MAXT], dX[FIT_MAXT];
#pragma unroll
            for (int t = 0; t < FIT_MAXT; ++t) {
              const int i = c + w + NW * t;
              pX[t] = kind[t] == 2 ? zbase + zstride * h : p2 + i * 64;
              sX[t] = kind[t] == 2 ? zstride : (nb - 1) * 64;
              dX[t] = kind[t] == 2 ? 0 : 64;
            }
            accumulate(h, j, p2 + c * 64, (nb - 1) * 64, 64, pX[0], sX[0], dX[0], pX[1], sX[1], dX[1], two);
          }
          // Ky tiles of column c (never stored); the diagonal row's partial result goes to the diagonal warp
#pragma unroll
          for (int t = 0; t < FIT_MAXT; ++t) {
            if (kind[t] != 0) {
              S0[t].a += S1[t].a; S0[t].b += S1[t].b;
              Kn[t] = kind[t] == 2 ? y_tile(c) : ky_tile(c + w + NW * t, c, dadd);
            }
          }
          if (kind[0] == 3) tile_store(dpart, lane, tile2{Kn[0].a - S0[0].a, Kn[0].b - S0[0].b});
        }
        named_bar_sync(1, FIT_THREADS);                             // inv(L_jj) published
        // ---- (d) rows below the diagonal: L(i,j) = C(i,j) inv(L_jj)^T -> pool and global; z_j ----
        {
          const tile2 Yinv = tile_load(linv, lane);
          double* colj = Lp + (long long)tile_index(j, j, nt) * 64;
#pragma unroll
          for (int t = 0; t < FIT_MAXT; ++t) {
            const int kd = slot_kind(j, t);
            if (kd == 1 || kd == 2) {
              const int i = j + w + NW * t;
              tile2 Lt{0.0, 0.0};
              tile_mma(Lt, Ccur[t], Yinv);
              if (kd == 1) {
                tile_store(pool + pool_idx(i, j) * 64, lane, Lt);
                tile_store(colj + (i - j) * 64, lane, Lt);
              } else if (r == 0) {
                zs[8 * j + 2 * q] = Lt.a; zs[8 * j + 2 * q + 1] = Lt.b;
              }
            }
          }
        }
        named_bar_sync(2, FIT_THREADS);
        // ---- (a) finish column c with the k = j term ----
        if (kind[0] != 0) {
          const tile2 Y = tile_load(pool + pool_idx(c, j) * 64, lane);
#pragma unroll
          for (int t = 0; t < FIT_MAXT; ++t) {
            if (kind[t] == 1 || kind[t] == 2) {
              const int i = c + w + NW * t;
              const tile2 X = kind[t] == 2 ? tile_load(zbase + zstride * j, 0) : tile_load(pool + pool_idx(i, j) * 64, lane);
              tile_mma(S0[t], X, Y);
              Ccur[t] = tile2{Kn[t].a - S0[t].a, Kn[t].b - S0[t].b};
            }
          }
        }
      }
    }
    __syncthreads();
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
  if (lane == 0) { s_red[w] = qs; s_red2[w] = hl; }
  __syncthreads();
  if (tid == 0) {
    double quad = 0.0, hls = 0.0;
    for (int i = 0; i <= NW; ++i) { quad += s_red[i]; hls += s_red2[i]; }
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
