// Phase A of the batched exact GP: assemble Ky = K(X,X) + (noise + 1e-8) I, factor it, solve z = L^-1 y, LML.
//
// Replaces GPy ExactGaussianInference.inference as reached from GPRegression(...) at
// core_navigation/script/gp_slip_node.py:35 (row a3): kern.K, pdinv -> jitchol -> dpotrf, dpotrs, logdet.
//
// One CTA per (candidate, window) problem: NW worker warps + one DIAGONAL warp.  Left-looking tile Cholesky over
// 8-wide tile columns with a one-column look-ahead, so that the only serial chain of the algorithm - the in-warp
// factorisation + inversion of the 8x8 diagonal tile (chol8_inv8, ~1.3k cycles of dependent shuffles / rsqrt / fma
// per column) - runs on the diagonal warp WHILE the workers accumulate the bulk of the next column.
//
// Row tiles are OWNED: row i belongs to worker i % NW for the whole factorisation (slot t = i / NW of that warp), so a
// worker's tiles of the current column, of the next column and the factor tiles it has just produced all stay in
// registers under compile-time indices, and the k = j term of the look-ahead needs one operand from shared memory only.
//
//   workers, step j:  (b) S(i, j+1) = sum_{k<j} L(i,k) L(j+1,k)^T for their rows i >= j+1 (one pass over the tile pool:
//                         1 + A 16-byte loads and 2A DMMA per k for A live rows), evaluate Ky(i, j+1) (the kernel
//                         matrix is never materialised); the owner of row j+1 parks the diagonal tile's partial sum and
//                         carries the forward solve for z along as two scalar FMA per k on the tiles it loads anyway
//                     --- barrier 1: inv(L_jj) published by the diagonal warp ---
//                     (d) L(i,j) = C(i,j) inv(L_jj)^T -> registers, tile pool, global (for phase B); z_j
//                     --- barrier 2 ---
//                     (a) C(i, j+1) = Ky(i, j+1) - S(i, j+1) - L(i,j) L(j+1,j)^T
//   diagonal warp:    chol8_inv8(C(j,j)) -> inv(L_jj); arrive on barrier 1; wait on barrier 2;
//                     C(j+1,j+1) = parked partial - L(j+1,j) L(j+1,j)^T.
//
// Every update and the triangular solve below the diagonal (through the explicitly inverted 8x8 diagonal tile) are
// tile_mma = FP64 DMMA.
// Output factor layout (consumed by gp_var.cuh / gp_grad.cuh): column-block-major tiles, tile (j,j) holds
// inv(L_jj) (lower triangular), tiles (i>j, j) hold L_ij.
#pragma once
#include "kernel_eval.cuh"

namespace cngp {

constexpr int NPAD = CNGP_MAX_N + 8;
// Shared-memory tile pool.  With h = ceil(nt/2): columns k < h are packed one after the other, column k holding its
// rows k..nt-1 (tile (i,k) at cb(k) + i - k, cb(k) = k nt - k(k-1)/2).  Rows < h are dead once column h starts, so
// column k = h + m re-uses the dead head of column m (h - m >= nt - k tiles): tile (i,k) at cb(m) + i - k.  Walking k
// for a fixed row therefore advances by nt-1, nt-2, ... tiles in BOTH ranges, with one pointer for all rows of a
// column.  392 tiles (196 KB) at nt = 32 instead of 528.
__host__ __device__ __forceinline__ int fit_pool_tiles(int nt) {
  const int h = (nt + 1) / 2;
  return h * nt - h * (h - 1) / 2;
}
// + slack: rows beyond nt-1 of a partially filled last slot are read (never used) by the unpredicated pool loop
constexpr size_t fit_smem_bytes(int nt, int slots_total) {
  return (size_t)(((nt + 1) / 2) * nt - ((nt + 1) / 2) * (((nt + 1) / 2) - 1) / 2 + (slots_total > nt ? slots_total - nt : 0)) * 512;
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

