// Phase B of the batched exact GP: predictive mean and variance at M test points per window.
//
// Replaces GPy PosteriorExact._raw_predict + Gaussian.predictive_values as reached from the per-point loop
// m.predict([[x]]) at core_navigation/script/gp_slip_node.py:47-50 (row a6):
//     mu = Kx' alpha,  tmp = dtrtrs(L, Kx),  var = max(Kxx - sum(tmp^2), 1e-15) + sigma_n^2.
//
// One WARP per 8 test points.  The warp holds the 8 x N block of K*^T in registers as nt accumulator-layout tiles
// (64 doubles per lane at N = 256) and runs a right-looking forward substitution V^T = K*^T L^-T over the tile
// columns of L: V_j = R_j inv(L_jj)^T, then R_j' -= V_j L(j',j)^T for every j' > j - all tile_mma (FP64 DMMA), with
// the rows of V never leaving registers and each L tile read once per warp as a single 16-byte load per lane.
// The mean is taken from the same V:  K*' alpha = K*' L^-T L^-1 y = V' z  with z = L^-1 y from phase A, so no
// back-substitution is needed.  K* itself is evaluated straight into the accumulator registers (never stored).
#pragma once
#include "kernel_eval.cuh"

namespace cngp {

struct VarArgs {
  KProg kp;
  const double* theta;
  long long theta_stride;
  int theta_mode;          // 0 shared, 1 per window
  const double* x;         // [n_windows][N]
  const double* xstar;     // [n_windows][M] or [M]
  long long xstar_stride;  // M or 0
  int N, nt, M, mt;        // mt = ceil(M / 8)
  long long window0;       // first window of this launch
  long long n_windows_launch;
  const double* L;         // [chunk][tiles][64]  (phase A output)
  const double* z;         // [chunk][nt*8]
  const int* status;       // [n_windows] (phase A), may be null
  double* mean;            // [n_windows][M]
  double* var;             // [n_windows][M]
};

template <int NT_MAX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gp_var_kernel(const VarArgs a) {
  __shared__ LeafConst hc_all[WARPS][CNGP_MAX_LEAVES];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const long long task = (long long)blockIdx.x * WARPS + w;
  if (task >= a.n_windows_launch * a.mt) return;
  const long long lw = task / a.mt;           // window within this launch
  const int m8 = (int)(task % a.mt);
  const long long win = a.window0 + lw;
  const double* th = a.theta + (a.theta_mode == 0 ? 0 : win) * a.theta_stride;
  const int N = a.N, nt = a.nt, M = a.M;
  LeafConst* hc = hc_all[w];
  if (lane < a.kp.n_leaves) hc[lane] = leaf_prepare(a.kp.leaf_type[lane], th + a.kp.leaf_param[lane]);
  __syncwarp();

  const double* xw = a.x + win * N;
  const int m = min(8 * m8 + r, M - 1);
  const double xm = a.xstar[(a.xstar_stride ? win * a.xstar_stride : 0) + m];

  // ---- K*^T block: R[J] = tile column jc = nt-1-J of the 8 x N block, rows = my 8 test points ----
  tile2 R[NT_MAX];
#pragma unroll
  for (int J = NT_MAX - 1; J >= 0; --J) {
    R[J] = tile2{0.0, 0.0};
    if (J < nt) {
      const int c0 = 8 * (nt - 1 - J) + 2 * q;
      if (c0 < N) R[J].a = keval<false>(a.kp, hc, xw[c0], xm, false);
      if (c0 + 1 < N) R[J].b = keval<false>(a.kp, hc, xw[c0 + 1], xm, false);
    }
  }

  const double* Lp = a.L + lw * (long long)tiles_in_lower(nt) * 64;
  const double* zp = a.z + lw * (long long)(nt * 8);
  double vs = 0.0, ms = 0.0;
#pragma unroll
  for (int J = NT_MAX - 1; J >= 0; --J) {
    if (J < nt) {
      const double* col = Lp + (long long)(J * (J + 1) / 2) * 64;
      const tile2 Yd = tile_load(col, lane);
      tile2 V{0.0, 0.0};
      tile_mma(V, R[J], Yd);
      const double2 zz = *reinterpret_cast<const double2*>(zp + 8 * (nt - 1 - J) + 2 * q);
      vs = fma(V.a, V.a, vs);
      vs = fma(V.b, V.b, vs);
      ms = fma(V.a, zz.x, ms);
      ms = fma(V.b, zz.y, ms);
      const tile2 nV{-V.a, -V.b};
#pragma unroll
      for (int J2 = J - 1; J2 >= 0; --J2) {
        const tile2 Yl = tile_load(col + (J - J2) * 64, lane);
        tile_mma(R[J2], nV, Yl);
      }
    }
  }
  vs += __shfl_xor_sync(0xffffffffu, vs, 1);
  vs += __shfl_xor_sync(0xffffffffu, vs, 2);
  ms += __shfl_xor_sync(0xffffffffu, ms, 1);
  ms += __shfl_xor_sync(0xffffffffu, ms, 2);
  if (q == 0 && 8 * m8 + r < M) {
    const double noise = th[a.kp.n_params];
    const double kss = kdiag_eval(a.kp, hc, xm);
    const bool bad = a.status && a.status[win] < 0;
    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
    a.mean[win * M + m] = bad ? nanv : ms;
    a.var[win * M + m] = bad ? nanv : fmax(kss - vs, CNGP_VAR_FLOOR) + noise;
  }
}

}  // namespace cngp
