// Phase B of the batched exact GP: predictive mean and variance at M test points per window.
//
// Replaces GPy PosteriorExact._raw_predict + Gaussian.predictive_values as reached from the per-point loop
// m.predict([[x]]) at core_navigation/script/gp_slip_node.py:47-50 (row a6):
//     mu = Kx' alpha,  tmp = dtrtrs(L, Kx),  var = max(Kxx - sum(tmp^2), 1e-15) + sigma_n^2.
//
// Work decomposition.  One WARP owns 8 test points: it holds the 8 x N block of K*^T in registers as nt
// accumulator-layout tiles (64 doubles per lane at N = 256) and runs a right-looking forward substitution
// V^T = K*^T L^-T over the tile columns of L:  V_j = R_j inv(L_jj)^T, then R_j' -= V_j L(j',j)^T for every j' > j -
// all tile_mma (FP64 DMMA); the rows of V never leave registers.  mean = V' z with z = L^-1 y from phase A (no
// back-substitution), var = k** - sum V^2.
//
// Data movement.  The substitution consumes the packed tiles of L in exactly their storage order, so a persistent CTA
// (one per SM) takes one window at a time and streams that window's factor (270 KB at N = 256) ONCE per round of
// WARPS tasks through a CTA-shared shared-memory ring filled by 1-D bulk asynchronous copies (cp.async.bulk -> UBLKCP,
// completion counted on mbarriers).  There is no producer warp: the last warp to release a ring slot issues the copy
// that refills it, VAR_NSLOT chunks ahead, across round and window boundaries.  Every tile then costs each warp one
// 16-byte LDS per lane and two DMMA.  K*^T is evaluated by a rolled loop (small code) into a per-warp staging buffer
// and picked up into the accumulator registers; it never touches global memory.
#pragma once
#include "kernel_eval.cuh"

namespace cngp {

constexpr int VAR_CT = 16;                      // tiles per chunk (8 KB)
constexpr int VAR_NSLOT = 4;                    // ring slots per CTA
constexpr int VAR_CHUNK_DOUBLES = VAR_CT * 64;
constexpr int VAR_CHUNK_BYTES = VAR_CHUNK_DOUBLES * 8;
constexpr int VAR_STAGE_TILES = 8;              // K* staging: tile columns per pass
constexpr size_t var_smem_bytes(int warps) {
  return (size_t)VAR_NSLOT * VAR_CHUNK_BYTES + (size_t)warps * VAR_STAGE_TILES * 512 + 128;
}

struct VarArgs {
  KProg kp;
  const double* theta;
  long long theta_stride;
  int theta_mode;          // 0 shared, 1 per window
  const double* xstar;     // [n_windows][M] or [M]
  long long xstar_stride;  // M or 0
  int N, nt, M, mt;        // mt = ceil(M / 8)
  long long window0;       // first window of this launch
  long long n_windows_launch;
  const double* L;         // [chunk][tiles][64]  (phase A output; one chunk of slack before the first window)
  const double* z;         // [chunk][nt*8]
  const double* feat;      // [chunk][4][nt*8]  per-point features of the training inputs (phase A output)
  const int* status;       // [n_windows] (phase A), may be null
  double* mean;            // [n_windows][M]
  double* var;             // [n_windows][M]
  int sigma_mode;          // 0: var;  1: write sigma = 2 sqrt(var) instead (gp_slip_node.py:61)
};

struct VarShared {
  unsigned long long full[VAR_NSLOT];   // mbarriers: chunk landed
  int cnt[VAR_NSLOT];                   // warps that have released the slot
  int cur_it, cur_round, cur_ce;        // producer cursor: next chunk to load
  int ce_max, nrounds, nwin_cta, n_tiles;
  const double* Lbase;
  uint32_t ring_u32, full_u32;
};

// Load the chunk under the producer cursor into `slot` and advance the cursor.  Called by one lane at a time (the
// prologue, then whichever lane performed the last release of a slot); kept out of line: it is the rare path.
static __device__ __noinline__ void var_issue_next(VarShared* sh, int slot) {
  const int it = sh->cur_it;
  if (it >= sh->nwin_cta) return;
  const int ce = sh->cur_ce;
  const long long lw = blockIdx.x + (long long)it * gridDim.x;
  const double* src = sh->Lbase + (lw + 1) * (long long)sh->n_tiles * 64 - (long long)(VAR_CT * ce + VAR_CT) * 64;
  mbar_expect_tx(sh->full_u32 + 8 * slot, VAR_CHUNK_BYTES);
  bulk_g2s(sh->ring_u32 + slot * VAR_CHUNK_BYTES, src, VAR_CHUNK_BYTES, sh->full_u32 + 8 * slot);
  if (ce > 0) {
    sh->cur_ce = ce - 1;
  } else {
    sh->cur_ce = sh->ce_max;
    if (sh->cur_round + 1 < sh->nrounds) sh->cur_round = sh->cur_round + 1;
    else { sh->cur_round = 0; sh->cur_it = it + 1; }
  }
}

// FULL: nt == NT_MAX and N == 8 nt (no padding) - drops every per-column guard from the unrolled code.
template <int NT_MAX, int WARPS, int KID, bool FULL>
__global__ void __launch_bounds__(WARPS * 32, 1) gp_var_kernel(const VarArgs a) {
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ LeafConst hc_all[KID == KID_GENERIC ? WARPS : 1][CNGP_MAX_LEAVES];
  __shared__ KProg kps;
  __shared__ VarShared sh;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int N = FULL ? NT_MAX * 8 : a.N, nt = FULL ? NT_MAX : a.nt, M = a.M, mt = a.mt;
  const int n8 = nt * 8;
  const int n_tiles = tiles_in_lower(nt);
  const int ce_max = (n_tiles - 1) / VAR_CT;      // chunk ids count from the END of a window's factor
  const int nchunks = ce_max + 1;
  const int nrounds = (mt + WARPS - 1) / WARPS;
  const int nwin_cta = (int)((a.n_windows_launch - blockIdx.x + gridDim.x - 1) / gridDim.x);   // windows of this CTA

  double* ring = reinterpret_cast<double*>(dsm);
  double* stage = ring + VAR_NSLOT * VAR_CHUNK_DOUBLES + (size_t)w * VAR_STAGE_TILES * 64 + 2 * lane;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t full_u32 = smem_u32(&sh.full[0]);

  if (threadIdx.x == 0) {
    if (KID == KID_GENERIC) kps = a.kp;
    for (int s = 0; s < VAR_NSLOT; ++s) { mbar_init(full_u32 + 8 * s, 1); sh.cnt[s] = 0; }
    sh.cur_it = 0; sh.cur_round = 0; sh.cur_ce = ce_max;
    sh.ce_max = ce_max; sh.nrounds = nrounds; sh.nwin_cta = nwin_cta; sh.n_tiles = n_tiles;
    sh.Lbase = a.L; sh.ring_u32 = ring_u32; sh.full_u32 = full_u32;
    mbar_fence_init();
    for (int s = 0; s < VAR_NSLOT; ++s) var_issue_next(&sh, s);   // global chunk g goes to slot g % VAR_NSLOT
  }
  __syncthreads();

  LeafConst* hc = hc_all[KID == KID_GENERIC ? w : 0];
  int g = 0;   // global chunk counter of this warp (over windows x rounds x chunks)

  for (int it = 0; it < nwin_cta; ++it) {
    const long long lw = blockIdx.x + (long long)it * gridDim.x;   // window within this launch
    const long long win = a.window0 + lw;
    const double* th = a.theta + (a.theta_mode == 0 ? 0 : win) * a.theta_stride;
    FastK<KID> fk;
    if (KID == KID_GENERIC) {
      __syncwarp();
      if (lane < a.kp.n_leaves) hc[lane] = leaf_prepare(a.kp.leaf_type[lane], th + a.kp.leaf_param[lane]);
      __syncwarp();
    } else {
      fk.init(th);
    }
    const double* fp = a.feat + lw * (long long)(4 * n8);
    const double* zp = a.z + lw * (long long)n8;
    const double noise = th[a.kp.n_params];
    const bool bad = a.status && a.status[win] < 0;

    for (int round = 0; round < nrounds; ++round) {
      const int m8 = round * WARPS + w;
      int slot = g % VAR_NSLOT;
      uint32_t parity = (uint32_t)(g / VAR_NSLOT) & 1u;
      g += nchunks;
      // Every warp releases every chunk (so no warp can run more than VAR_NSLOT chunks ahead of another, which
      // is what makes waiting on a phase PARITY safe); the last one out refills the slot VAR_NSLOT chunks ahead.
      auto release_chunk = [&]() {
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          const int old = atomicAdd(&sh.cnt[slot], 1);
          if (old == WARPS - 1) {
            sh.cnt[slot] = 0;
            __threadfence_block();
            var_issue_next(&sh, slot);
          }
        }
        if (++slot == VAR_NSLOT) { slot = 0; parity ^= 1u; }
      };
      if (m8 >= mt) {   // no task for this warp in the (last) round: just keep the ring protocol going
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(full_u32 + 8 * slot, parity);
          release_chunk();
        }
        continue;
      }

      const int m = min(8 * m8 + r, M - 1);
      const double xm = a.xstar[(a.xstar_stride ? win * a.xstar_stride : 0) + m];
      PointFeat fm{xm, 0.0, 0.0, 0.0};
      if (KID != KID_GENERIC) fm = fk.point(xm);

      // ---- K*^T block into R: rolled evaluation through the per-lane staging slots ----
      tile2 R[NT_MAX];
#pragma unroll
      for (int p = NT_MAX / VAR_STAGE_TILES - 1; p >= 0; --p) {
#pragma unroll 1
        for (int jj = VAR_STAGE_TILES - 1; jj >= 0; --jj) {
          const int J = p * VAR_STAGE_TILES + jj;
          double2 v = make_double2(0.0, 0.0);
          if (FULL || J < nt) {
            const int c0 = 8 * (nt - 1 - J) + 2 * q;
            const double2 x2 = *reinterpret_cast<const double2*>(fp + c0);
            if (KID == KID_GENERIC) {
              if (c0 < N) v.x = keval_generic_cross(&kps, hc, x2.x, xm);
              if (c0 + 1 < N) v.y = keval_generic_cross(&kps, hc, x2.y, xm);
            } else {
              const double2 xx2 = *reinterpret_cast<const double2*>(fp + n8 + c0);
              double2 cc2 = make_double2(0.0, 0.0), ss2 = make_double2(0.0, 0.0);
              if (KID == KID_RBF_PER) {
                cc2 = *reinterpret_cast<const double2*>(fp + 2 * n8 + c0);
                ss2 = *reinterpret_cast<const double2*>(fp + 3 * n8 + c0);
              }
              if (c0 < N) v.x = fk.eval(PointFeat{x2.x, xx2.x, cc2.x, ss2.x}, fm, false);
              if (c0 + 1 < N) v.y = fk.eval(PointFeat{x2.y, xx2.y, cc2.y, ss2.y}, fm, false);
            }
          }
          *reinterpret_cast<double2*>(stage + jj * 64) = v;
        }
#pragma unroll
        for (int jj = 0; jj < VAR_STAGE_TILES; ++jj) {
          const double2 v = *reinterpret_cast<const double2*>(stage + jj * 64);
          R[p * VAR_STAGE_TILES + jj] = tile2{v.x, v.y};
        }
      }

      // ---- forward substitution over the streamed tiles ----
      double vs = 0.0, ms = 0.0;
      const double* cptr = ring + slot * VAR_CHUNK_DOUBLES + 2 * lane;
      // tile at distance e (1-based) from the end of the factor: chunk id and offset inside the chunk are
      // compile-time after unrolling; the ring slot advances by one per chunk.
      auto next_tile = [&](const int e, const bool first) -> tile2 {
        const int ce = (e - 1) / VAR_CT;
        if ((!FULL && first) || (e - 1) % VAR_CT == VAR_CT - 1) mbar_wait(full_u32 + 8 * slot, parity);
        const double2 v = *reinterpret_cast<const double2*>(cptr + (VAR_CT * ce + VAR_CT - e) * 64);
        if ((e - 1) % VAR_CT == 0) {     // chunk drained by this warp
          release_chunk();
          cptr = ring + slot * VAR_CHUNK_DOUBLES + 2 * lane;
        }
        return tile2{v.x, v.y};
      };

#pragma unroll
      for (int J = NT_MAX - 1; J >= 0; --J) {
        if (FULL || J < nt) {
          const int e0 = (J + 1) * (J + 2) / 2;   // distance from the end of tile (J, d = 0)
          const tile2 Yd = next_tile(e0, J == nt - 1);
          tile2 V{0.0, 0.0};
          tile_mma(V, R[J], Yd);
          const double2 zz = *reinterpret_cast<const double2*>(zp + 8 * (nt - 1 - J) + 2 * q);
          vs = fma(V.a, V.a, vs);
          vs = fma(V.b, V.b, vs);
          ms = fma(V.a, zz.x, ms);
          ms = fma(V.b, zz.y, ms);
          const tile2 nV{-V.a, -V.b};
          // two tiles per step with their DMMA pairs interleaved (a dependent DMMA costs 26 cycles, issue 16)
#pragma unroll
          for (int J2 = J - 1; J2 >= 1; J2 -= 2) {
            const tile2 Ya = next_tile(e0 - (J - J2), false);
            const tile2 Yb = next_tile(e0 - (J - J2) - 1, false);
            dmma884(R[J2].a, R[J2].b, nV.a, Ya.a);
            dmma884(R[J2 - 1].a, R[J2 - 1].b, nV.a, Yb.a);
            dmma884(R[J2].a, R[J2].b, nV.b, Ya.b);
            dmma884(R[J2 - 1].a, R[J2 - 1].b, nV.b, Yb.b);
          }
          if (J & 1) {   // J tiles below the diagonal: one left over when J is odd
            const tile2 Yl = next_tile(e0 - J, false);
            tile_mma(R[0], nV, Yl);
          }
        }
      }
      vs += __shfl_xor_sync(0xffffffffu, vs, 1);
      vs += __shfl_xor_sync(0xffffffffu, vs, 2);
      ms += __shfl_xor_sync(0xffffffffu, ms, 1);
      ms += __shfl_xor_sync(0xffffffffu, ms, 2);
      if (q == 0 && 8 * m8 + r < M) {
        const double kss = (KID == KID_GENERIC) ? kdiag_eval(kps, hc, xm) : fk.kdiag(xm);
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        a.mean[win * M + m] = bad ? nanv : ms;
        const double vv = fmax(kss - vs, CNGP_VAR_FLOOR) + noise;
        a.var[win * M + m] = bad ? nanv : (a.sigma_mode ? 2.0 * sqrt(vv) : vv);
      }
    }
  }
}

}  // namespace cngp
