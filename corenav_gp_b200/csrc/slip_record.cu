// Slip extraction + GP window recorder for B independent drives: the step
// immediately upstream of the GP.  Replaces, per odometry update, CoreNav::Update at
// core_navigation/src/CoreNav.cpp:176-183 (wheel speeds), :190 + :560-581 (row 0 of eul_to_dcm), :244-258 (slip,
// dead-band, clamp) and :264-329 (the recorder state machine that fills core_nav/GP_Input).
//
// One WARP per drive.  The per-update arithmetic is elementwise, so a warp takes 32 consecutive updates at a time with
// coalesced loads (12 doubles in, 1 out per update: the kernel is HBM-bound by construction - 104 algorithmic bytes per
// update).  The recorder is a sequential state machine over the updates, but it only looks at one bit per update
// ("valid driving sample") plus the rare stop command, so after a ballot every lane replays the 32 steps of the chunk
// redundantly in registers (warp-uniform, no divergence, no shared memory) and keeps the output position of its own
// update; recorded samples are then written by their own lanes.
// The sum order of vlin and the explicit __dmul_rn/__dadd_rn (no fma contraction) follow the C oracle, so slip
// differs from the CPU value only through the last-ulp differences of sin/cos.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cngp.h"

namespace {

__device__ __forceinline__ double std_max(double x, double y) { return (x < y) ? y : x; }   // std::max semantics

struct SlipArgs {
  const double *joint, *att, *vel, *cmd, *stop_cmd;
  long long B;
  int T;
  cngp_slip_config c;
  int max_windows, cap;
  double *slip, *time_array, *slip_array;
  int *n_samples, *published, *stop_update, *n_windows;
};

constexpr int SLIP_WARPS = 8;

__global__ void __launch_bounds__(SLIP_WARPS * 32) slip_record_kernel(const SlipArgs a) {
  const int lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * SLIP_WARPS + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const int T = a.T;
  const double* joint = a.joint + b * (long long)T * 4;
  const double* att = a.att + b * (long long)T * 3;
  const double* vel = a.vel + b * (long long)T * 3;
  const double* cmd = a.cmd + b * (long long)T;
  const double* scmd = a.stop_cmd ? a.stop_cmd + b * (long long)T : nullptr;
  double* tarr = a.time_array + b * (long long)a.max_windows * a.cap;
  double* sarr = a.slip_array + b * (long long)a.max_windows * a.cap;
  int* nsam = a.n_samples + b * a.max_windows;
  int* publ = a.published + b * a.max_windows;
  int* stup = a.stop_update + b * a.max_windows;

  // recorder state (identical in every lane)
  double count = 0.0, start = 0.0, stop = 0.0, cmd_stop = 0.0;
  bool first_driving = true, gp_flag = false, new_stop = false;
  int n_win = 0, n_cur = 0;

  for (int k0 = 0; k0 < T; k0 += 32) {
    const int k = k0 + lane;
    const bool in = k < T;
    double slip = 0.0, cm = 0.0, sc = 0.0;
    bool has_stop = false;
    if (in) {
      const double2 j01 = *reinterpret_cast<const double2*>(joint + 4 * (long long)k);
      const double2 j23 = *reinterpret_cast<const double2*>(joint + 4 * (long long)k + 2);
      const double vFL = __dmul_rn(-j01.x, a.c.wheel_radius), vFR = __dmul_rn(j01.y, a.c.wheel_radius);
      const double vBL = __dmul_rn(-j23.x, a.c.wheel_radius), vBR = __dmul_rn(j23.y, a.c.wheel_radius);
      const double rear = __dadd_rn(vBL, vBR) / 2.0;
      const double the = att[3 * (long long)k + 1], psi = att[3 * (long long)k + 2];
      double spsi, cpsi, sthe, cthe;
      sincos(psi, &spsi, &cpsi);
      sincos(the, &sthe, &cthe);
      const double v0 = vel[3 * (long long)k], v1 = vel[3 * (long long)k + 1], v2 = vel[3 * (long long)k + 2];
      const double vlin = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(cthe, cpsi), v0), __dmul_rn(__dmul_rn(cthe, spsi), v1)),
                                    __dmul_rn(-sthe, v2));
      slip = std_max(std_max(__dadd_rn(vFR, -vlin) / vFR, __dadd_rn(vBR, -vlin) / vBR),
                     std_max(__dadd_rn(vFL, -vlin) / vFL, __dadd_rn(vBL, -vlin) / vBL));
      if (fabs(rear) < a.c.rear_min) slip = 0.0;
      if (slip < -1.0) slip = -1.0;
      if (slip > 1.0) slip = 1.0;
      cm = cmd[k];
      if (scmd) { sc = scmd[k]; has_stop = !isnan(sc); }
      if (a.slip) a.slip[b * (long long)T + k] = slip;
    }
    const bool valid = in && slip != 0.0 && slip != -1.0 && slip != 1.0 && fabs(cm) > a.c.cmd_min;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned smask = __ballot_sync(0xffffffffu, has_stop);
    const int nk = min(32, T - k0);
    int my_win = -1, my_pos = 0;
    // Fast paths: chunks in which the recorder cannot change mode need no replay.  count runs over c0+1 .. c0+nk.
    const double c_last = count + nk;
    bool replay = true;
    if (smask == 0u) {
      if (first_driving) {
        replay = vmask != 0u;                              // unarmed and nothing valid: only the counter advances
      } else if (c_last < stop) {                          // armed, the window cannot close inside this chunk
        if (!gp_flag) {
          const bool rec = valid && (count + (lane + 1)) > start;
          const unsigned rmask = __ballot_sync(0xffffffffu, rec);
          if (rec) { my_win = n_win; my_pos = n_cur + __popc(rmask & ((1u << lane) - 1u)); }
          n_cur += __popc(rmask);
        }
        replay = false;
      } else if (gp_flag && !new_stop && count + 1.0 >= stop && c_last - stop <= 99.0) {
        replay = false;                                    // window closed, no command pending, too early to re-initialise
      }
    }
    if (!replay) count = c_last;
    for (int i = 0; replay && i < nk; ++i) {
      if ((smask >> i) & 1u) { cmd_stop = __shfl_sync(0xffffffffu, sc, i); new_stop = true; }
      count += 1.0;
      if ((vmask >> i) & 1u) {
        if (first_driving) {
          start = count + a.c.arm_delay;
          stop = start + a.c.window;
          first_driving = false;
        }
        if (count > start && count < stop && !gp_flag) {
          if (lane == i) { my_win = n_win; my_pos = n_cur; }
          ++n_cur;
        }
        if (count >= stop) {
          if (!gp_flag) {
            gp_flag = true;
            if (lane == 0 && n_win < a.max_windows) {
              nsam[n_win] = n_cur;
              publ[n_win] = n_cur >= a.c.min_samples;
              stup[n_win] = k0 + i;
            }
            ++n_win;
            n_cur = 0;
          }
          if (new_stop) {
            new_stop = false;
            start = stop + ceil(cmd_stop) * 10 + 10 + 50;
            stop = start + a.c.window;
            gp_flag = false;
          }
        }
        // count / 10 - stop / 10 > 10 (CoreNav.cpp:323) needs count - stop of about 100: the two divisions are only
        // evaluated when that is possible (all three are integer-valued doubles)
        if (!first_driving && count - stop > 99.0 && count / 10 - stop / 10 > 10) {
          n_cur = 0;
          first_driving = true;
          gp_flag = false;
        }
      }
    }
    if (my_win >= 0 && my_win < a.max_windows && my_pos < a.cap) {
      tarr[(long long)my_win * a.cap + my_pos] = (double)(k + 1);
      sarr[(long long)my_win * a.cap + my_pos] = slip;
    }
  }
  if (lane == 0) a.n_windows[b] = n_win;
}

}  // namespace

extern "C" int cngp_launch_slip_record(const double* joint, const double* att, const double* vel, const double* cmd,
                                       const double* stop_cmd, long long B, int T, const cngp_slip_config* cfg,
                                       int max_windows, int cap, double* slip, double* time_array, double* slip_array,
                                       int* n_samples, int* published, int* stop_update, int* n_windows, cudaStream_t s) {
  SlipArgs a{joint, att, vel, cmd, stop_cmd, B, T, *cfg, max_windows, cap, slip, time_array, slip_array,
             n_samples, published, stop_update, n_windows};
  const long long grid = (B + SLIP_WARPS - 1) / SLIP_WARPS;
  slip_record_kernel<<<(unsigned)grid, SLIP_WARPS * 32, 0, s>>>(a);
  return (int)cudaGetLastError();
}
