/* CPU restatement of the slip extraction + GP window recorder of CoreNav::Update.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ as the checker of cngp_slip_record_batch.  Nothing under corenav_gp_b200/
 * may call it.  PARITY UNPINNED: the reference holds no test or golden vector for this code and cannot be built here
 * (ROS + Eigen); the restatement follows the source line by line and is cross-checked against an independent numpy
 * transcription (tests/test_oracle_slip.py).
 *
 * Follows core_navigation/src/CoreNav.cpp (one call of slip_record_one = the odometry updates of one drive):
 *   :176        odomUptCount = odomUptCount + 1                      (a double, CoreNav.h:265)
 *   :178-183    wheel speeds from the joint rates, rearVel = (vBL + vBR) / 2
 *   :190,560-581 Cn2bUnc = eul_to_dcm(att): only row 0 is needed, (cthe cpsi, cthe spsi, -sthe)
 *   :244-245    vlin = row0(Cn2bUnc v), slip = max over the four wheels of (v_wheel - vlin) / v_wheel
 *   :247-258    slip = 0 when |rearVel| < 0.001, clamp to [-1, 1]
 *   :264-329    recorder: first valid driving sample arms the window (start = count + 10, stop = start + 150);
 *               samples with start < count < stop are recorded; at count >= stop the window is published if it holds
 *               at least 15 samples; a stop command re-arms (start = stop + ceil(cmd_stop) 10 + 60); a drive that
 *               outlives its window by more than 10 s re-initialises.
 * Inputs per update k: joint[4] (rad/s), att[3] (pre-update attitude), vel[3] (post-update INS velocity, what
 * CoreNav.cpp:244 reads), cmd (cmd[0]), stop_cmd (seconds-to-stop received since the previous update, NaN = none).   */
#include <math.h>
#include <stdint.h>

typedef struct {
  double wheel_radius;   /* InsConst.h:17  0.11 */
  double cmd_min;        /* CoreNav.cpp:264  0.2 */
  double rear_min;       /* :247  0.001 */
  int32_t arm_delay;     /* :270  10 */
  int32_t window;        /* :271  150 */
  int32_t min_samples;   /* :300  15 */
} slip_cfg;

void slip_oracle_default_cfg(slip_cfg* c) {
  c->wheel_radius = 0.11; c->cmd_min = 0.2; c->rear_min = 0.001; c->arm_delay = 10; c->window = 150; c->min_samples = 15;
}

/* std::max on doubles: returns its first argument unless it is smaller than the second (NaN-order sensitive) */
static double std_max(double x, double y) { return (x < y) ? y : x; }

double slip_oracle_slip(const double* joint, const double* att, const double* vel, const slip_cfg* c) {
  const double vFL = -joint[0] * c->wheel_radius, vFR = joint[1] * c->wheel_radius;
  const double vBL = -joint[2] * c->wheel_radius, vBR = joint[3] * c->wheel_radius;
  const double rear = (vBL + vBR) / 2.0;
  const double cpsi = cos(att[2]), spsi = sin(att[2]), cthe = cos(att[1]), sthe = sin(att[1]);
  const double vlin = ((cthe * cpsi) * vel[0] + (cthe * spsi) * vel[1]) + (-sthe) * vel[2];
  double slip = std_max(std_max((vFR - vlin) / vFR, (vBR - vlin) / vBR), std_max((vFL - vlin) / vFL, (vBL - vlin) / vBL));
  if (fabs(rear) < c->rear_min) slip = 0.0;
  if (slip < -1.0) slip = -1.0;
  if (slip > 1.0) slip = 1.0;
  return slip;
}

/* One drive of T odometry updates.  Outputs: slip[T]; up to max_windows windows of at most `cap` samples each
 * (time_array / slip_array [max_windows][cap], n_samples[max_windows], published[max_windows], stop_update[max_windows]
 * = index k of the update that closed the window, where CoreNav.cpp:291-292 latches savePos / P_pred).
 * Returns the number of windows closed. */
int32_t slip_record_one(const double* joint, const double* att, const double* vel, const double* cmd,
                        const double* stop_cmd, int32_t T, const slip_cfg* c, int32_t max_windows, int32_t cap,
                        double* slip_out, double* time_array, double* slip_array, int32_t* n_samples,
                        int32_t* published, int32_t* stop_update) {
  double count = 0.0, start = 0.0, stop = 0.0;          /* CoreNav.cpp:1066-1068 */
  int first_driving = 1, gp_flag = 0, new_stop = 0;
  double cmd_stop = 0.0;
  int32_t n_win = 0, n_cur = 0;
  for (int32_t k = 0; k < T; ++k) {
    if (stop_cmd && !isnan(stop_cmd[k])) { cmd_stop = stop_cmd[k]; new_stop = 1; }     /* CoreNav.cpp:756-758 */
    count = count + 1.0;
    const double slip = slip_oracle_slip(joint + 4 * k, att + 3 * k, vel + 3 * k, c);
    if (slip_out) slip_out[k] = slip;
    if (slip != 0.0 && slip != -1.0 && slip != 1.0 && fabs(cmd[k]) > c->cmd_min) {
      if (first_driving) {
        start = count + c->arm_delay;
        stop = start + c->window;
        first_driving = 0;
      }
      if (count > start && count < stop && !gp_flag) {
        if (n_win < max_windows && n_cur < cap) {
          time_array[(int64_t)n_win * cap + n_cur] = count;
          slip_array[(int64_t)n_win * cap + n_cur] = slip;
        }
        ++n_cur;
      }
      if (count >= stop) {
        if (!gp_flag) {
          gp_flag = 1;
          if (n_win < max_windows) {
            n_samples[n_win] = n_cur;
            published[n_win] = n_cur >= c->min_samples;
            stop_update[n_win] = k;
          }
          ++n_win;
          n_cur = 0;
        }
        if (new_stop) {
          new_stop = 0;
          start = stop + ceil(cmd_stop) * 10 + 10 + 50;
          stop = start + c->window;
          gp_flag = 0;
        }
      }
      if (!first_driving && count / 10 - stop / 10 > 10) {
        n_cur = 0;
        first_driving = 1;
        gp_flag = 0;
      }
    }
  }
  return n_win;
}

void slip_record_batch(const double* joint, const double* att, const double* vel, const double* cmd, const double* stop_cmd,
                       int64_t B, int32_t T, const slip_cfg* c, int32_t max_windows, int32_t cap, double* slip_out,
                       double* time_array, double* slip_array, int32_t* n_samples, int32_t* published,
                       int32_t* stop_update, int32_t* n_windows) {
  for (int64_t b = 0; b < B; ++b)
    n_windows[b] = slip_record_one(joint + b * T * 4, att + b * T * 3, vel + b * T * 3, cmd + b * T,
                                   stop_cmd ? stop_cmd + b * T : 0, T, c, max_windows, cap,
                                   slip_out ? slip_out + b * T : 0, time_array + b * max_windows * cap,
                                   slip_array + b * max_windows * cap, n_samples + b * max_windows,
                                   published + b * max_windows, stop_update + b * max_windows);
}
