/* CPU oracle for the stop-predictor half of the hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  Nothing under corenav_gp_b200/
 * links, includes or calls it.
 *
 * PARITY PIN: the reference holds no test or golden vector for this path, so the restatement is pinned on the
 * reference's own code: oracle/_ref compiles /root/reference/gp_predictor/src/gp_predictor.cpp UNMODIFIED against
 * stand-in ROS / Eigen / message headers (oracle/ref_stubs/, oracle/ref_gp_predictor.py), and tests/test_ref_stop.py
 * holds this file to it - decisions equal, xy-error trace / final P / R_IP to 1e-12 - directly and through the
 * committed vectors tests/golden/stop_ref_golden.npz.  (Eigen and roscpp themselves are absent from the image; what
 * the stand-ins assume is written in their headers.)  This is a plain-C restatement of
 *   /root/reference/gp_predictor/src/gp_predictor.cpp:30-46   unpack of the SetStopping response
 *   /root/reference/gp_predictor/src/gp_predictor.cpp:64-124  look-ahead loop
 *   /root/reference/gp_predictor/src/gp_predictor.cpp:144-178 llh_to_enu
 *   /root/reference/core_navigation/src/CoreNav.cpp:652-676   server-side packing (H aliasing quirk)
 * also checked by closed-form known-answer tests in tests/test_oracle_stop.py.
 *
 * Reference quirks reproduced on purpose:
 *   q1  H is read as H(r,c) = Hvec[r*4 + c] for r<4, c<15 (gp_predictor.cpp:38-42) - the aliasing
 *       index both sides use; fix_h_packing != 0 switches to the intended r*15+c.
 *   q2  sigma is used directly as the sigma-point offset and the UT "covariance" is squared again
 *       inside max(floor^2, cov^2) (gp_predictor.cpp:70-82).
 *   q3  init_llh / init_ecef are explicit inputs (the reference never calls LoadParameters).
 *   q5  order inside a step: propagate, then (every ratio-th step) update + i++, then error check.
 *   q6  only the +3 sigma point is used for the trigger.
 *
 * Every dot product is an ascending-index chain of fma() starting from 0, with additive terms
 * (Q, R, identity) applied after the chain; the CUDA look-ahead kernel follows the same order so the
 * (triggered, i) decision is reproducible bit for bit.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -fPIC -shared -o oracle/_build/libstop_oracle.so oracle/stop_oracle.c -lm
 */
#include <math.h>
#include <string.h>

typedef struct {
  double v_nom;     /* 0.8    gp_predictor.cpp:73-75 */
  double floor_a;   /* 0.03   :80-81 */
  double floor_b;   /* 0.05   :82-83 */
  double track;     /* 0.685  :85 (T_r_) */
  double scale;     /* 25     :88 */
  double thresh;    /* 3.00   :102 */
  int ratio;        /* 5      :64,67 (IMU steps per odometry update) */
  int fix_h_packing;/* 0 = reference behaviour */
  int trig_mode;    /* 0 = libm sin/cos/tan (what the reference calls); 1 = the deterministic det_* below, whose
                       operation sequence the CUDA kernel repeats bit for bit */
  int pad_;
  double init_llh[3];   /* config/init_params.yaml:13-16 */
  double init_ecef[3];  /* config/init_params.yaml:9-12 */
} stop_cfg;

void stop_oracle_default_cfg(stop_cfg* c) {
  c->v_nom = 0.8; c->floor_a = 0.03; c->floor_b = 0.05; c->track = 0.685; c->scale = 25.0;
  c->thresh = 3.0; c->ratio = 5; c->fix_h_packing = 0; c->trig_mode = 0; c->pad_ = 0;
  c->init_llh[0] = 0.693457963620326; c->init_llh[1] = -1.39498384275845; c->init_llh[2] = 334.993517334743;
  c->init_ecef[0] = 859153.015300000; c->init_ecef[1] = -4836303.72660000; c->init_ecef[2] = 4055378.50100000;
}


/* Deterministic sin / cos / tan: Cody-Waite reduction by pi/2 in three fma steps, then the classic degree-13 / 14
 * minimax polynomials on [-pi/4, pi/4] evaluated as fma Horner chains.  Only IEEE-exact operations (fma, mul, add,
 * div, rint) in a fixed order, so a host compiled with -ffp-contract=off and a CUDA device compiled with -fmad=false
 * return the same bits; the result is within 2 ulp of libm for |x| < 1e5 (tests/test_oracle_stop.py).  This is what
 * makes the ZUPT decision reproducible at a 0-ulp margin (trig_mode = 1); trig_mode = 0 keeps libm as the reference. */
static void det_sincos(double x, double* sn, double* cs) {
  const double k = rint(x * 6.36619772367581382433e-01);
  double r = fma(-k, 1.57079632679489655800e+00, x);
  r = fma(-k, 6.12323399573676603587e-17, r);
  r = fma(-k, -1.49738490485916983000e-33, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  const double c = w + (((1.0 - w) - hz) + (z * z) * pc);
  const long long q = (long long)k & 3;
  *sn = (q == 0) ? s : (q == 1) ? c : (q == 2) ? -s : -c;
  *cs = (q == 0) ? c : (q == 1) ? -s : (q == 2) ? -c : s;
}
void stop_oracle_det_sincos(double x, double* sn, double* cs) { det_sincos(x, sn, cs); }

static void trig3(double x, int mode, double* sn, double* cs, double* tn) {
  if (mode) {
    det_sincos(x, sn, cs);
    *tn = *sn / *cs;
  } else {
    *sn = sin(x); *cs = cos(x); *tn = tan(x);
  }
}

/* gp_predictor.cpp:144-178 */
void stop_oracle_llh_to_enu(double lat, double lon, double height, const stop_cfg* c, double enu[3]) {
  double phi = lat, lambda = lon, h = height;
  double a = 6378137.0000, b = 6356752.3142;
  double boa = b / a;
  double e = sqrt(1 - boa * boa);
  double sinphi, cosphi, tp, sinlam, coslam, unused;
  trig3(phi, c->trig_mode, &sinphi, &cosphi, &tp);
  trig3(lambda, c->trig_mode, &sinlam, &coslam, &unused);
  double tan2phi = tp * tp;
  double tmp2 = 1 - e * e;
  double tmpden = sqrt(1 + tmp2 * tan2phi);
  double x1 = (a * coslam) / tmpden + h * coslam * cosphi;
  double y1 = (a * sinlam) / tmpden + h * sinlam * cosphi;
  double tmp3 = sqrt(1 - e * e * sinphi * sinphi);
  double z1 = (a * tmp2 * sinphi) / tmp3 + h * sinphi;
  double d0 = x1 - c->init_ecef[0], d1 = y1 - c->init_ecef[1], d2 = z1 - c->init_ecef[2];
  double sinPhi, cosPhi, sinLam, cosLam;
  trig3(c->init_llh[0], c->trig_mode, &sinPhi, &cosPhi, &unused);
  trig3(c->init_llh[1], c->trig_mode, &sinLam, &cosLam, &unused);
  double R[3][3] = {{-1 * sinLam, cosLam, 0},
                    {(-1 * sinPhi) * cosLam, (-1 * sinPhi) * sinLam, cosPhi},
                    {cosPhi * cosLam, cosPhi * sinLam, sinPhi}};
  for (int r = 0; r < 3; ++r) {
    double acc = 0.0;
    acc = fma(R[r][0], d0, acc);
    acc = fma(R[r][1], d1, acc);
    acc = fma(R[r][2], d2, acc);
    enu[r] = acc;
  }
}

/* closed-form 4x4 inverse (adjugate / determinant through 2x2 minors), the same math as Eigen's fixed-size
 * Matrix4d::inverse() used at gp_predictor.cpp:90 */
static void inv4(const double m[4][4], double o[4][4]) {
  double s0 = m[0][0] * m[1][1] - m[1][0] * m[0][1];
  double s1 = m[0][0] * m[1][2] - m[1][0] * m[0][2];
  double s2 = m[0][0] * m[1][3] - m[1][0] * m[0][3];
  double s3 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
  double s4 = m[0][1] * m[1][3] - m[1][1] * m[0][3];
  double s5 = m[0][2] * m[1][3] - m[1][2] * m[0][3];
  double c5 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  double c4 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  double c3 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  double c2 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  double c1 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  double c0 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  double id = 1.0 / det;
  o[0][0] = ( m[1][1] * c5 - m[1][2] * c4 + m[1][3] * c3) * id;
  o[0][1] = (-m[0][1] * c5 + m[0][2] * c4 - m[0][3] * c3) * id;
  o[0][2] = ( m[3][1] * s5 - m[3][2] * s4 + m[3][3] * s3) * id;
  o[0][3] = (-m[2][1] * s5 + m[2][2] * s4 - m[2][3] * s3) * id;
  o[1][0] = (-m[1][0] * c5 + m[1][2] * c2 - m[1][3] * c1) * id;
  o[1][1] = ( m[0][0] * c5 - m[0][2] * c2 + m[0][3] * c1) * id;
  o[1][2] = (-m[3][0] * s5 + m[3][2] * s2 - m[3][3] * s1) * id;
  o[1][3] = ( m[2][0] * s5 - m[2][2] * s2 + m[2][3] * s1) * id;
  o[2][0] = ( m[1][0] * c4 - m[1][1] * c2 + m[1][3] * c0) * id;
  o[2][1] = (-m[0][0] * c4 + m[0][1] * c2 - m[0][3] * c0) * id;
  o[2][2] = ( m[3][0] * s4 - m[3][1] * s2 + m[3][3] * s0) * id;
  o[2][3] = (-m[2][0] * s4 + m[2][1] * s2 - m[2][3] * s0) * id;
  o[3][0] = (-m[1][0] * c3 + m[1][1] * c1 - m[1][2] * c0) * id;
  o[3][1] = ( m[0][0] * c3 - m[0][1] * c1 + m[0][2] * c0) * id;
  o[3][2] = (-m[3][0] * s3 + m[3][1] * s1 - m[3][2] * s0) * id;
  o[3][3] = ( m[2][0] * s3 - m[2][1] * s1 + m[2][2] * s0) * id;
}

/* gp_predictor.cpp:69-88: unscented transform of slip -> odometry velocity noise R_IP */
void stop_oracle_ut_R(double mean, double sigma, const stop_cfg* c, double R[4][4]) {
  double chi0_slip = mean, chi1_slip = mean + sigma, chi2_slip = mean - sigma;
  double chi0 = c->v_nom / (1.0 - chi0_slip);
  double chi1 = c->v_nom / (1.0 - chi1_slip);
  double chi2 = c->v_nom / (1.0 - chi2_slip);
  double est = (chi0 + chi1 + chi2) / 3.0;
  double cov = ((chi0 - est) * (chi0 - est) + (chi1 - est) * (chi1 - est) + (chi2 - est) * (chi2 - est)) / 3.0;
  double c2 = cov * cov;
  double fa = c->floor_a * c->floor_a, fb = c->floor_b * c->floor_b;
  double R2[4] = {fmax(fa, c2), fmax(fa, c2), fmax(fb, c2), fb};
  double R1[4][4] = {{0.5, 0.5, 0.0, 0.0},
                     {1 / c->track, -1 / c->track, 0.0, 0.0},
                     {0.0, 0.0, 1.0, 0.0},
                     {0.0, 0.0, 0.0, 1.0}};
  double B[4][4];
  for (int a = 0; a < 4; ++a)
    for (int k = 0; k < 4; ++k) B[a][k] = (c->scale * R1[a][k]) * R2[k];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc = fma(B[a][k], R1[b][k], acc);
      R[a][b] = acc;
    }
}

static void propagate(double P[15][15], const double F[15][15], const double Q[15][15]) {
  double T[15][15];
  for (int r = 0; r < 15; ++r)
    for (int k = 0; k < 15; ++k) {
      double acc = 0.0;
      for (int j = 0; j < 15; ++j) acc = fma(F[r][j], P[j][k], acc);
      T[r][k] = acc;
    }
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 15; ++c) {
      double acc = 0.0;
      for (int k = 0; k < 15; ++k) acc = fma(T[r][k], F[c][k], acc);
      P[r][c] = acc + Q[r][c];
    }
}

static void joseph_update(double P[15][15], const double H[4][15], const double R[4][4], double* K_out) {
  double PHt[15][4], HP[4][15], S[4][4], Si[4][4], K[15][4], IKH[15][15], T[15][15], KR[15][4];
  for (int r = 0; r < 15; ++r)
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      for (int c = 0; c < 15; ++c) acc = fma(P[r][c], H[m][c], acc);
      PHt[r][m] = acc;
    }
  /* S = (H P) H' + R: C++ parses  H_*P_pred*H_.transpose()  (gp_predictor.cpp:90) left to right */
  for (int m = 0; m < 4; ++m)
    for (int c = 0; c < 15; ++c) {
      double acc = 0.0;
      for (int j = 0; j < 15; ++j) acc = fma(H[m][j], P[j][c], acc);
      HP[m][c] = acc;
    }
  for (int m = 0; m < 4; ++m)
    for (int n = 0; n < 4; ++n) {
      double acc = 0.0;
      for (int c = 0; c < 15; ++c) acc = fma(HP[m][c], H[n][c], acc);
      S[m][n] = acc + R[m][n];
    }
  inv4(S, Si);
  for (int r = 0; r < 15; ++r)
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      for (int n = 0; n < 4; ++n) acc = fma(PHt[r][n], Si[n][m], acc);
      K[r][m] = acc;
    }
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 15; ++c) {
      double acc = 0.0;
      for (int m = 0; m < 4; ++m) acc = fma(K[r][m], H[m][c], acc);
      IKH[r][c] = (r == c ? 1.0 : 0.0) - acc;
    }
  for (int r = 0; r < 15; ++r)
    for (int k = 0; k < 15; ++k) {
      double acc = 0.0;
      for (int j = 0; j < 15; ++j) acc = fma(IKH[r][j], P[j][k], acc);
      T[r][k] = acc;
    }
  for (int r = 0; r < 15; ++r)
    for (int n = 0; n < 4; ++n) {
      double acc = 0.0;
      for (int m = 0; m < 4; ++m) acc = fma(K[r][m], R[m][n], acc);
      KR[r][n] = acc;
    }
  for (int r = 0; r < 15; ++r)
    for (int c = 0; c < 15; ++c) {
      double a1 = 0.0, a2 = 0.0;
      for (int k = 0; k < 15; ++k) a1 = fma(T[r][k], IKH[c][k], a1);
      for (int n = 0; n < 4; ++n) a2 = fma(KR[r][n], K[c][n], a2);
      P[r][c] = a1 + a2;
    }
  if (K_out)
    for (int r = 0; r < 15; ++r)
      for (int m = 0; m < 4; ++m) K_out[r * 4 + m] = K[r][m];
}

/* GpPredictor::GPCallBack look-ahead (gp_predictor.cpp:64-124) for one window.
 * mean/sigma: GP_Output arrays of length M.  Pvec/Qvec/STMvec: row-major 15x15 (r*15+c), Hvec: 60 doubles as
 * packed by CoreNav::setStopping_, pos: savePos (lat, lon rad; h m).
 * Outputs: *triggered (0/1), *i_stop = number of odometry updates performed when the loop ended (the
 * reference's `i`), *step_stop = slip_i at the trigger (or 5*M if none), *xy_err = last xy error computed,
 * xy_trace (optional, length ratio*M) = xy error per step, P_out (optional, 225) = final covariance.
 */
int stop_oracle_lookahead_ex(const double* mean, const double* sigma, int M,
                             const double* Pvec, const double* Qvec, const double* STMvec, const double* Hvec,
                             const double* pos, const stop_cfg* c,
                             int* triggered, int* i_stop, int* step_stop, double* xy_err,
                             double* xy_trace, double* P_out, double* K_out, double* R_out);

int stop_oracle_lookahead(const double* mean, const double* sigma, int M,
                          const double* Pvec, const double* Qvec, const double* STMvec, const double* Hvec,
                          const double* pos, const stop_cfg* c,
                          int* triggered, int* i_stop, int* step_stop, double* xy_err,
                          double* xy_trace, double* P_out) {
  return stop_oracle_lookahead_ex(mean, sigma, M, Pvec, Qvec, STMvec, Hvec, pos, c, triggered, i_stop, step_stop, xy_err,
                                  xy_trace, P_out, 0, 0);
}

/* the same, also returning the reference's public members K_pred (15 x 4) and R_IP (4 x 4) of the last update */
int stop_oracle_lookahead_ex(const double* mean, const double* sigma, int M,
                             const double* Pvec, const double* Qvec, const double* STMvec, const double* Hvec,
                             const double* pos, const stop_cfg* c,
                             int* triggered, int* i_stop, int* step_stop, double* xy_err,
                             double* xy_trace, double* P_out, double* K_out, double* R_out) {
  double P[15][15], Q[15][15], F[15][15], H[4][15];
  for (int r = 0; r < 15; ++r)
    for (int col = 0; col < 15; ++col) {
      P[r][col] = Pvec[r * 15 + col];
      Q[r][col] = Qvec[r * 15 + col];
      F[r][col] = STMvec[r * 15 + col];
    }
  for (int r = 0; r < 4; ++r)
    for (int col = 0; col < 15; ++col) H[r][col] = Hvec[c->fix_h_packing ? r * 15 + col : r * 4 + col];
  double enu0[3], enu3[3], R[4][4];
  stop_oracle_llh_to_enu(pos[0], pos[1], pos[2], c, enu0);
  int i = 0, trig = 0, step = c->ratio * M;
  double xy = 0.0;
  for (int slip_i = 0; slip_i < c->ratio * M; ++slip_i) {
    propagate(P, F, Q);
    if (slip_i % c->ratio == 0) {
      stop_oracle_ut_R(mean[i], sigma[i], c, R);
      joseph_update(P, H, R, K_out);
      if (R_out)
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) R_out[a * 4 + b] = R[a][b];
      i++;
    }
    stop_oracle_llh_to_enu(pos[0] + 3.0 * sqrt(fabs(P[6][6])), pos[1] + 3.0 * sqrt(fabs(P[7][7])),
                           pos[2] + 3.0 * sqrt(fabs(P[8][8])), c, enu3);
    double dx = enu3[0] - enu0[0], dy = enu3[1] - enu0[1];
    xy = sqrt(dx * dx + dy * dy);
    if (xy_trace) xy_trace[slip_i] = xy;
    if (xy > c->thresh) { trig = 1; step = slip_i; break; }
  }
  *triggered = trig; *i_stop = i; *step_stop = step; *xy_err = xy;
  if (P_out)
    for (int r = 0; r < 15; ++r)
      for (int col = 0; col < 15; ++col) P_out[r * 15 + col] = P[r][col];
  return 0;
}

/* batch driver for the CPU baseline: B windows.  per_window is a bit mask saying which context arrays carry one
 * entry per window (bit0 P, bit1 Q, bit2 STM, bit3 Hvec, bit4 pos); the others are shared by all windows. */
int stop_oracle_lookahead_batch(const double* mean, const double* sigma, int B, int M,
                                const double* Pvec, const double* Qvec, const double* STMvec, const double* Hvec,
                                const double* pos, int per_window, const stop_cfg* c,
                                int* triggered, int* i_stop, int* step_stop, double* xy_err) {
  for (long b = 0; b < B; ++b) {
    stop_oracle_lookahead(mean + b * M, sigma + b * M, M,
                          Pvec + ((per_window & 1) ? b * 225 : 0), Qvec + ((per_window & 2) ? b * 225 : 0),
                          STMvec + ((per_window & 4) ? b * 225 : 0), Hvec + ((per_window & 8) ? b * 60 : 0),
                          pos + ((per_window & 16) ? b * 3 : 0), c,
                          triggered + b, i_stop + b, step_stop + b, xy_err + b, 0, 0);
  }
  return 0;
}
