// Batched EKF context generation: the state-transition matrix, process-noise covariance
// and odometry measurement matrix that CoreNav hands to the stop predictor through SetStopping (row a9), for B
// independent operating points - what a Monte-Carlo run of the look-ahead needs on the device instead of B service
// round trips.  Replaces, per window, the pure functions
//     CoreNav::insErrorStateModel_LNF   core_navigation/src/CoreNav.cpp:411-470   -> STM  (15x15, row-major)
//     CoreNav::calc_Q                   core_navigation/src/CoreNav.cpp:471-527   -> Q    (15x15, row-major)
// evaluated as CoreNav::Propagate does (radii of curvature :63-66, transport and Earth rate :60,68-72), with
// C_b^n = eul_to_dcm(att)^T (:560-581), and the 4x15 odometry H of CoreNav.cpp:191-220 in its instantaneous form
// (the time-averaged integrals of :116-122 replaced by their integrands), packed with the aliasing index
// HvecData[r*4+c] of CoreNav::setStopping_ (:669-673).  The reference quirks are kept: the geocentric
// latitude is formed from lat*180/PI inside sin/cos (:412), F23(1,2) uses lon*h (:437).
//
// One thread per operating point; the inputs are 12 doubles and the outputs 510 doubles per window (4.2 KB), so the
// kernel is bound by its HBM writes.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cngp.h"

namespace {

constexpr double OMEGA_IE = 7.292115e-5, R0 = 6378137.0, FLAT = 1.0 / 298.257223563, ECC = 0.0818191909425;
constexpr double PI_INS = 3.14159265358979;       // InsConst.h:20 (truncated on purpose)

struct M3 { double m[3][3]; };

__device__ __forceinline__ M3 skew(const double v[3]) {
  M3 s{{{0.0, -v[2], v[1]}, {v[2], 0.0, -v[0]}, {-v[1], v[0], 0.0}}};
  return s;
}
__device__ __forceinline__ M3 mul(const M3& a, const M3& b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return c;
}
__device__ __forceinline__ M3 tr(const M3& a) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[j][i];
  return c;
}
// out[(3 br + i) * 15 + 3 bc + j] = s * a(i,j) (+ 1 on the block diagonal when `eye`)
__device__ __forceinline__ void put(double* out, int br, int bc, const M3& a, double s, bool eye = false) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) out[(3 * br + i) * 15 + 3 * bc + j] = s * a.m[i][j] + ((eye && i == j) ? 1.0 : 0.0);
}
__device__ __forceinline__ void put_t(double* out, int br, int bc, const M3& a, double s) { put(out, br, bc, tr(a), s); }

__global__ void __launch_bounds__(128) ekf_context_kernel(const double* __restrict__ llh, const double* __restrict__ vel,
                                                          const double* __restrict__ att, const double* __restrict__ fib,
                                                          long long B, double dt, double dt_odo, double* __restrict__ STM,
                                                          double* __restrict__ Q, double* __restrict__ Hvec) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double lat = llh[3 * b], lon = llh[3 * b + 1], h = llh[3 * b + 2];
  const double v[3] = {vel[3 * b], vel[3 * b + 1], vel[3 * b + 2]};
  const double f[3] = {fib[3 * b], fib[3 * b + 1], fib[3 * b + 2]};
  const double phi = att[3 * b], the = att[3 * b + 1], psi = att[3 * b + 2];

  // nav -> body DCM (c3 c2) c1 and its transpose C_b^n
  double sphi, cphi, sthe, cthe, spsi, cpsi;
  sincos(phi, &sphi, &cphi); sincos(the, &sthe, &cthe); sincos(psi, &spsi, &cpsi);
  const M3 c1{{{cpsi, spsi, 0.0}, {-spsi, cpsi, 0.0}, {0.0, 0.0, 1.0}}};
  const M3 c2{{{cthe, 0.0, -sthe}, {0.0, 1.0, 0.0}, {sthe, 0.0, cthe}}};
  const M3 c3{{{1.0, 0.0, 0.0}, {0.0, cphi, sphi}, {0.0, -sphi, cphi}}};
  const M3 Cnb = mul(mul(c3, c2), c1);
  const M3 Cbn = tr(Cnb);

  double slat, clat;
  sincos(lat, &slat, &clat);
  const double tlat = tan(lat);
  const double e2s2 = 1.0 - ECC * ECC * slat * slat;
  const double R_N = R0 * (1.0 - ECC * ECC) / pow(e2s2, 1.5);
  const double R_E = R0 / sqrt(e2s2);
  const double rn = R_N + h, re = R_E + h;
  const double w_ie[3] = {OMEGA_IE * clat, 0.0, -OMEGA_IE * slat};
  const double w_en[3] = {v[1] / re, -v[0] / rn, -v[1] * tlat / re};
  const double w_in[3] = {w_en[0] + w_ie[0], w_en[1] + w_ie[1], w_en[2] + w_ie[2]};

  // ---- state transition ----
  const double tc = (1.0 - FLAT) * (1.0 - FLAT);
  const double latdeg = lat * 180.0 / PI_INS;
  const double geo = atan2(tc * sin(latdeg), cos(latdeg));
  const double sg = sin(geo);
  const double r_geo = sqrt(R0 * R0 / (1.0 + (1.0 / tc - 1.0) * sg * sg));
  const double s2l = sin(2.0 * lat);
  const double g0 = 9.780318 * (1.0 + 5.3024e-3 * slat * slat - 5.9e-6 * s2l * s2l);
  const double sec2 = (1.0 / clat) * (1.0 / clat);

  M3 F11 = skew(w_in);
  const M3 F12{{{0.0, -1.0 / re, 0.0}, {1.0 / rn, 0.0, 0.0}, {0.0, tlat / re, 0.0}}};
  const M3 F13{{{OMEGA_IE * slat, 0.0, v[1] / (re * re)},
                {0.0, 0.0, -v[0] / (rn * rn)},
                {OMEGA_IE * clat + v[1] / (re * (clat * clat)), 0.0, -v[1] * tlat / (re * re)}}};
  const double cf[3] = {Cbn.m[0][0] * f[0] + Cbn.m[0][1] * f[1] + Cbn.m[0][2] * f[2],
                        Cbn.m[1][0] * f[0] + Cbn.m[1][1] * f[1] + Cbn.m[1][2] * f[2],
                        Cbn.m[2][0] * f[0] + Cbn.m[2][1] * f[1] + Cbn.m[2][2] * f[2]};
  M3 F21 = skew(cf);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { F11.m[i][j] = -F11.m[i][j]; F21.m[i][j] = -F21.m[i][j]; }
  const M3 F22{{{v[2] / rn, -(2.0 * v[1] * tlat / re) - 2.0 * OMEGA_IE * slat, v[0] / rn},
                {v[1] * tlat / re + 2.0 * OMEGA_IE * slat, (v[0] * tlat + v[2]) / re, v[1] / re + 2.0 * OMEGA_IE * clat},
                {-2.0 * v[0] / rn, -2.0 * (v[1] / re) - 2.0 * OMEGA_IE * clat, 0.0}}};
  const M3 F23{{{-(v[1] * v[1] * sec2 / re) - 2.0 * v[1] * OMEGA_IE * clat, 0.0,
                 v[1] * v[1] * tlat / (re * re) - v[0] * v[2] / (rn * rn)},
                {(v[0] * v[1] * sec2 / re) + 2.0 * v[0] * OMEGA_IE * clat - 2.0 * v[2] * OMEGA_IE * slat, 0.0,
                 -((v[0] * v[1] * tlat + lon * h) / (re * re))},
                {2.0 * v[1] * OMEGA_IE * slat, 0.0, (v[1] * v[1] / (re * re) + v[0] * v[0] / (rn * rn)) - 2.0 * g0 / r_geo}}};
  const M3 T{{{1.0 / rn, 0.0, 0.0}, {0.0, 1.0 / (re * clat), 0.0}, {0.0, 0.0, -1.0}}};   // F32 = T_rn_p
  const M3 F33{{{0.0, 0.0, -v[0] / (rn * rn)},
                {(v[1] * slat) / (re * (clat * clat)), 0.0, -v[1] / ((re * re) * clat)},
                {0.0, 0.0, 0.0}}};
  const M3 Z{{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}};

  double* S = STM + b * 225;
  put(S, 0, 0, F11, dt, true); put(S, 0, 1, F12, dt); put(S, 0, 2, F13, dt); put(S, 0, 3, Z, 0.0); put(S, 0, 4, Cbn, dt);
  put(S, 1, 0, F21, dt); put(S, 1, 1, F22, dt, true); put(S, 1, 2, F23, dt); put(S, 1, 3, Cbn, dt); put(S, 1, 4, Z, 0.0);
  put(S, 2, 0, Z, 0.0); put(S, 2, 1, T, dt); put(S, 2, 2, F33, dt, true); put(S, 2, 3, Z, 0.0); put(S, 2, 4, Z, 0.0);
  put(S, 3, 0, Z, 0.0); put(S, 3, 1, Z, 0.0); put(S, 3, 2, Z, 0.0); put(S, 3, 3, Z, 0.0, true); put(S, 3, 4, Z, 0.0);
  put(S, 4, 0, Z, 0.0); put(S, 4, 1, Z, 0.0); put(S, 4, 2, Z, 0.0); put(S, 4, 3, Z, 0.0); put(S, 4, 4, Z, 0.0, true);

  // ---- process noise (Groves 14.2.6 as coded at CoreNav.cpp:479-519) ----
  const double gg = 9.80665;
  const double sig_gyro = 1.6 * PI_INS / 180 / 3600, sig_arw = .1 * (PI_INS / 180) * sqrt(3600.0) / 3600;
  const double sig_acc = 3.2e-6 * gg, sig_vrw = 0.008 * sqrt(3600.0) / 3600;
  const double Srg = sig_arw * sig_arw * dt, Sra = sig_vrw * sig_vrw * dt;
  const double Sbad = sig_acc * sig_acc / dt, Sbgd = sig_gyro * sig_gyro / dt;
  const double dt2 = dt * dt, dt3 = dt2 * dt, dt4 = dt3 * dt, dt5 = dt4 * dt, dt6 = dt5 * dt, dt7 = dt6 * dt;
  const M3 I{{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}};
  const M3 FF = mul(F21, tr(F21));
  const M3 TF = mul(T, F21), TFF = mul(T, FF), TT = mul(T, T), TFFT = mul(TFF, T);
  const M3 TC = mul(T, Cbn), FC = mul(F21, Cbn), CF = mul(Cbn, F21), TFC = mul(TF, Cbn);
  const double q21 = 0.5 * Srg * dt2 + 0.25 * Sbgd * dt4, q31 = (1.0 / 3.0) * Srg * dt3 + (1.0 / 5.0) * Sbgd * dt5;
  double* Qo = Q + b * 225;
  put(Qo, 0, 0, I, Srg * dt + (1.0 / 3.0) * Sbgd * dt3); put_t(Qo, 0, 1, F21, q21); put_t(Qo, 0, 2, TF, q31);
  put(Qo, 0, 3, Z, 0.0); put(Qo, 0, 4, Cbn, 0.5 * Sbgd * dt2);
  put(Qo, 1, 0, F21, q21);
  {
    M3 Q22, Q32, Q33;
    const double a22 = Sra * dt + (1.0 / 3.0) * Sbad * dt3, a32 = 0.5 * Sra * dt2 + 0.25 * Sbad * dt4;
    const double b32 = 0.25 * Srg * dt4 + (1.0 / 6.0) * Sbgd * dt6;
    const double a33 = (1.0 / 3.0) * Sra * dt3 + (1.0 / 5.0) * Sbad * dt5, b33 = (1.0 / 5.0) * Srg * dt5 + (1.0 / 7.0) * Sbgd * dt7;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        Q22.m[i][j] = a22 * I.m[i][j] + q31 * FF.m[i][j];
        Q32.m[i][j] = a32 * T.m[i][j] + b32 * TFF.m[i][j];
        Q33.m[i][j] = a33 * TT.m[i][j] + b33 * TFFT.m[i][j];
      }
    put(Qo, 1, 1, Q22, 1.0); put_t(Qo, 1, 2, Q32, 1.0);
    put(Qo, 2, 1, Q32, 1.0); put(Qo, 2, 2, Q33, 1.0);
  }
  put(Qo, 1, 3, Cbn, 0.5 * Sbad * dt2); put(Qo, 1, 4, FC, (1.0 / 3.0) * Sbgd * dt3);
  put(Qo, 2, 0, TF, q31); put(Qo, 2, 3, TC, (1.0 / 3.0) * Sbad * dt3); put(Qo, 2, 4, TFC, 0.25 * Sbgd * dt4);
  put(Qo, 3, 0, Z, 0.0); put_t(Qo, 3, 1, Cbn, 0.5 * Sbad * dt2); put_t(Qo, 3, 2, TC, (1.0 / 3.0) * Sbad * dt3);
  put(Qo, 3, 3, I, Sbad * dt); put(Qo, 3, 4, Z, 0.0);
  put_t(Qo, 4, 0, Cbn, 0.5 * Sbgd * dt2); put_t(Qo, 4, 1, CF, (1.0 / 3.0) * Sbgd * dt3); /* F21^T Cbn^T */ put_t(Qo, 4, 2, TFC, 0.25 * Sbgd * dt4);
  put(Qo, 4, 3, Z, 0.0); put(Qo, 4, 4, I, Sbgd * dt);

  // ---- odometry H (4x15), packed with the reference's aliasing index: later writes win ----
  if (Hvec) {
    const M3 CV = mul(Cnb, skew(v));
    double H[4][15];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 15; ++c) H[r][c] = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H[0][j] = -CV.m[0][j]; H[0][3 + j] = -Cnb.m[0][j];
      H[1][9 + j] = -(cthe * Cbn.m[2][j]) / dt_odo;
      H[2][j] = -CV.m[1][j]; H[2][3 + j] = -Cnb.m[1][j];
      H[3][j] = -CV.m[2][j]; H[3][3 + j] = -Cnb.m[2][j];
    }
    double* Ho = Hvec + b * 60;
    for (int i = 0; i < 60; ++i) Ho[i] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 15; ++c) Ho[r * 4 + c] = H[r][c];
  }
}

}  // namespace

extern "C" int cngp_launch_ekf_context(const double* llh, const double* vel, const double* att, const double* fib,
                                       long long B, double dt, double dt_odo, double* STM, double* Q, double* Hvec,
                                       cudaStream_t s) {
  const long long grid = (B + 127) / 128;
  ekf_context_kernel<<<(unsigned)grid, 128, 0, s>>>(llh, vel, att, fib, B, dt, dt_odo, STM, Q, Hvec);
  return (int)cudaGetLastError();
}
