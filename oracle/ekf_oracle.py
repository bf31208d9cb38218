"""CPU checker of cngp_ekf_context_batch: STM / Q / packed H of the SetStopping service for B operating points.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference has no test for CoreNav::insErrorStateModel_LNF / calc_Q
(core_navigation/src/CoreNav.cpp:411-527) and cannot be built here (ROS + Eigen).  The arithmetic is the numpy
restatement that generates the benchmark contexts (corenav_gp_b200/synthetic.py: stm_lnf, calc_q, eul_to_dcm,
pack_hvec - input generation, no GP arithmetic); tests/test_oracle_ekf.py pins it on closed-form known answers
(Q symmetric PSD and block structure, STM -> I as dt -> 0, STM = I + F dt against a finite difference of the
analytic F, the H aliasing pattern of SURVEY App. B q1)."""
from __future__ import annotations

import numpy as np

from corenav_gp_b200 import synthetic as syn


def context_one(llh, vel, att, f_ib_b, dt=0.02, dt_odo=0.1):
    C = syn.eul_to_dcm(*att)          # nav -> body (CoreNav.cpp:560-581)
    Cbn = C.T
    R_N, R_E = syn.radii(llh[0])
    w_ie = np.array([syn.OMEGA_IE * np.cos(llh[0]), 0.0, -syn.OMEGA_IE * np.sin(llh[0])])            # CoreNav.cpp:60
    w_en = np.array([vel[1] / (R_E + llh[2]), -vel[0] / (R_N + llh[2]), -vel[1] * np.tan(llh[0]) / (R_E + llh[2])])  # :68-70
    STM = syn.stm_lnf(llh, vel, dt, Cbn, w_en + w_ie, f_ib_b)
    Q = syn.calc_q(llh, dt, Cbn, f_ib_b)
    vss = syn._skew(vel)
    H = np.zeros((4, 15))
    H[0, 0:3] = -(C @ vss)[0]; H[0, 3:6] = -C[0]
    H[1, 9:12] = -(np.cos(att[1]) * C.T[2]) / dt_odo
    H[2, 0:3] = -(C @ vss)[1]; H[2, 3:6] = -C[1]
    H[3, 0:3] = -(C @ vss)[2]; H[3, 3:6] = -C[2]
    return STM.reshape(225), Q.reshape(225), syn.pack_hvec(H), H


def context(llh, vel, att, f_ib_b, dt=0.02, dt_odo=0.1):
    B = len(llh)
    S, Q, Hv = np.zeros((B, 225)), np.zeros((B, 225)), np.zeros((B, 60))
    for b in range(B):
        S[b], Q[b], Hv[b], _ = context_one(llh[b], vel[b], att[b], f_ib_b[b], dt, dt_odo)
    return S, Q, Hv
