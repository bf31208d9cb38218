"""CPU: known answers that pin oracle/ekf_oracle.py (the numpy restatement of CoreNav::insErrorStateModel_LNF / calc_Q,
CoreNav.cpp:411-527, and of the H packing of CoreNav.cpp:669-673)."""
import numpy as np

from corenav_gp_b200 import synthetic as syn
from oracle import ekf_oracle as eo


def one(i=0, dt=0.02):
    p = syn.operating_points(i, 1)
    return eo.context_one(p["llh"][0], p["vel"][0], p["att"][0], p["f_ib_b"][0], dt)


def test_block_structure_of_stm():
    S, _, _, _ = one()
    S = S.reshape(15, 15)
    assert np.array_equal(S[9:, :9], np.zeros((6, 9))) and np.array_equal(S[9:, 9:], np.eye(6))   # bias states are constant
    assert np.array_equal(S[6:9, 0:3], np.zeros((3, 3))) and np.array_equal(S[0:3, 9:12], np.zeros((3, 3)))
    assert np.array_equal(S[0:3, 12:15], S[3:6, 9:12])                                              # both are C_b^n dt
    C = S[0:3, 12:15] / 0.02
    np.testing.assert_allclose(C @ C.T, np.eye(3), atol=1e-14)                                      # a rotation


def test_stm_is_identity_plus_f_dt():
    """STM(dt) - I is exactly linear in dt: (STM(2 dt) - I) = 2 (STM(dt) - I), and STM(dt -> 0) -> I."""
    S1, S2 = one(dt=0.01)[0].reshape(15, 15), one(dt=0.02)[0].reshape(15, 15)
    np.testing.assert_allclose(S2 - np.eye(15), 2.0 * (S1 - np.eye(15)), rtol=1e-12, atol=1e-15)   # (1 + x dt) - 1 on the diagonal
    np.testing.assert_allclose(one(dt=1e-12)[0].reshape(15, 15), np.eye(15), atol=1e-9)


def test_attitude_block_is_minus_skew_of_transport_plus_earth_rate():
    p = syn.operating_points(4, 1)
    llh, vel = p["llh"][0], p["vel"][0]
    S = eo.context_one(llh, vel, p["att"][0], p["f_ib_b"][0], 0.02)[0].reshape(15, 15)
    R_N, R_E = syn.radii(llh[0])
    w = np.array([vel[1] / (R_E + llh[2]) + syn.OMEGA_IE * np.cos(llh[0]), -vel[0] / (R_N + llh[2]),
                  -vel[1] * np.tan(llh[0]) / (R_E + llh[2]) - syn.OMEGA_IE * np.sin(llh[0])])
    F11 = (S[0:3, 0:3] - np.eye(3)) / 0.02
    np.testing.assert_allclose(F11, -syn._skew(w), rtol=1e-9, atol=1e-16)       # (S - I)/dt loses digits: |F11| ~ 1e-4
    assert np.allclose(F11, -F11.T, atol=1e-15)


def test_q_is_symmetric_except_for_the_reference_quirk_and_psd():
    """CoreNav.cpp:512 codes Q52 = F21^T Cbn^T = (Cbn F21)^T while Q25 = F21 Cbn (:503), so the velocity / gyro-bias
    blocks of the reference's Q are not transposes of each other; everything else is symmetric.  The quirk is kept."""
    for i in range(5):
        Q = one(i)[1].reshape(15, 15)
        D = Q - Q.T
        D[3:6, 12:15] = 0.0
        D[12:15, 3:6] = 0.0
        assert np.abs(D).max() <= 1e-12 * np.abs(Q).max()
        assert np.abs(Q[3:6, 12:15] - Q[12:15, 3:6].T).max() > 0.0
        w = np.linalg.eigvalsh((Q + Q.T) / 2)
        assert w.min() > -1e-12 * w.max()
        assert np.array_equal(Q[0:3, 9:12], np.zeros((3, 3))) and np.array_equal(Q[9:12, 12:15], np.zeros((3, 3)))
        d = np.diag(Q)
        assert d[9] == d[10] == d[11] and d[12] == d[13] == d[14]               # bias random walks are isotropic


def test_q_known_values():
    """Q44 = Sbad dt I and Q55 = Sbgd dt I reduce to the squared in-run bias instabilities (CoreNav.cpp:480-491)."""
    Q = one()[1].reshape(15, 15)
    assert np.isclose(Q[9, 9], (3.2e-6 * 9.80665) ** 2, rtol=1e-14)
    assert np.isclose(Q[12, 12], (1.6 * syn.PI_INS / 180 / 3600) ** 2, rtol=1e-14)


def test_h_packing_aliases_as_the_reference_does():
    _, _, Hv, H = one(2)
    # HvecData[r*4+c] with r outer, c inner (CoreNav.cpp:669-673): row r overwrites entries 4r .. 4r+10 of row r-1
    assert np.array_equal(Hv[0:4], H[0, 0:4])
    assert np.array_equal(Hv[4:8], H[1, 0:4])
    assert np.array_equal(Hv[8:12], H[2, 0:4])
    assert np.array_equal(Hv[12:27], H[3, :])
    assert np.array_equal(Hv[27:], np.zeros(33))
    # reading it back the reference's way (gp_predictor.cpp:41-46) gives the effective H of SURVEY App. B q1
    Heff = np.array([[Hv[r * 4 + c] for c in range(15)] for r in range(4)])
    assert np.array_equal(Heff[3], H[3]) and not np.array_equal(Heff[0], H[0])
