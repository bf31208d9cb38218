"""CPU tests of the stop-predictor oracle (oracle/stop_oracle.c), the restatement of GpPredictor::GPCallBack
(gp_predictor/src/gp_predictor.cpp:58-130,144-178).  The reference has no test for it ("parity unpinned"), so the C
restatement is pinned against an independent numpy transcription written from the same lines with matrix products
(different summation order => tolerance, not bit equality), and the documented quirks are asserted one by one
(SURVEY.md App. B)."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import stop_oracle as so


def np_llh_to_enu(lat, lon, h, init_llh=syn.INIT_LLH, init_ecef=syn.INIT_ECEF):
    a, b = 6378137.0, 6356752.3142                      # gp_predictor.cpp:150-151 (function-local constants, q9)
    e = np.sqrt(1 - (b / a) ** 2)
    t2 = np.tan(lat) ** 2
    den = np.sqrt(1 + (1 - e * e) * t2)
    x1 = a * np.cos(lon) / den + h * np.cos(lon) * np.cos(lat)
    y1 = a * np.sin(lon) / den + h * np.sin(lon) * np.cos(lat)
    z1 = a * (1 - e * e) * np.sin(lat) / np.sqrt(1 - e * e * np.sin(lat) ** 2) + h * np.sin(lat)
    sp, cp, sl, cl = np.sin(init_llh[0]), np.cos(init_llh[0]), np.sin(init_llh[1]), np.cos(init_llh[1])
    R = np.array([[-sl, cl, 0], [-sp * cl, -sp * sl, cp], [cp * cl, cp * sl, sp]])
    return R @ (np.array([x1, y1, z1]) - init_ecef)


def np_ut_R(mean, sigma, v=0.8, fa=0.03, fb=0.05, track=0.685, scale=25.0):
    chi = v / (1.0 - np.array([mean, mean + sigma, mean - sigma]))     # gp_predictor.cpp:69-75
    est = chi.mean()
    cov = np.mean((chi - est) ** 2)
    R2 = np.diag([max(fa ** 2, cov ** 2), max(fa ** 2, cov ** 2), max(fb ** 2, cov ** 2), fb ** 2])   # :80-83 (q2)
    R1 = np.array([[0.5, 0.5, 0, 0], [1 / track, -1 / track, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    return scale * R1 @ R2 @ R1.T


def np_lookahead(mean, sigma, P, Q, F, Hvec, pos, thresh=3.0, ratio=5, fix_h=False):
    P, Q, F = (np.array(a, dtype=float).reshape(15, 15) for a in (P, Q, F))
    H = np.array([[Hvec[r * 15 + c] if fix_h else Hvec[r * 4 + c] for c in range(15)] for r in range(4)])
    enu0 = np_llh_to_enu(*pos)
    i, trace = 0, []
    for slip_i in range(ratio * len(mean)):
        P = F @ P @ F.T + Q                                                    # :66
        if slip_i % ratio == 0:                                                # :67
            R = np_ut_R(mean[i], sigma[i])
            K = P @ H.T @ np.linalg.inv(H @ P @ H.T + R)                       # :90
            A = np.eye(15) - K @ H
            P = A @ P @ A.T + K @ R @ K.T                                      # :91
            i += 1
        e3 = np_llh_to_enu(pos[0] + 3 * np.sqrt(abs(P[6, 6])), pos[1] + 3 * np.sqrt(abs(P[7, 7])),
                           pos[2] + 3 * np.sqrt(abs(P[8, 8])))                 # :95
        xy = np.hypot(e3[0] - enu0[0], e3[1] - enu0[1])                        # :98-99
        trace.append(xy)
        if xy > thresh:                                                        # :102
            return True, i, slip_i, xy, np.array(trace)
    return False, i, ratio * len(mean), xy, np.array(trace)


def gp_out(M, seed=0):
    rng = np.random.default_rng(seed)
    k = np.arange(M)
    return (0.05 * np.exp(-k / 80.0) * rng.uniform(-1, 1) + 0.01 * rng.standard_normal(M),
            2.0 * np.sqrt(1e-3 + 0.01 * (1 - np.exp(-k / 150.0))))


def test_llh_to_enu_origin_and_offsets():
    assert np.max(np.abs(so.llh_to_enu(*syn.INIT_LLH) - np_llh_to_enu(*syn.INIT_LLH))) < 1e-8
    # init_ecef in config/init_params.yaml:9-12 is the ECEF of init_llh to ~cm, so the origin maps close to (0,0,0)
    assert np.linalg.norm(so.llh_to_enu(*syn.INIT_LLH)) < 1.0
    # one metre north / east / up
    R_N, R_E = syn.radii(syn.INIT_LLH[0])
    d = so.llh_to_enu(syn.INIT_LLH[0] + 1.0 / R_N, syn.INIT_LLH[1], syn.INIT_LLH[2]) - so.llh_to_enu(*syn.INIT_LLH)
    assert abs(d[1] - 1.0) < 1e-3 and abs(d[0]) < 1e-6
    d = so.llh_to_enu(syn.INIT_LLH[0], syn.INIT_LLH[1] + 1.0 / (R_E * np.cos(syn.INIT_LLH[0])), syn.INIT_LLH[2]) \
        - so.llh_to_enu(*syn.INIT_LLH)
    assert abs(d[0] - 1.0) < 1e-3
    d = so.llh_to_enu(syn.INIT_LLH[0], syn.INIT_LLH[1], syn.INIT_LLH[2] + 1.0) - so.llh_to_enu(*syn.INIT_LLH)
    assert abs(d[2] - 1.0) < 1e-6


@pytest.mark.parametrize("mean,sigma", [(0.0, 0.1), (0.05, 0.3), (-0.1, 0.02), (0.3, 0.9)])
def test_ut_R_matches_numpy(mean, sigma):
    assert np.allclose(so.ut_R(mean, sigma), np_ut_R(mean, sigma), rtol=1e-13, atol=0)


def test_ut_R_floors_and_double_square():
    R = so.ut_R(0.0, 1e-6)                  # tiny sigma: all three floors active
    assert np.allclose(R, np_ut_R(0.0, 1e-6)) and R[3, 3] == pytest.approx(25 * 0.05 ** 2)
    # q2: the UT "covariance" is squared again inside max(floor^2, c^2)
    chi = 0.8 / (1 - np.array([0.2, 0.9, -0.5]))
    c = np.mean((chi - chi.mean()) ** 2)
    assert so.ut_R(0.2, 0.7)[2, 2] == pytest.approx(25 * max(0.05 ** 2, c * c))


@pytest.mark.parametrize("s0,M", [(0.5, 120), (0.3, 200), (0.8, 40)])
def test_lookahead_matches_numpy_transcription(s0, M):
    c = syn.lookahead_context(s0)
    mean, sigma = gp_out(M, seed=int(s0 * 10))
    ref = np_lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"])
    out = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], want_trace=True)
    assert (out["triggered"], out["i_stop"], out["step_stop"]) == ref[:3]
    n = ref[4].size
    assert np.max(np.abs(out["xy_trace"][:n] - ref[4]) / np.maximum(1.0, ref[4])) < 1e-7


def test_lookahead_quirks():
    c = syn.lookahead_context(0.5)
    mean, sigma = gp_out(50, seed=1)
    # q5: the first error check already includes one propagation and one update => i_stop >= 1 even at step 0
    out = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], so.default_cfg(thresh=1e-6))
    assert out["triggered"] and out["step_stop"] == 0 and out["i_stop"] == 1
    # q8: no trigger inside the horizon => triggered False, i_stop = M, step_stop = ratio*M
    out = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], so.default_cfg(thresh=1e9))
    assert not out["triggered"] and out["i_stop"] == 50 and out["step_stop"] == 250
    # q1: the aliasing H index changes the result relative to the intended row-major packing
    hv_true = c["H"].reshape(60)
    a = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], so.default_cfg(thresh=1e9))
    b = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], hv_true, c["pos"], so.default_cfg(thresh=1e9, fix_h_packing=1))
    assert abs(a["xy_err"] - b["xy_err"]) > 1e-6
    # packed vector: only indices 0..26 are ever written by CoreNav::setStopping_ (CoreNav.cpp:669-673)
    assert np.all(c["Hvec"][27:] == 0.0)


def test_lookahead_batch_equals_single():
    B, M = 6, 30
    rng = np.random.default_rng(2)
    mean = 0.02 * rng.standard_normal((B, M))
    sigma = 0.2 + 0.1 * rng.random((B, M))
    c = syn.lookahead_context(syn.window_sigmas(0, B))
    ref = so.lookahead_batch(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"])
    for b in range(B):
        one = so.lookahead(mean[b], sigma[b], c["P"][b], c["Q"], c["STM"], c["Hvec"], c["pos"])
        assert one["triggered"] == bool(ref["triggered"][b]) and one["i_stop"] == ref["i_stop"][b]
        assert one["xy_err"] == ref["xy_err"][b]


def test_synthetic_context_is_sane():
    c = syn.lookahead_context(0.5)
    F, Q = c["STM"].reshape(15, 15), c["Q"].reshape(15, 15)
    assert np.allclose(np.diag(F), 1.0, atol=1e-3) and np.allclose(Q, Q.T, atol=1e-25)
    assert np.all(np.linalg.eigvalsh(Q) > -1e-20)
    s = syn.window_sigmas(0, 1000)
    assert 0.2 <= s.min() and s.max() <= 0.8 and abs(s.mean() - 0.5) < 0.03
    # windows are reproducible shard by shard
    a = syn.slip_windows(100, 8, 64)[1]
    b = np.concatenate([syn.slip_windows(100, 3, 64)[1], syn.slip_windows(103, 5, 64)[1]])
    assert np.array_equal(a, b)
