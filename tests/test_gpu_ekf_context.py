"""GPU parity of cngp_ekf_context_batch (STM / Q / packed H behind SetStopping; CoreNav.cpp:411-527) against the numpy
oracle: 1e-11 relative per matrix (CUDA vs glibc sin/cos/tan/pow differ in the last ulp; exact zeros stay exact), and
the generated contexts drive the look-ahead to the same ZUPT decisions."""
import numpy as np
import pytest
import torch

from corenav_gp_b200 import synthetic as syn
from oracle import ekf_oracle as eo
from oracle import stop_oracle as so

pytestmark = pytest.mark.gpu


def close(got, ref, atol=0.0):
    """1e-11 relative on every entry, structural zeros exact.  atol: H holds entries that are pure cancellation residue
    (row0(C) x v with v along the body x axis: ~1e-17 from O(1) products), compared absolutely."""
    assert np.array_equal(got == 0.0, ref == 0.0), "structural zeros must be exact"
    err = np.abs(got - ref) - atol
    assert (err <= 1e-11 * np.abs(ref)).all(), (err / np.maximum(np.abs(ref), 1e-300)).max()


@pytest.mark.parametrize("B", [1, 257])
def test_matches_oracle(gp_ctx, B):
    p = syn.operating_points(0, B)
    S, Q, H = gp_ctx.ekf_context(p["llh"], p["vel"], p["att"], p["f_ib_b"], dt=0.02)
    rS, rQ, rH = eo.context(p["llh"], p["vel"], p["att"], p["f_ib_b"], dt=0.02)
    close(S, rS); close(Q, rQ); close(H, rH, atol=1e-15)


def test_device_buffers_other_steps_and_no_h(gp_ctx):
    p = syn.operating_points(500, 64)
    dev = {k: torch.from_numpy(v).cuda() for k, v in p.items()}
    S, Q, H = gp_ctx.ekf_context(dev["llh"], dev["vel"], dev["att"], dev["f_ib_b"], dt=0.005, dt_odo=0.05, want_h=False)
    assert H is None
    rS, rQ, _ = eo.context(p["llh"], p["vel"], p["att"], p["f_ib_b"], dt=0.005, dt_odo=0.05)
    close(S.cpu().numpy(), rS); close(Q.cpu().numpy(), rQ)


def test_generated_contexts_reproduce_zupt_decisions(gp_ctx):
    """N4 -> row a9-a12: look-ahead on device-generated contexts = C oracle on oracle-generated contexts."""
    B, M = 48, 200
    p = syn.operating_points(100, B)
    S, Q, H = gp_ctx.ekf_context(p["llh"], p["vel"], p["att"], p["f_ib_b"])
    rS, rQ, rH = eo.context(p["llh"], p["vel"], p["att"], p["f_ib_b"])
    rng = np.random.default_rng(0)
    k = np.arange(M)
    mean = 0.05 * np.exp(-k / 80.0)[None, :] * rng.uniform(-1, 1, (B, 1))
    sigma = 2.0 * np.sqrt(1e-3 + 0.01 * (1 - np.exp(-k / 150.0)))[None, :] * rng.uniform(0.7, 1.3, (B, 1))
    P0 = syn.lookahead_context(syn.window_sigmas(0, B, lo=0.3, hi=0.9))["P"]
    pos = np.tile(syn.INIT_LLH, (B, 1))
    out = gp_ctx.zupt_lookahead(mean, sigma, P0, Q, S, H, pos)
    ref = so.lookahead_batch(mean, sigma, P0, rQ, rS, rH, pos)
    assert 0 < ref["triggered"].sum()
    for key in ("triggered", "i_stop", "step_stop"):
        assert np.array_equal(out[key], ref[key]), key


def test_ekf_covariance_recursion_matches_numpy(gp_ctx):
    """cngp_ekf_covariance_batch (row N4, the P half): P <- STM P STM' + Q every IMU step, Joseph-form odometry update
    every 5th (CoreNav.cpp:101, 226-230), against a numpy transcription with matrix products (different summation
    order: 1e-11)."""
    B, n_steps = 5, 750                              # one 150-count recording window at 5 IMU steps per odometry update
    c = syn.lookahead_context(syn.window_sigmas(0, B))
    P0 = c["P"]
    F, Q, H = c["STM"].reshape(15, 15), c["Q"].reshape(15, 15), c["H"]
    R = np.diag([0.05 ** 2, 0.1 ** 2, 0.03 ** 2, 0.03 ** 2])
    out = gp_ctx.ekf_covariance(P0, c["Q"], c["STM"], H.reshape(60), R.reshape(16), n_steps)
    for b in range(B):
        P = P0[b].reshape(15, 15).copy()
        for k in range(n_steps):
            P = F @ P @ F.T + Q
            if k % 5 == 0:
                K = P @ H.T @ np.linalg.inv(H @ P @ H.T + R)
                A = np.eye(15) - K @ H
                P = A @ P @ A.T + K @ R @ K.T
        got = out[b].reshape(15, 15)
        assert np.max(np.abs(got - P)) / np.max(np.abs(P)) < 1e-10
    # per-window R and Q
    Rb = np.stack([R.reshape(16) * (1 + 0.1 * b) for b in range(B)])
    Qb = np.stack([c["Q"] * (1 + 0.2 * b) for b in range(B)])
    out2 = gp_ctx.ekf_covariance(P0, Qb, c["STM"], H.reshape(60), Rb, 50)
    P = P0[3].reshape(15, 15).copy()
    Q3, R3 = Qb[3].reshape(15, 15), Rb[3].reshape(4, 4)
    for k in range(50):
        P = F @ P @ F.T + Q3
        if k % 5 == 0:
            K = P @ H.T @ np.linalg.inv(H @ P @ H.T + R3)
            A = np.eye(15) - K @ H
            P = A @ P @ A.T + K @ R3 @ K.T
    assert np.max(np.abs(out2[3].reshape(15, 15) - P)) / np.max(np.abs(P)) < 1e-10
