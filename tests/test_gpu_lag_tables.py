"""The lag-table path of the batched predictor (csrc/cngp_common.cuh "Lag tables").

The reference's stamps are integer update counts (CoreNav.cpp:286) and its grid has step 1 (gp_slip_node.py:45); for a
stationary expression the covariance then depends on the integer lag only, and phase A builds one table per window that
replaces every evaluation of K(X,X) and K(X,X*).  These tests hold the table path to the oracle AND to the lazy
evaluators (CNGP_NO_LAG_TABLES=1) on the shapes that exercise its edges: gaps, duplicates, negative and large stamps,
grids that overlap the training stamps, spans that do not fit the table, mixed batches, padded windows."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-9


def relerr(a, ref):
    """max |a - ref| / max(|ref|, 1e-3 scale) with scale = max |ref| (true relative error away from zero crossings)."""
    a, ref = np.asarray(a, dtype=float), np.asarray(ref, dtype=float)
    scale = max(float(np.max(np.abs(ref))), 1e-300)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1e-3 * scale)))


def check(gp_ctx, kname, th, x, y, xs, monkeypatch, expect_same_as_lazy=1e-10):
    B = x.shape[0]
    monkeypatch.delenv("CNGP_NO_LAG_TABLES", raising=False)
    mean, var, lml, status = gp_ctx.predict(kname, th, x, y, xs)
    monkeypatch.setenv("CNGP_NO_LAG_TABLES", "1")
    mean0, var0, lml0, status0 = gp_ctx.predict(kname, th, x, y, xs)
    monkeypatch.delenv("CNGP_NO_LAG_TABLES", raising=False)
    assert np.array_equal(status, status0) and np.all(status == 0)
    e = go.KernelExpr(kname)
    for b in range(B):
        thb = th if th.ndim == 1 else th[b]
        xsb = xs if xs.ndim == 1 else xs[b]
        inf = go.inference(e, thb[:-1], thb[-1], x[b], y[b])
        mu, v = go.predict(e, thb[:-1], thb[-1], x[b], y[b], xsb, inf)
        assert relerr(mean[b], mu) < TOL and relerr(var[b], v) < TOL and relerr(lml[b], inf.lml) < TOL, (kname, b)
    assert relerr(mean, mean0) < expect_same_as_lazy and relerr(var, var0) < expect_same_as_lazy
    assert relerr(lml, lml0) < expect_same_as_lazy


@pytest.mark.parametrize("kname,theta", [
    ("rbf", [0.01, 10.0]),
    ("rbf+stdperiodic", [0.01, 10.0, 0.0025, 37.0, 1.0]),
    ("mat32+bias", [0.02, 7.0, 0.001]),
    ("mat52+white", [0.5, 15.0, 1e-4]),
    ("ratquad+stdperiodic*rbf", [0.01, 8.0, 1.5, 0.5, 37.0, 1.2, 0.01, 50.0]),
])
@pytest.mark.parametrize("N,M", [(256, 600), (100, 600), (37, 13), (255, 601)])
def test_table_path_matches_oracle_and_lazy_path(gp_ctx, monkeypatch, kname, theta, N, M):
    x, y = syn.slip_windows(21, 3, N)
    xs = syn.test_grid(x[0], M)
    check(gp_ctx, kname, np.array(theta + [1e-3]), x, y, xs, monkeypatch)


def test_reference_grid_overlapping_the_training_stamps(gp_ctx, monkeypatch):
    """gp_slip_node.py:45: X_ = arange(X.min(), X.max() + 600, 1) starts INSIDE the training span (lag 0 and negative
    lags), per-window grids, per-window hyper-parameters."""
    B, N = 4, 134
    x, y = syn.slip_windows(5, B, N)
    x = x + 7.0 * np.arange(B)[:, None]
    xs = np.stack([np.arange(x[b].min(), x[b].max() + 600.0, 1.0) for b in range(B)])
    th = np.array([0.01, 10.0, 0.0025, 37.0, 1.0, 1e-3])[None, :] * (1.0 + 0.05 * np.arange(B))[:, None]
    check(gp_ctx, "rbf+stdperiodic", th, x, y, xs, monkeypatch)


def test_gaps_duplicates_negative_and_large_stamps(gp_ctx, monkeypatch):
    rng = np.random.default_rng(3)
    N, M = 96, 50
    base = np.cumsum(rng.integers(0, 4, N)).astype(float)        # gaps of 0..3 counts: duplicate stamps included
    x = np.stack([base, base - 500.0, base + 3.0e6])
    y = 0.05 * np.sin(base / 9.0)[None, :] + 0.01 * rng.standard_normal((3, N))
    xs = np.stack([np.arange(M) + x[b].max() + 1 for b in range(3)])
    # duplicates make Ky singular up to the noise: keep the noise at 1e-2 so the comparison is well conditioned
    check(gp_ctx, "rbf+mat32", np.array([0.01, 10.0, 0.005, 20.0, 1e-2]), x, y, xs, monkeypatch)


def test_span_beyond_the_table_falls_back(gp_ctx, monkeypatch):
    N, M = 40, 16
    x = np.stack([np.arange(N) * 60.0, np.arange(N) * 1.0])       # window 0: span 2340 > 1024 lags
    y = 0.05 * np.cos(np.arange(N) / 5.0)[None, :] * np.ones((2, 1))
    xs = np.stack([x[b].max() + 1 + np.arange(M) for b in range(2)])
    check(gp_ctx, "rbf", np.array([0.01, 100.0, 1e-3]), x, y, xs, monkeypatch)


def test_mixed_batch_integer_and_fractional_windows(gp_ctx, monkeypatch):
    B, N, M = 6, 128, 600
    x, y = syn.slip_windows(9, B, N)
    x[1] += 0.5
    x[4] += 0.125
    xs = np.stack([syn.test_grid(x[b], M) for b in range(B)])
    check(gp_ctx, "rbf+stdperiodic", syn.theta_for("rbf+stdperiodic"), x, y, xs, monkeypatch, expect_same_as_lazy=1e-10)


def test_fractional_grid_disables_the_table(gp_ctx, monkeypatch):
    B, N, M = 2, 64, 33
    x, y = syn.slip_windows(2, B, N)
    xs = syn.test_grid(x[0], M) + 0.25
    check(gp_ctx, "rbf", syn.theta_for("rbf"), x, y, xs, monkeypatch, expect_same_as_lazy=1e-10)


def test_lml_sweep_uses_the_table_for_the_factorisation(gp_ctx, monkeypatch):
    B, N, C = 3, 256, 3
    x, y = syn.slip_windows(7, B, N)
    th = syn.theta_for("rbf+stdperiodic")[None, :] * np.array([[1.0], [1.3], [0.75]])
    lml, grad, status = gp_ctx.lml_grad("rbf+stdperiodic", th, x, y)
    monkeypatch.setenv("CNGP_NO_LAG_TABLES", "1")
    lml0, grad0, status0 = gp_ctx.lml_grad("rbf+stdperiodic", th, x, y)
    assert relerr(lml, lml0) < 1e-12 and relerr(grad, grad0) < 1e-9
    e = go.KernelExpr("rbf+stdperiodic")
    for c in range(C):
        for b in range(B):
            inf = go.inference(e, th[c, :-1], th[c, -1], x[b], y[b], want_grad=True)
            assert relerr(lml[c, b], inf.lml) < TOL
