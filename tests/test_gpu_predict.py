"""GPU parity of cngp_predict_batch / cngp_lml_grad_batch against the CPU oracle (through the C ABI).

Tolerances (BASELINE.json north_star): 1e-9 in FP64, evaluated as |d| <= tol * max(1, |ref|) for mean,
variance and log marginal likelihood.
"""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-9


def close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) if a.size else 0.0


KERNELS = [
    ("rbf", [0.01, 10.0]),
    ("rbf+stdperiodic", [0.01, 10.0, 0.0025, 37.0, 1.0]),
    ("rbf*brownian", [0.01, 10.0, 0.05]),
    ("mat32+bias", [0.02, 7.0, 0.001]),
    ("mat52*linear+white", [0.5, 15.0, 1e-5, 1e-4]),
    ("ratquad+stdperiodic*rbf", [0.01, 8.0, 1.5, 0.5, 37.0, 1.2, 0.01, 50.0]),
    ("(rbf+linear)*brownian+white", [0.02, 12.0, 1e-6, 0.03, 2e-4]),
]


@pytest.mark.parametrize("N,M", [(100, 600), (256, 600), (128, 600), (8, 8), (37, 13), (255, 601), (1, 5)])
@pytest.mark.parametrize("kname,theta", KERNELS[:3])
def test_predict_matches_oracle_shapes(gp_ctx, kname, theta, N, M):
    B = 3
    x, y = syn.slip_windows(11, B, N)
    xs = syn.test_grid(x[0], M)
    th = np.array(theta + [1e-3])
    mean, var, lml, status = gp_ctx.predict(kname, th, x, y, xs)
    assert np.all(status == 0)
    e = go.KernelExpr(kname)
    for b in range(B):
        inf = go.inference(e, th[:-1], th[-1], x[b], y[b])
        mu, v = go.predict(e, th[:-1], th[-1], x[b], y[b], xs, inf)
        assert close(mean[b], mu) < TOL
        assert close(var[b], v) < TOL
        assert close(lml[b], inf.lml) < TOL


@pytest.mark.parametrize("kname,theta", KERNELS)
def test_predict_all_families(gp_ctx, kname, theta):
    B, N, M = 4, 120, 77
    x, y = syn.slip_windows(3, B, N)
    x = x + np.linspace(0, 0.3, N)[None, :] * np.arange(B)[:, None]      # ragged, non-integer time stamps
    xs = np.stack([syn.test_grid(x[b], M) for b in range(B)])             # per-window grids
    rng = np.random.default_rng(5)
    th = np.array(theta + [2e-3])[None, :] * rng.uniform(0.8, 1.25, (B, len(theta) + 1))   # per-window hypers
    mean, var, lml, status = gp_ctx.predict(kname, th, x, y, xs)
    assert np.all(status == 0)
    e = go.KernelExpr(kname)
    for b in range(B):
        inf = go.inference(e, th[b, :-1], th[b, -1], x[b], y[b])
        mu, v = go.predict(e, th[b, :-1], th[b, -1], x[b], y[b], xs[b], inf)
        assert close(mean[b], mu) < TOL
        assert close(var[b], v) < TOL
        assert close(lml[b], inf.lml) < TOL


@pytest.mark.parametrize("kname,theta", KERNELS)
def test_lml_grad_matches_oracle(gp_ctx, kname, theta):
    B, N, C = 3, 96, 4
    x, y = syn.slip_windows(100, B, N)
    rng = np.random.default_rng(9)
    th = np.array(theta + [5e-3])[None, :] * rng.uniform(0.7, 1.4, (C, len(theta) + 1))
    lml, grad, status = gp_ctx.lml_grad(kname, th, x, y)
    assert np.all(status == 0)
    e = go.KernelExpr(kname)
    for c in range(C):
        for b in range(B):
            inf = go.inference(e, th[c, :-1], th[c, -1], x[b], y[b], want_grad=True)
            assert close(lml[c, b], inf.lml) < TOL
            # gradients are sums of O(N^2) terms of size |dL_dK| |dK|: compare relative to the gradient scale
            scale = np.maximum(1.0, np.abs(inf.grad))
            assert np.max(np.abs(grad[c, b] - inf.grad) / scale) < 1e-8, (kname, c, b, grad[c, b], inf.grad)


def test_lml_grad_full_size(gp_ctx):
    B, N, C = 2, 256, 3
    x, y = syn.slip_windows(7, B, N)
    base = syn.theta_for("rbf+stdperiodic")
    th = base[None, :] * np.array([[1.0], [1.3], [0.75]])
    lml, grad, status = gp_ctx.lml_grad("rbf+stdperiodic", th, x, y)
    e = go.KernelExpr("rbf+stdperiodic")
    for c in range(C):
        for b in range(B):
            inf = go.inference(e, th[c, :-1], th[c, -1], x[b], y[b], want_grad=True)
            assert close(lml[c, b], inf.lml) < TOL
            assert np.max(np.abs(grad[c, b] - inf.grad) / np.maximum(1.0, np.abs(inf.grad))) < 1e-8


def test_slipval_fixture(gp_ctx, slipval):
    """The only real slip series of the reference (core_navigation/script/slipVal.csv), deployed kernel, all-ones hypers."""
    t, s = slipval
    xtr, ytr = go.split_train(t, s)
    grid = go.prediction_grid(t)
    th = np.ones(4)
    mean, var, lml, status = gp_ctx.predict("rbf*brownian", th, xtr[None], ytr[None], grid)
    e = go.KernelExpr("rbf*brownian")
    inf = go.inference(e, th[:-1], th[-1], xtr, ytr)
    mu, v = go.predict(e, th[:-1], th[-1], xtr, ytr, grid, inf)
    assert status[0] == 0
    assert close(lml[0], inf.lml) < TOL and close(mean[0], mu) < TOL and close(var[0], v) < TOL


def test_not_positive_definite_reports_status(gp_ctx):
    # duplicate time stamps + zero noise + linear kernel -> rank-deficient Ky (only the 1e-8 jitter on the diagonal)
    N = 16
    x = np.tile(np.arange(N, dtype=float), (2, 1))
    x[1, :] = 5.0
    y = np.zeros((2, N))
    th = np.array([-1.0, 0.0])           # negative variance: not PD
    mean, var, lml, status = gp_ctx.predict("linear", th, x, y, np.arange(4.0))
    assert np.all(status < 0)
    assert np.all(np.isnan(mean)) and np.all(np.isnan(lml))


def test_device_pointer_path_matches_host_path(gp_ctx):
    import torch
    B, N, M = 5, 128, 600
    x, y = syn.slip_windows(40, B, N)
    xs = syn.test_grid(x[0], M)
    th = syn.theta_for("rbf+stdperiodic")
    m0, v0, l0, s0 = gp_ctx.predict("rbf+stdperiodic", th, x, y, xs)
    dx, dy, dxs, dth = (torch.from_numpy(a).cuda() for a in (x, y, xs, th))
    m1, v1, l1, s1 = gp_ctx.predict("rbf+stdperiodic", dth, dx, dy, dxs)
    torch.cuda.synchronize()
    assert np.array_equal(m0, m1.cpu().numpy()) and np.array_equal(v0, v1.cpu().numpy())
    assert np.array_equal(l0, l1.cpu().numpy())


@pytest.mark.parametrize("pinned", [False, True])
def test_pipelined_host_path_matches_device_path(gp_ctx, pinned):
    """Host-memory calls with B >= 1024 run in four slabs with the copies on a second stream (cngp_api.cu): ragged slab
    sizes, per-window theta and xstar, pageable and pinned buffers - results equal the device-pointer path bit for bit."""
    import torch
    B, N, M = 1027, 40, 37
    x, y = syn.slip_windows(7, B, N)
    xs = np.stack([syn.test_grid(x[b], M) + 0.25 * (b % 3) for b in range(B)])
    th = np.tile(syn.theta_for("rbf"), (B, 1)) * (1.0 + 0.1 * (np.arange(B) % 5))[:, None]
    if pinned:
        hx, hy, hxs, hth = (torch.from_numpy(a).pin_memory() for a in (x, y, xs, th))
        out = (torch.empty(B, M, dtype=torch.float64).pin_memory(), torch.empty(B, M, dtype=torch.float64).pin_memory(),
               torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
        m0, v0, l0, s0 = gp_ctx.predict("rbf", hth, hx, hy, hxs, out=out)
        m0, v0, l0 = m0.numpy(), v0.numpy(), l0.numpy()
    else:
        m0, v0, l0, s0 = gp_ctx.predict("rbf", th, x, y, xs)
    dx, dy, dxs, dth = (torch.from_numpy(a).cuda() for a in (x, y, xs, th))
    m1, v1, l1, s1 = gp_ctx.predict("rbf", dth, dx, dy, dxs)
    torch.cuda.synchronize()
    assert np.array_equal(m0, m1.cpu().numpy()) and np.array_equal(v0, v1.cpu().numpy())
    assert np.array_equal(l0, l1.cpu().numpy())
    assert np.isfinite(m0).all() and (v0 > 0).all()
