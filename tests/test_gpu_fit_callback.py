"""GPU tests of cngp_lml_grad_windows, cngp_optimize_batch and cngp_gp_slip_batch (rows a1, a4, a5, a7) against the
CPU oracle: the batched L-BFGS-B fit must reach the optimum scipy's fmin_l_bfgs_b reaches on the oracle objective
(fitted hyper-parameters are compared through the LML they reach), and the callback at fixed
hyper-parameters must match the oracle's callback to 1e-9."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def test_lml_grad_windows_matches_oracle(gp_ctx):
    import ctypes as C
    from corenav_gp_b200 import _lib as L
    from corenav_gp_b200.api import parse_kernel
    B, N = 5, 90
    x, y = syn.slip_windows(21, B, N)
    k = parse_kernel("rbf*brownian")
    P = k.n_params + 1
    wmap = np.array([4, 0, 2, 2, 1, 3, 0], dtype=np.int32)           # windows repeat and are out of order
    rng = np.random.default_rng(0)
    th = np.array([0.02, 9.0, 0.05, 2e-3])[None, :] * rng.uniform(0.7, 1.4, (wmap.size, P))
    lml = np.empty(wmap.size)
    grad = np.empty((wmap.size, P))
    st = np.empty(wmap.size, dtype=np.int32)
    gp_ctx._bind_stream(False)
    rc = gp_ctx.lib.cngp_lml_grad_windows(gp_ctx.h, C.byref(k), th.ctypes.data, wmap.size, wmap.ctypes.data,
                                          x.ctypes.data, y.ctypes.data, B, N, lml.ctypes.data, grad.ctypes.data,
                                          st.ctypes.data, L.MEM_HOST)
    assert rc == 0 and np.all(st == 0)
    e = go.KernelExpr("rbf*brownian")
    for p, w in enumerate(wmap):
        inf = go.inference(e, th[p, :-1], th[p, -1], x[w], y[w], want_grad=True)
        assert rel(lml[p], inf.lml) < TOL
        assert rel(grad[p], inf.grad) < 1e-8


@pytest.mark.parametrize("kname", ["rbf*brownian", "rbf", "rbf+stdperiodic"])
def test_optimize_reaches_scipy_optimum_on_slipval(gp_ctx, slipval, kname):
    t, s = slipval
    xtr, ytr = go.split_train(t, s)
    theta, lml, nfev = gp_ctx.optimize(kname, xtr[None], ytr[None])
    e = go.KernelExpr(kname)
    th_ref, noise_ref, lml_ref, nev_ref = go.optimize(e, xtr, ytr)
    assert np.all(theta > 0) and nfev[0] <= 1000
    # Optimiser trajectories are not reproducible to 1e-9 across implementations: the fitted
    # optimum is compared through the LML it reaches - at least the oracle's; on a multimodal surface (periodic
    # leaf) a 1e-16 difference in the first evaluations may legitimately end in a better local optimum.
    assert lml[0] >= lml_ref - 1e-6 * abs(lml_ref), (lml[0], lml_ref)
    if "periodic" not in kname:
        assert abs(lml[0] - lml_ref) < 1e-4 * abs(lml_ref), (lml[0], lml_ref)
    # the reported LML is the oracle's LML at the returned hyper-parameters
    assert rel(lml[0], go.inference(e, theta[0, :-1], theta[0, -1], xtr, ytr).lml) < TOL


def test_optimize_batch_of_windows(gp_ctx):
    """Windows converge after different numbers of evaluations; the active list shrinks round by round."""
    B, N = 12, 64
    x, y = syn.slip_windows(300, B, N)
    theta, lml, nfev = gp_ctx.optimize("rbf", x, y)
    e = go.KernelExpr("rbf")
    assert len(set(nfev.tolist())) > 1
    for b in range(B):
        _, _, lml_ref, _ = go.optimize(e, x[b], y[b])
        assert lml[b] >= lml_ref - 1e-6 * abs(lml_ref), (b, lml[b], lml_ref)
        assert rel(lml[b], go.inference(e, theta[b, :-1], theta[b, -1], x[b], y[b]).lml) < TOL


def test_gp_slip_callback_fixed_theta_matches_oracle(gp_ctx, slipval):
    t, s = slipval
    th = np.array([0.5, 6.0, 0.01, 2e-3])
    mean, sigma, status = gp_ctx.gp_slip("rbf*brownian", t[None], s[None], theta=th)
    mu, sg = go.gp_slip_callback(t, s, go.KernelExpr("rbf*brownian"), theta=th[:-1], noise=th[-1])
    assert status[0] == 0 and mean.shape == (1, mu.size)
    assert mu.size == 421                     # grid = ceil(46.0 + 600 - 26.2) = 620 points, entries [199:] are published
    assert rel(mean[0], mu) < TOL and rel(sigma[0], sg) < TOL


def test_gp_slip_callback_batch_integer_counts(gp_ctx):
    """Time stamps as the EKF records them (odomUptCount values, CoreNav.cpp:286): n = 149 -> N = 134, 599 kept points."""
    B, n = 6, 149
    x, y = syn.slip_windows(50, B, n)
    x = x + 10.0 * np.arange(B)[:, None]                      # each window starts at its own count
    th = syn.theta_for("rbf*brownian")
    mean, sigma, status = gp_ctx.gp_slip("rbf*brownian", x, y, theta=th)
    assert mean.shape == (B, 599) and np.all(status == 0)
    e = go.KernelExpr("rbf*brownian")
    for b in range(B):
        mu, sg = go.gp_slip_callback(x[b], y[b], e, theta=th[:-1], noise=th[-1])
        assert rel(mean[b], mu) < TOL and rel(sigma[b], sg) < TOL


def test_gp_slip_callback_with_fit(gp_ctx, slipval):
    """theta = None: all-ones start + m.optimize() (gp_slip_node.py:31-36), then predict.  The fitted optimum is compared
    through the outputs it produces: the published mean/sigma agree with the oracle's fit-then-predict to 1e-5."""
    t, s = slipval
    mean, sigma, status = gp_ctx.gp_slip("rbf*brownian", t[None], s[None])
    mu, sg = go.gp_slip_callback(t, s)
    assert status[0] == 0
    assert np.max(np.abs(mean[0] - mu)) < 1e-5 and np.max(np.abs(sigma[0] - sg) / sg) < 1e-4


def test_gp_slip_rejects_ragged_spans(gp_ctx):
    from corenav_gp_b200.api import CngpError
    x, y = syn.slip_windows(0, 2, 40)
    x[1] *= 1.5
    with pytest.raises(CngpError):
        gp_ctx.gp_slip("rbf", x, y, theta=syn.theta_for("rbf"))


def test_python_node_mirror(slipval):
    """corenav_gp_b200.gp_slip_node keeps the reference script's surface: callback(data) publishes a GP_Output."""
    from corenav_gp_b200 import gp_slip_node as node
    t, s = slipval
    got = []
    node.pub.sink = got.append
    th = np.array([0.5, 6.0, 0.01, 2e-3])
    msg = node.callback(node.GP_Input(time_array=t.tolist(), slip_array=s.tolist()), theta=th)
    node.pub.sink = None
    mu, sg = go.gp_slip_callback(t, s, go.KernelExpr("rbf*brownian"), theta=th[:-1], noise=th[-1])
    assert got and got[0] is msg and node.pub.last is msg
    assert rel(msg.mean, mu) < TOL and rel(msg.sigma, sg) < TOL


def test_optimize_on_device_buffers_matches_host_call(gp_ctx):
    """cngp_optimize_batch_mem with CUDA tensors: the whole fit (objective, gradient, L-BFGS-B state machines) stays on
    the device; same optima as the host-buffer call, finished windows are skipped while the others still iterate."""
    import torch
    B, N = 37, 64
    x, y = syn.slip_windows(300, B, N)
    th_h, lml_h, it_h = gp_ctx.optimize("rbf", x, y)
    th_d, lml_d, it_d = gp_ctx.optimize("rbf", torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(th_h, th_d.cpu().numpy()) and np.array_equal(lml_h, lml_d.cpu().numpy())
    assert np.array_equal(it_h, it_d.cpu().numpy())
    assert it_h.min() < it_h.max(), "windows should need different numbers of evaluations"
    e = go.KernelExpr("rbf")
    for b in (0, 11, 36):
        inf = go.inference(e, th_h[b, :-1], th_h[b, -1], x[b], y[b], want_grad=True)
        assert rel(lml_h[b], inf.lml) < TOL                       # the reported LML is the LML at the returned theta
        _, _, lml_ref, _ = go.optimize(e, x[b], y[b])
        assert lml_h[b] >= lml_ref - 1e-6 * abs(lml_ref)
