"""FP32 mode of the batched predictor (cngp_config.precision = CNGP_PRECISION_F32; BASELINE.json north_star: outputs
within 1e-4 of the reference in FP32 mode for predictive mean, variance and log-likelihood).

The comparison is against the FP64 oracle, on the BASELINE shapes (configs[0]: N = 100 SE; configs[1]: N = 256
SE + periodic; configs[3]: N = 128) plus the fallback path (non-integer stamps, non-stationary kernels -> K* through
the interpreter) and padded windows.

Error measure of this mode: NORMWISE relative error per window, max|d| / max|ref| < 1e-4, for mean, variance and LML.
An exact GP solve in FP32 carries cond(Ky) eps32 ~ 3e3 x 6e-8 ~ 2e-4 of the output's scale in the worst case (observed:
<= 1e-5 for the mean, <= 1e-5 for the variance), and the predictive mean crosses zero, so no FP32 path - this one or a
CPU one - can hold 1e-4 relative on entries far below the scale.  The entrywise relative error over the entries within
a decade of the scale is printed beside it (a few 1e-5)."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
TOL32 = 1e-4


def relerr(a, ref, floor=1.0):
    a, ref = np.asarray(a, dtype=float), np.asarray(ref, dtype=float)
    scale = max(float(np.max(np.abs(ref))), 1e-300)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor * scale)))


def run(ctx, kname, th, x, y, xs):
    mean, var, lml, status = ctx.predict(kname, th, x, y, xs)
    assert np.all(status == 0)
    e = go.KernelExpr(kname)
    worst = [0.0, 0.0, 0.0, 0.0]
    for b in range(x.shape[0]):
        thb = th if th.ndim == 1 else th[b]
        xsb = xs if xs.ndim == 1 else xs[b]
        inf = go.inference(e, thb[:-1], thb[-1], x[b], y[b])
        mu, v = go.predict(e, thb[:-1], thb[-1], x[b], y[b], xsb, inf)
        worst = [max(worst[0], relerr(mean[b], mu)), max(worst[1], relerr(var[b], v)), max(worst[2], relerr(lml[b], inf.lml)),
                 max(worst[3], relerr(mean[b], mu, floor=0.1))]
    return worst, (mean, var, lml)


@pytest.mark.parametrize("kname,N,M", [("rbf", 100, 600), ("rbf+stdperiodic", 256, 600), ("rbf+stdperiodic", 128, 600),
                                       ("rbf", 37, 13), ("rbf+stdperiodic", 255, 601), ("rbf", 8, 40), ("rbf", 1, 5)])
def test_fp32_mode_within_1e4_on_baseline_shapes(gp_ctx32, kname, N, M):
    x, y = syn.slip_windows(31, 4, N)
    xs = syn.test_grid(x[0], M)
    worst, _ = run(gp_ctx32, kname, syn.theta_for(kname), x, y, xs)
    print(f"fp32 mode {kname} N={N} M={M}: normwise rel err mean {worst[0]:.2e} var {worst[1]:.2e} lml {worst[2]:.2e}; "
          f"entrywise (|ref| >= 0.1 scale) mean {worst[3]:.2e}")
    assert max(worst[:3]) < TOL32, worst


@pytest.mark.parametrize("kname,theta", [("rbf*brownian", [0.01, 10.0, 0.05]), ("mat32+bias", [0.02, 7.0, 0.001]),
                                         ("ratquad+stdperiodic*rbf", [0.01, 8.0, 1.5, 0.5, 37.0, 1.2, 0.01, 50.0])])
def test_fp32_mode_interpreter_path(gp_ctx32, kname, theta):
    """Non-integer stamps / non-stationary kernels: K* is staged through the interpreter instead of the lag table."""
    B, N, M = 3, 120, 77
    x, y = syn.slip_windows(3, B, N)
    x = x + np.linspace(0, 0.3, N)[None, :] * (1 + np.arange(B))[:, None]
    xs = np.stack([syn.test_grid(x[b], M) for b in range(B)])
    th = np.array(theta + [2e-3])[None, :] * np.linspace(0.9, 1.1, B)[:, None]
    worst, _ = run(gp_ctx32, kname, th, x, y, xs)
    assert max(worst[:3]) < TOL32, worst


def test_fp32_mode_differs_from_fp64_and_leaves_fp64_context_alone(gp_ctx, gp_ctx32):
    B, N, M = 5, 256, 600
    x, y = syn.slip_windows(2, B, N)
    xs = syn.test_grid(x[0], M)
    th = syn.theta_for("rbf+stdperiodic")
    m64, v64, l64, _ = gp_ctx.predict("rbf+stdperiodic", th, x, y, xs)
    m32, v32, l32, _ = gp_ctx32.predict("rbf+stdperiodic", th, x, y, xs)
    assert relerr(m32, m64) < TOL32 and relerr(v32, v64) < TOL32
    assert relerr(m32, m64) > 1e-9, "FP32 mode returned FP64-exact results: the mode switch is not taking effect"
    assert np.array_equal(l32, l64)                          # the factorisation (and the LML) stays FP64
    m64b, v64b, _, _ = gp_ctx.predict("rbf+stdperiodic", th, x, y, xs)
    assert np.array_equal(m64, m64b) and np.array_equal(v64, v64b)


def test_fp32_mode_node_callback(gp_ctx32):
    """cngp_gp_slip_batch (rows a1-a7) in FP32 mode: sigma = 2 sqrt(var) within 1e-4."""
    B, n = 3, 149
    t, s = syn.slip_windows(8, B, n)
    th = np.array([0.01, 10.0, 1e-3])
    mean, sigma, status = gp_ctx32.gp_slip("rbf", t, s, theta=th)
    e = go.KernelExpr("rbf")
    for b in range(B):
        ref = go.gp_slip_callback(t[b], s[b], e, theta=th[:-1], noise=th[-1])
        assert relerr(mean[b], ref[0]) < TOL32 and relerr(sigma[b], ref[1]) < TOL32


def test_set_precision_switches_an_existing_context():
    from corenav_gp_b200.api import GpContext
    ctx = GpContext(device=0)
    x, y = syn.slip_windows(1, 2, 64)
    xs = syn.test_grid(x[0], 32)
    th = syn.theta_for("rbf")
    a = ctx.predict("rbf", th, x, y, xs)
    ctx.set_precision("f32")
    b = ctx.predict("rbf", th, x, y, xs)
    ctx.set_precision("f64")
    c = ctx.predict("rbf", th, x, y, xs)
    ctx.close()
    assert np.array_equal(a[0], c[0]) and not np.array_equal(a[0], b[0]) and relerr(b[0], a[0]) < TOL32
