"""True relative errors of the FP64 path on the BASELINE.json shapes, the GPy jitter ladder, and the r^2 clip.

BASELINE.json north_star: "1e-9 relative in FP64 ... for predictive mean, variance and log-likelihood".  The error
measure here is the TRUE relative error |d| / |ref| of every entry whose magnitude is at least 1e-3 of the largest entry
of its output vector (the predictive mean crosses zero: entries below that floor are measured against the floor, i.e.
to 1e-12 of the scale); variance and LML never come near zero, for them it is the plain relative error.  The numbers
are printed and written to gpurun_out/relerr_fp64.json (copied to profiles/ by the round script)."""
import json
import os

import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from corenav_gp_b200.api import GpContext
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, ref, floor=1e-3):
    a, ref = np.asarray(a, dtype=float), np.asarray(ref, dtype=float)
    scale = max(float(np.max(np.abs(ref))), 1e-300)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor * scale)))


SHAPES = [
    ("configs[0] N=100 SE", "rbf", 100, 600),
    ("configs[1] N=256 SE+periodic", "rbf+stdperiodic", 256, 600),
    ("configs[3] N=128 SE+periodic", "rbf+stdperiodic", 128, 600),
    ("deployed kernel rbf*brownian N=134", "rbf*brownian", 134, 600),
]


def test_true_relative_error_on_baseline_shapes(gp_ctx):
    report = {}
    for name, kname, N, M in SHAPES:
        B = 6
        x, y = syn.slip_windows(77, B, N)
        xs = syn.test_grid(x[0], M)
        th = syn.theta_for(kname)
        mean, var, lml, status = gp_ctx.predict(kname, th, x, y, xs)
        assert np.all(status == 0)
        e = go.KernelExpr(kname)
        em = ev = el = 0.0
        for b in range(B):
            inf = go.inference(e, th[:-1], th[-1], x[b], y[b])
            mu, v = go.predict(e, th[:-1], th[-1], x[b], y[b], xs, inf)
            em = max(em, relerr(mean[b], mu))
            ev = max(ev, float(np.max(np.abs(var[b] - v) / np.abs(v))))
            el = max(el, abs(lml[b] - inf.lml) / abs(inf.lml))
        report[name] = {"mean": em, "var": ev, "lml": el}
        print(f"{name}: true relative error mean {em:.2e} var {ev:.2e} lml {el:.2e}")
        assert em < 1e-9 and ev < 1e-9 and el < 1e-9, (name, em, ev, el)
    # configs[2]: LML and gradient of the Kernel Selection families at N = 256
    fams = {"rbf": [0.01, 10.0], "mat32": [0.02, 7.0], "mat52": [0.02, 9.0], "ratquad": [0.01, 8.0, 1.5],
            "rbf+stdperiodic": [0.01, 10.0, 0.0025, 37.0, 1.0], "rbf*brownian": [0.01, 10.0, 0.05],
            "mat32+linear": [0.02, 7.0, 1e-6], "rbf+bias": [0.01, 10.0, 0.001]}
    x, y = syn.slip_windows(5, 2, 256)
    for kname, theta in fams.items():
        th = np.array(theta + [2e-3])[None, :]
        lml, grad, status = gp_ctx.lml_grad(kname, th, x, y)
        e = go.KernelExpr(kname)
        el = eg = 0.0
        for b in range(2):
            inf = go.inference(e, th[0, :-1], th[0, -1], x[b], y[b], want_grad=True)
            el = max(el, abs(lml[0, b] - inf.lml) / abs(inf.lml))
            eg = max(eg, relerr(grad[0, b], inf.grad))
        report[f"configs[2] {kname} N=256"] = {"lml": el, "grad": eg}
        print(f"configs[2] {kname}: true relative error lml {el:.2e} grad (floor 1e-3 of the largest component) {eg:.2e}")
        assert el < 1e-9 and eg < 1e-8, (kname, el, eg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "relerr_fp64.json"), "w") as f:
        json.dump(report, f, indent=1)


@pytest.mark.parametrize("deficit,tries", [(0.0, 1), (5e-5, 3)])
def test_jitter_ladder_matches_jitchol(deficit, tries):
    """GPy's jitchol (util/linalg.py; oracle go.jitchol): when dpotrf fails, retry with jitter = mean(diag) 1e-6, x10 per
    try.  Ky = 1 1' + (noise + 1e-8) I with noise = -1e-8 - deficit is exactly singular (deficit 0: the second pivot is an
    exact 0) or indefinite (deficit 5e-5: needs jitter 1e-4, the third rung).  status = number of tries used."""
    N, M = 64, 24
    rng = np.random.default_rng(2)
    x = np.tile(np.arange(N, dtype=float), (2, 1))
    y = 0.1 * rng.standard_normal((2, N))
    xs = N + np.arange(M, dtype=float)
    th = np.array([1.0, -1e-8 - deficit])
    e = go.KernelExpr("bias")
    ctx = GpContext(device=0, jitter_retry=True)
    mean, var, lml, status = ctx.predict("bias", th, x, y, xs)
    ctx.close()
    for b in range(2):
        inf = go.inference(e, th[:-1], th[-1], x[b], y[b])
        assert inf.jitter > 0 and abs(inf.jitter / ((1.0 - deficit) * 1e-6 * 10 ** (tries - 1)) - 1) < 1e-9
        assert status[b] == tries, (status, tries)
        mu, v = go.predict(e, th[:-1], th[-1], x[b], y[b], xs, inf)
        # cond(Ky + jitter I) ~ N / (jitter - deficit) ~ 1e6 .. 6e7: the two factorizations agree to cond x eps
        assert relerr(mean[b], mu, floor=1.0) < 1e-6 and relerr(var[b], v, floor=1.0) < 1e-6
        assert abs(lml[b] - inf.lml) < 1e-6 * abs(inf.lml)
    # without the ladder the same windows are reported as not positive definite
    ctx0 = GpContext(device=0, jitter_retry=False)
    m0, v0, l0, s0 = ctx0.predict("bias", th, x, y, xs)
    ctx0.close()
    assert np.all(s0 < 0) and np.all(np.isnan(m0))


@pytest.mark.parametrize("kname,theta", [("rbf", [0.5, 0.7]), ("rbf+stdperiodic", [0.5, 0.7, 0.1, 37.0, 1.0]),
                                         ("rbf*brownian", [0.5, 0.7, 1e-6]), ("mat32", [0.5, 0.7])])
def test_r2_clip_with_large_stamps(gp_ctx, kname, theta):
    """GPy forms r^2 = -2 x x' + (x^2 + x'^2) and clips it at 0.  At |x| ~ 1e6 the expanded form is off by ~1e-4 either
    way; with a length scale below 1 an unclipped negative r^2 would change K by ~1e-4.  Near-duplicate, non-integer
    stamps (so the lazy evaluators run, not the lag table) around 1e6: parity must hold at 1e-9."""
    rng = np.random.default_rng(6)
    N, M = 48, 16
    x = 1.0e6 + np.sort(rng.uniform(0.0, 12.0, (2, N)), axis=1) + 0.123
    k = x[:, 1::5].shape[1]
    x[:, 1::5] = x[:, 0::5][:, :k] + rng.uniform(0.003, 0.012, (2, k))     # close pairs: r^2 ~ 5e-5 under a noise of 1.2e-4
    x = np.sort(x, axis=1)
    y = 0.05 * np.sin(x - 1.0e6) + 0.01 * rng.standard_normal((2, N))
    xs = np.stack([x[b].max() + 0.37 + np.arange(M) for b in range(2)])
    th = np.array(theta + [1e-2])
    mean, var, lml, status = gp_ctx.predict(kname, th, x, y, xs)
    assert np.all(status == 0)
    e = go.KernelExpr(kname)
    neg = 0
    for b in range(2):
        r2 = -2.0 * np.multiply.outer(x[b], x[b]) + (np.square(x[b])[:, None] + np.square(x[b])[None, :])
        neg += int((r2 < 0).sum())
        inf = go.inference(e, th[:-1], th[-1], x[b], y[b])
        mu, v = go.predict(e, th[:-1], th[-1], x[b], y[b], xs[b], inf)
        assert relerr(mean[b], mu) < 1e-9 and relerr(var[b], v) < 1e-9 and abs(lml[b] - inf.lml) < 1e-9 * abs(inf.lml)
    assert neg > 0, "fixture must produce negative unclipped r^2 values"
