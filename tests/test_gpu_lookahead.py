"""GPU parity of cngp_zupt_lookahead_batch / cngp_llh_to_enu against the C oracle (decisions bit-identical)."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import stop_oracle as so

pytestmark = pytest.mark.gpu


def gp_outputs(B, M, seed=0):
    """Plausible GP_Output arrays: mean decaying to ~0, 2-sigma band widening (Kernel Selection/docs/ours.jpg)."""
    rng = np.random.default_rng(seed)
    k = np.arange(M)
    mean = 0.05 * np.exp(-k / 80.0)[None, :] * rng.uniform(-1, 1, (B, 1)) + 0.01 * rng.standard_normal((B, M))
    sigma = 2.0 * np.sqrt(1e-3 + 0.01 * (1 - np.exp(-k / 150.0)))[None, :] * rng.uniform(0.7, 1.3, (B, 1))
    return mean, sigma


def test_llh_to_enu(gp_ctx):
    rng = np.random.default_rng(1)
    llh = syn.INIT_LLH[None, :] + rng.normal(0, [1e-6, 1e-6, 2.0], (64, 3))
    enu = gp_ctx.llh_to_enu(llh)
    ref = np.stack([so.llh_to_enu(*p) for p in llh])
    assert np.max(np.abs(enu - ref)) < 1e-8          # metres; ECEF magnitudes are ~6e6 so this is ~1e-15 relative
    # (CUDA and glibc sin/cos/tan differ in the last ulp, which is all that is left here)


@pytest.mark.parametrize("M", [600, 37, 1])
def test_lookahead_matches_oracle_shared_context(gp_ctx, M):
    B = 64
    mean, sigma = gp_outputs(B, M)
    ctx = syn.lookahead_context(0.5)
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    for k in ("triggered", "i_stop", "step_stop"):
        assert np.array_equal(out[k], ref[k]), k
    assert np.max(np.abs(out["xy_err"] - ref["xy_err"]) / np.maximum(1, np.abs(ref["xy_err"]))) < 1e-9


def test_lookahead_per_window_context_mixed_triggers(gp_ctx):
    B, M = 256, 150
    mean, sigma = gp_outputs(B, M, seed=3)
    s = syn.window_sigmas(0, B)
    ctx = syn.lookahead_context(s)
    assert ctx["P"].shape == (B, 225)
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    assert 0 < ref["triggered"].sum() < B, "fixture must contain both triggering and non-triggering windows"
    for k in ("triggered", "i_stop", "step_stop"):
        assert np.array_equal(out[k], ref[k]), k


def test_lookahead_near_threshold_decisions(gp_ctx):
    """Adversarial: put the threshold a hair above / below the error reached at a given step.

    The matrix part of the look-ahead is bit-reproducible (same fma order on both sides), but xy_err is a difference of
    ENU coordinates formed from ECEF values of magnitude 5e6 m (gp_predictor.cpp:161-167), so last-ulp differences
    between CUDA's and glibc's sin/cos/tan move it by ~1e-9 m.  Decisions are therefore identical whenever the error
    is not within ~1e-8 m of the threshold; margins of 1e-8 and 1e-7 relative (3e-8 m, 3e-7 m) are tested."""
    M = 200
    mean, sigma = gp_outputs(1, M, seed=5)
    ctx = syn.lookahead_context(0.3)
    tr = so.lookahead(mean[0], sigma[0], ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                      so.default_cfg(thresh=1e9), want_trace=True)["xy_trace"]
    for step in (3, 57, 500, 999):
        for rel in (-1e-7, 1e-7, -1e-8, 1e-8):
            thr = tr[step] * (1 + rel)
            ref = so.lookahead(mean[0], sigma[0], ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                               so.default_cfg(thresh=thr))
            out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                                        gp_ctx.stop_config(thresh=thr))
            assert bool(out["triggered"][0]) == ref["triggered"]
            assert int(out["i_stop"][0]) == ref["i_stop"] and int(out["step_stop"][0]) == ref["step_stop"]


def test_lookahead_fix_h_packing_and_ratio(gp_ctx):
    B, M = 8, 50
    mean, sigma = gp_outputs(B, M, seed=8)
    ctx = syn.lookahead_context(0.6)
    hv = np.zeros(60)
    hv[:] = ctx["H"].reshape(60)        # intended row-major packing
    for ratio in (5, 1, 3):
        cfg_o = so.default_cfg(fix_h_packing=1, ratio=ratio, thresh=2.0)
        cfg_g = gp_ctx.stop_config(fix_h_packing=1, ratio=ratio, thresh=2.0)
        out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], hv, ctx["pos"], cfg_g)
        ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], hv, ctx["pos"], cfg_o)
        for k in ("triggered", "i_stop", "step_stop"):
            assert np.array_equal(out[k], ref[k]), (k, ratio)


@pytest.mark.parametrize("kernel", ["warp", "cta"])
def test_both_kernel_shapes_match_the_oracle(gp_ctx, monkeypatch, kernel):
    """Large batches run one warp per window, small ones (<= 592 windows, the reference's single callback) one CTA per
    window; both keep the oracle's fma order, so the decisions - and the two kernels' xy_err - are identical."""
    monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", kernel)
    B, M = 96, 260
    mean, sigma = gp_outputs(B, M, seed=5)
    ctx = syn.lookahead_context(syn.window_sigmas(50, B, lo=0.3, hi=0.9))
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    assert 0 < ref["triggered"].sum()
    for k in ("triggered", "i_stop", "step_stop"):
        assert np.array_equal(out[k], ref[k]), k
    assert np.max(np.abs(out["xy_err"] - ref["xy_err"]) / np.maximum(1, np.abs(ref["xy_err"]))) < 1e-9
    monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", "cta" if kernel == "warp" else "warp")
    other = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    assert np.array_equal(out["xy_err"], other["xy_err"]) and np.array_equal(out["step_stop"], other["step_stop"])


def test_dense_stm_takes_the_general_propagation(gp_ctx, monkeypatch):
    """An STM whose bias rows are not unit rows disables the copy shortcut (both kernels)."""
    B, M = 8, 60
    mean, sigma = gp_outputs(B, M, seed=9)
    ctx = syn.lookahead_context(0.6)
    F = ctx["STM"].reshape(15, 15).copy()
    F[10, 2] = 1e-4                                        # couples a bias state: rows 9..14 are no longer unit rows
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], F.reshape(225), ctx["Hvec"], ctx["pos"])
    for kernel in ("warp", "cta"):
        monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", kernel)
        out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], F.reshape(225), ctx["Hvec"], ctx["pos"])
        for k in ("triggered", "i_stop", "step_stop"):
            assert np.array_equal(out[k], ref[k]), (kernel, k)
