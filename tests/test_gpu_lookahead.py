"""GPU parity of cngp_zupt_lookahead_batch / cngp_llh_to_enu.

Chain of trust: the reference's own GPCallBack compiled unmodified (oracle/_ref) == oracle/stop_oracle.c with libm
trigonometry to 1e-12 (tests/test_ref_stop.py); stop_oracle.c with trig_mode = 1 differs from that only in sin/cos/tan
(a deterministic implementation, <= 1 ulp from libm) and is reproduced by the CUDA kernels BIT FOR BIT: decisions and
xy_err are compared with array_equal, including thresholds placed exactly on a value of the error trace."""
import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import ref_gp_predictor as rg
from oracle import stop_oracle as so

pytestmark = pytest.mark.gpu


def ocfg(**kw):
    """Oracle configuration with the deterministic trigonometry the CUDA kernels use."""
    return so.default_cfg(trig_mode=1, **kw)


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
    ref = np.stack([so.llh_to_enu(*p, ocfg()) for p in llh])
    assert np.array_equal(enu, ref)                  # same operation sequence on both sides
    libm = np.stack([so.llh_to_enu(*p) for p in llh])
    assert np.max(np.abs(enu - libm)) < 1e-8         # metres; ECEF magnitudes are ~6e6 so this is ~1e-15 relative


@pytest.mark.parametrize("M", [600, 37, 1])
def test_lookahead_matches_oracle_shared_context(gp_ctx, M):
    B = 64
    mean, sigma = gp_outputs(B, M)
    ctx = syn.lookahead_context(0.5)
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"], ocfg())
    for k in ("triggered", "i_stop", "step_stop", "xy_err"):
        assert np.array_equal(out[k], ref[k]), k


def test_lookahead_per_window_context_mixed_triggers(gp_ctx):
    B, M = 256, 150
    mean, sigma = gp_outputs(B, M, seed=3)
    s = syn.window_sigmas(0, B)
    ctx = syn.lookahead_context(s)
    assert ctx["P"].shape == (B, 225)
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"], ocfg())
    assert 0 < ref["triggered"].sum() < B, "fixture must contain both triggering and non-triggering windows"
    for k in ("triggered", "i_stop", "step_stop", "xy_err"):
        assert np.array_equal(out[k], ref[k]), k


def test_lookahead_near_threshold_decisions(gp_ctx):
    """Adversarial: put the threshold EXACTLY on the error reached at a given step (no trigger there: the test is
    `xy_err > thresh`, gp_predictor.cpp:102), one ulp below it (trigger), one ulp above, and a hair away.

    The whole look-ahead is bit-reproducible - same fma order in the matrix algebra, the same deterministic sin/cos on
    both sides, IEEE sqrt and division - so the decision is identical at a 0-ulp margin."""
    M = 200
    mean, sigma = gp_outputs(1, M, seed=5)
    ctx = syn.lookahead_context(0.3)
    tr = so.lookahead(mean[0], sigma[0], ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                      ocfg(thresh=1e9), want_trace=True)["xy_trace"]
    for step in (0, 3, 57, 500, 999):
        for thr in (tr[step], np.nextafter(tr[step], 0.0), np.nextafter(tr[step], 1e9), tr[step] * (1 - 1e-12),
                    tr[step] * (1 + 1e-12), tr[step] * (1 - 1e-8), tr[step] * (1 + 1e-8)):
            ref = so.lookahead(mean[0], sigma[0], ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                               ocfg(thresh=thr))
            out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"],
                                        gp_ctx.stop_config(thresh=thr))
            assert bool(out["triggered"][0]) == ref["triggered"]
            assert int(out["i_stop"][0]) == ref["i_stop"] and int(out["step_stop"][0]) == ref["step_stop"]
            assert out["xy_err"][0] == ref["xy_err"]


def test_lookahead_fix_h_packing_and_ratio(gp_ctx):
    B, M = 8, 50
    mean, sigma = gp_outputs(B, M, seed=8)
    ctx = syn.lookahead_context(0.6)
    hv = np.zeros(60)
    hv[:] = ctx["H"].reshape(60)        # intended row-major packing
    for ratio in (5, 1, 3):
        cfg_o = ocfg(fix_h_packing=1, ratio=ratio, thresh=2.0)
        cfg_g = gp_ctx.stop_config(fix_h_packing=1, ratio=ratio, thresh=2.0)
        out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], hv, ctx["pos"], cfg_g)
        ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], hv, ctx["pos"], cfg_o)
        for k in ("triggered", "i_stop", "step_stop"):
            assert np.array_equal(out[k], ref[k]), (k, ratio)


@pytest.mark.parametrize("kernel", ["tc", "warp", "cta"])
def test_both_kernel_shapes_match_the_oracle(gp_ctx, monkeypatch, kernel):
    """Three implementations of the same bits: the tensor-core kernel (default; one DMMA = an ascending fma chain, see
    tools/dmma_semantics.cu) and the two scalar kernels of round 1 (one warp / one CTA per window).  All keep the
    oracle's fma order, so the decisions - and every kernel's xy_err - are identical."""
    monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", kernel)
    B, M = 96, 260
    mean, sigma = gp_outputs(B, M, seed=5)
    ctx = syn.lookahead_context(syn.window_sigmas(50, B, lo=0.3, hi=0.9))
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"], ocfg())
    assert 0 < ref["triggered"].sum()
    for k in ("triggered", "i_stop", "step_stop", "xy_err"):
        assert np.array_equal(out[k], ref[k]), k
    for k2 in ("tc", "warp", "cta"):
        monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", k2)
        other = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
        assert np.array_equal(out["xy_err"], other["xy_err"]) and np.array_equal(out["step_stop"], other["step_stop"]), k2


def test_dense_stm_takes_the_general_propagation(gp_ctx, monkeypatch):
    """An STM whose bias rows are not unit rows disables the copy shortcut (both kernels)."""
    B, M = 8, 60
    mean, sigma = gp_outputs(B, M, seed=9)
    ctx = syn.lookahead_context(0.6)
    F = ctx["STM"].reshape(15, 15).copy()
    F[10, 2] = 1e-4                                        # couples a bias state: rows 9..14 are no longer unit rows
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], F.reshape(225), ctx["Hvec"], ctx["pos"], ocfg())
    for kernel in ("tc", "warp", "cta"):
        monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", kernel)
        out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], F.reshape(225), ctx["Hvec"], ctx["pos"])
        for k in ("triggered", "i_stop", "step_stop"):
            assert np.array_equal(out[k], ref[k]), (kernel, k)


def test_lookahead_matches_the_reference_binary(gp_ctx):
    """CUDA kernel against the reference's OWN GPCallBack (oracle/_ref: gp_predictor.cpp compiled unmodified; built in
    the authoring container, shipped as a .so) and against the committed vectors it produced (tests/golden/
    stop_ref_golden.npz): decisions equal, xy_err to 1e-9 relative (libm vs deterministic trigonometry on ECEF-sized
    intermediates: 1 ulp of sin at 6.4e6 m is 1e-9 m), final decisions equal for every window of the fixture."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stop_ref_golden.npz"))
    out = gp_ctx.zupt_lookahead(g["mean"], g["sigma"], g["P"], g["Q"], g["STM"], g["Hvec"], g["pos"])
    assert np.array_equal(out["triggered"], g["triggered"])
    assert np.array_equal(out["i_stop"], g["i_stop"])
    assert np.array_equal(out["step_stop"], g["step_stop"])
    assert np.max(np.abs(out["xy_err"] - g["xy_err"]) / np.abs(g["xy_err"])) < 1e-9
    if rg.available():
        for b in range(0, g["mean"].shape[0], 7):
            r = rg.gp_callback(g["mean"][b], g["sigma"][b], g["P"][b], g["Q"][b], g["STM"][b], g["Hvec"][b], g["pos"][b])
            assert r["triggered"] == bool(out["triggered"][b]) and r["i_stop"] == out["i_stop"][b]
            assert abs(r["xy_err"] - out["xy_err"][b]) < 1e-9 * r["xy_err"]


def test_tensor_core_kernel_many_windows_per_cta_and_monte_carlo_contexts(gp_ctx):
    """More windows than 4 x 148 (eight warps per CTA), per-window P and Q of the Monte-Carlo fixture: some windows run
    the whole horizon without a trigger, some trigger at step 0."""
    B, M = 1500, 120
    mean, sigma = gp_outputs(B, M, seed=12)
    ctx = syn.monte_carlo_contexts(1000, B)
    out = gp_ctx.zupt_lookahead(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"])
    ref = so.lookahead_batch(mean, sigma, ctx["P"], ctx["Q"], ctx["STM"], ctx["Hvec"], ctx["pos"], ocfg())
    assert 0 < ref["triggered"].sum() < B
    for k in ("triggered", "i_stop", "step_stop", "xy_err"):
        assert np.array_equal(out[k], ref[k]), k


@pytest.mark.parametrize("kernel", ["tc", "warp", "cta"])
def test_final_state_matches_the_reference_members(gp_ctx, monkeypatch, kernel):
    """cngp_zupt_lookahead_batch_ex: P_pred / K_pred / R_IP as the reference leaves them in its public members after the
    callback (gp_predictor.h:36-43) - against the oracle bit for bit, and against the values the reference binary itself
    produced (golden file, 1e-12: only the rounding inside a matrix product differs)."""
    import os
    monkeypatch.setenv("CNGP_LOOKAHEAD_KERNEL", kernel)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stop_ref_golden.npz"))
    out = gp_ctx.zupt_lookahead(g["mean"], g["sigma"], g["P"], g["Q"], g["STM"], g["Hvec"], g["pos"], want_state=True)
    scale = np.abs(g["P_final"]).max(axis=1, keepdims=True)
    assert np.max(np.abs(out["P_final"] - g["P_final"]) / scale) < 1e-12
    assert np.max(np.abs(out["K_final"] - g["K_final"]) / np.abs(g["K_final"]).max(axis=1, keepdims=True)) < 1e-11
    assert np.max(np.abs(out["R_final"] - g["R_final"]) / np.abs(g["R_final"]).max(axis=1, keepdims=True)) < 1e-14
    for b in (0, 5, 17, 40):
        o = so.lookahead(g["mean"][b], g["sigma"][b], g["P"][b], g["Q"][b], g["STM"][b], g["Hvec"][b], g["pos"][b], ocfg())
        assert np.array_equal(out["P_final"][b].reshape(15, 15), o["P"])
        assert np.array_equal(out["K_final"][b].reshape(15, 4), o["K"])
        assert np.array_equal(out["R_final"][b].reshape(4, 4), o["R"])
