"""GPU parity of cngp_slip_record_batch (slip extraction + GP_Input recorder, CoreNav.cpp:244-329) against the C oracle.
Recorder decisions (window membership, counts, publication) are bit-identical; slip itself differs only by the last-ulp
differences between CUDA's and glibc's sin/cos (|d| <= 1e-14 on values of O(0.1))."""
import numpy as np
import pytest
import torch

from corenav_gp_b200 import synthetic as syn
from oracle import slip_oracle as so

pytestmark = pytest.mark.gpu


def compare(out, ref, host=True):
    g = (lambda a: a) if host else (lambda a: a.cpu().numpy())
    for k in ("n_windows", "n_samples", "published", "stop_update"):
        assert np.array_equal(g(out[k]), ref[k]), k
    assert np.array_equal(g(out["time_array"]), ref["time_array"])
    np.testing.assert_allclose(g(out["slip_array"]), ref["slip_array"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(g(out["slip"]), ref["slip"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("T", [420, 33, 1])
def test_matches_oracle_host_buffers(gp_ctx, T):
    d = syn.drives(0, 96, T=T)
    out = gp_ctx.slip_record(d["joint"], d["att"], d["vel"], d["cmd"], d["stop_cmd"], max_windows=3, cap=149)
    ref = so.slip_record(d["joint"], d["att"], d["vel"], d["cmd"], d["stop_cmd"], max_windows=3, cap=149)
    compare(out, ref)
    if T == 420:
        assert (ref["n_windows"] >= 2).any() and ref["published"][:, 0].all()


def test_device_buffers_no_stop_commands_and_overflow(gp_ctx):
    d = syn.drives(1000, 257, T=700, stop_events=False)
    dev = {k: torch.from_numpy(v).cuda() for k, v in d.items() if k != "stop_cmd"}
    out = gp_ctx.slip_record(dev["joint"], dev["att"], dev["vel"], dev["cmd"], None, max_windows=1, cap=40)
    ref = so.slip_record(d["joint"], d["att"], d["vel"], d["cmd"], None, max_windows=1, cap=40)
    compare(out, ref, host=False)
    assert (ref["n_windows"] >= 2).all() and (ref["n_samples"][:, 0] > 40).all()


def test_recorded_window_feeds_the_gp(gp_ctx):
    """End to end from raw odometry: recorder -> GP_Input -> cngp_gp_slip_batch (makes the Monte-Carlo
    configuration start at the wheel encoders)."""
    d = syn.drives(7, 4, T=300, stop_events=False)
    out = gp_ctx.slip_record(d["joint"], d["att"], d["vel"], d["cmd"], None, max_windows=1, cap=149)
    n = int(out["n_samples"][:, 0].min())
    assert n >= 100
    t, s = out["time_array"][:, 0, :n], out["slip_array"][:, 0, :n]
    mean, sigma, status = gp_ctx.gp_slip("rbf", t, s, theta=syn.theta_for("rbf"))
    assert (status >= 0).all() and np.isfinite(mean).all() and (sigma > 0).all()
