"""The ROS-free C++ GpPredictor (include/gp_predictor_b200.hpp, corenav_gp_b200/host/gp_predictor.cpp; SURVEY.md row
a13) played through a small CLI driver.  CPU test: host logic against a test double of the C ABI backed by the C
oracle.  GPU test: the same driver linked against the real libcngp.so, compared with the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from corenav_gp_b200 import synthetic as syn
from oracle import stop_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host", "_build")
SRCS = [os.path.join(ROOT, "tests", "host", "gp_predictor_cli.cpp"),
        os.path.join(ROOT, "corenav_gp_b200", "host", "gp_predictor.cpp")]


def build_cli(fake: bool) -> str:
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "gp_predictor_cli_fake" if fake else "gp_predictor_cli")
    if fake:
        so.build()
        obj = os.path.join(BUILD, "fake_cngp.o")
        subprocess.check_call(["gcc", "-O2", "-c", os.path.join(ROOT, "tests", "host", "fake_cngp.c"), "-o", obj])
        link = [obj, os.path.join(ROOT, "oracle", "_build", "libstop_oracle.so"),
                "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")]
    else:
        link = [os.path.join(ROOT, "corenav_gp_b200", "libcngp.so"), "-Wl,-rpath," + os.path.join(ROOT, "corenav_gp_b200")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, *SRCS, *link, "-lm"])
    return exe


def write_case(path, M=120, s0=0.5, seed=0):
    rng = np.random.default_rng(seed)
    k = np.arange(M)
    mean = 0.05 * np.exp(-k / 80.0) * rng.uniform(-1, 1) + 0.01 * rng.standard_normal(M)
    sigma = 2.0 * np.sqrt(1e-3 + 0.01 * (1 - np.exp(-k / 150.0)))
    c = syn.lookahead_context(s0)
    blob = np.concatenate([c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], [float(M)], mean, sigma])
    blob.astype(np.float64).tofile(path)
    return mean, sigma, c


def run(exe, path, advance, service_ok=1):
    r = subprocess.run([exe, path, repr(float(advance)), str(service_ok)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def check_against_oracle(exe, tmp_path):
    path = str(tmp_path / "case.bin")
    mean, sigma, c = write_case(path, M=120, s0=0.7)
    ref = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"], so.default_cfg(trig_mode=1))
    assert ref["triggered"]
    out = run(exe, path, advance=0.25)
    assert out["returned"] == 1 and out["published"] == 1 and out["flag"] == 0
    assert out["i"] == ref["i_stop"] and out["slip_i"] == ref["step_stop"]
    assert out["xy_errSlip"] == ref["xy_err"]
    # public matrix members after the callback (gp_predictor.h:36-46): P_pred is left at the final look-ahead covariance,
    # K_pred / R_IP / R_IP_2 at the last update, ins_enu_slip* at the last error-observer evaluation
    assert np.array_equal(np.array(out["P_pred"]).reshape(15, 15), ref["P"])
    assert np.array_equal(np.array(out["K_pred"]).reshape(15, 4), ref["K"])
    assert np.array_equal(np.array(out["R_IP"]).reshape(4, 4), ref["R"])
    R1 = np.array(out["R_IP_1"]).reshape(4, 4)
    R2 = np.array(out["R_IP_2"]).reshape(4, 4)
    assert np.allclose(25.0 * R1 @ R2 @ R1.T, ref["R"], rtol=1e-14, atol=0)          # gp_predictor.cpp:88
    e0, e3 = np.array(out["ins_enu_slip"]), np.array(out["ins_enu_slip3p"])
    assert abs(np.hypot(*(e3 - e0)[:2]) - ref["xy_err"]) < 1e-9                      # gp_predictor.cpp:99
    assert np.all(np.array(out["ins_enu_slip_3p"]) != e3)
    # gp_predictor.cpp:107-116: stop in (t_gp + i/10 - now) seconds ...
    assert out["stop_cmd"] == pytest.approx(ref["i_stop"] / 10.0 - 0.25, abs=1e-12)
    # ... or 0.5 s when that moment has already passed
    late = run(exe, path, advance=ref["i_stop"] / 10.0 + 1.0)
    assert late["published"] == 1 and late["stop_cmd"] == 0.5
    # a failed service call: nothing runs, nothing is published (gp_predictor.cpp:53-58)
    failed = run(exe, path, advance=0.0, service_ok=0)
    assert failed["published"] == 0 and failed["returned"] == 0 and failed["i"] == 0
    # no trigger inside the horizon: nothing is published
    path2 = str(tmp_path / "case2.bin")
    mean, sigma, c = write_case(path2, M=20, s0=0.2, seed=1)
    ref2 = so.lookahead(mean, sigma, c["P"], c["Q"], c["STM"], c["Hvec"], c["pos"])
    assert not ref2["triggered"]
    quiet = run(exe, path2, advance=0.0)
    assert quiet["published"] == 0 and quiet["i"] == 20 and quiet["slip_i"] == 100
    # llh_to_enu through the class: one metre up from the saved position
    up = np.array(quiet["enu_up"]) - so.llh_to_enu(*syn.INIT_LLH)
    assert abs(up[2] - 1.0) < 1e-6


def test_gp_predictor_host_logic_cpu(tmp_path):
    check_against_oracle(build_cli(fake=True), tmp_path)


@pytest.mark.gpu
def test_gp_predictor_on_gpu(tmp_path):
    check_against_oracle(build_cli(fake=False), tmp_path)


def build_slip_cli() -> str:
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "gp_slip_cli")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "gp_slip_cli.cpp"),
                           os.path.join(ROOT, "corenav_gp_b200", "host", "gp_slip_predict.cpp"),
                           os.path.join(ROOT, "corenav_gp_b200", "libcngp.so"),
                           "-Wl,-rpath," + os.path.join(ROOT, "corenav_gp_b200")])
    return exe


def test_gp_slip_predict_is_exported_and_links():
    """CPU: the C++ mirror of the node callback (gp_slip_node.py:16-63) is part of libgp_predictor_b200.so and the driver
    links against the C ABI (no compute: there is no GPU here)."""
    from corenav_gp_b200 import build
    lib = build.build_host()
    syms = subprocess.run(["nm", "-DC", lib], capture_output=True, text=True).stdout
    assert "gp_slip_predict(cngp_ctx*, core_nav::GP_Input const&, char const*, double const*, int)" in syms
    assert os.path.exists(build_slip_cli())


def slip_window(n=149, seed=0):
    rng = np.random.default_rng(seed)
    t = 21.0 + np.arange(n)
    return t, 0.02 + 0.05 * np.sin(2 * np.pi * np.arange(n) / 37.0) + 0.03 * rng.standard_normal(n)


@pytest.mark.gpu
@pytest.mark.parametrize("fit", [False, True])
def test_gp_slip_predict_matches_the_python_node(gp_ctx, tmp_path, fit):
    """GPU: gp_slip_predict (C++) and corenav_gp_b200.gp_slip_node / GpContext.gp_slip (Python) are the same ABI call -
    identical bits - and the fixed-hyper-parameter result is the oracle's callback to 1e-9."""
    from oracle import gp_oracle as go
    exe = build_slip_cli()
    t, s = slip_window()
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([[float(t.size)], t, s]).tofile(src)
    r = subprocess.run([exe, str(src), str(dst)] + (["fit"] if fit else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout.strip().splitlines()[-1])["seq"] == 7
    out = np.fromfile(dst)
    m = int(out[0])
    mean, sigma = out[1:1 + m], out[1 + m:1 + 2 * m]
    theta = None if fit else np.array([0.01, 10.0, 0.05, 1e-3])
    pm, ps, st = gp_ctx.gp_slip("rbf*brownian", t[None], s[None], theta=theta)
    assert st[0] >= 0 and pm.shape[1] == m == 599
    assert np.array_equal(pm[0], mean) and np.array_equal(ps[0], sigma)
    if not fit:
        rm, rs = go.gp_slip_callback(t, s, go.KernelExpr("rbf*brownian"), theta=theta[:-1], noise=theta[-1])
        assert np.max(np.abs(mean - rm) / np.maximum(1, np.abs(rm))) < 1e-9
        assert np.max(np.abs(sigma - rs) / np.maximum(1, np.abs(rs))) < 1e-9
