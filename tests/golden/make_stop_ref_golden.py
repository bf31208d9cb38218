"""Generate tests/golden/stop_ref_golden.npz: inputs and outputs of the reference's OWN GpPredictor::GPCallBack.

Run in the authoring container (where /root/reference exists):  python tests/golden/make_stop_ref_golden.py
The outputs come from oracle/_ref - /root/reference/gp_predictor/src/gp_predictor.cpp compiled unmodified against the
stand-in headers of oracle/ref_stubs/ (see oracle/ref_gp_predictor.py) - NOT from oracle/stop_oracle.c, so the fixture
pins the restatement and the CUDA kernel to the reference's code.  /root/reference cannot travel to the GPU box; this
file can.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from corenav_gp_b200 import synthetic as syn   # noqa: E402  (input generators only: contexts, no GP arithmetic)
from oracle import ref_gp_predictor as rg      # noqa: E402


def make(B=48, M=240, seed=2024):
    rng = np.random.default_rng(seed)
    k = np.arange(M)
    # GP_Output-like arrays (Kernel Selection/docs/ours.jpg: mean relaxing to ~0, band widening), with a few windows of
    # large slip / large sigma so that the UT covariance term beats the floors of gp_predictor.cpp:80-83
    mean = 0.08 * np.exp(-k / 70.0)[None, :] * rng.uniform(-1, 1, (B, 1)) + 0.01 * rng.standard_normal((B, M))
    sigma = 2.0 * np.sqrt(1e-3 + 0.01 * (1 - np.exp(-k / 150.0)))[None, :] * rng.uniform(0.6, 1.4, (B, 1))
    mean[::5] += 0.35
    sigma[::6] *= 2.5
    s = rng.uniform(0.15, 0.72, B)                      # horizontal 1-sigma of P0: some windows never trigger
    ctx = syn.lookahead_context(s)
    pos = np.broadcast_to(ctx["pos"], (B, 3)).copy() if ctx["pos"].ndim == 1 else ctx["pos"].copy()
    pos += rng.normal(0, [2e-6, 2e-6, 3.0], (B, 3))
    arr = {n: (np.broadcast_to(ctx[n], (B,) + ctx[n].shape).copy() if ctx[n].ndim == 1 else ctx[n].copy())
           for n in ("P", "Q", "STM", "Hvec")}
    out = dict(triggered=np.zeros(B, np.int32), i_stop=np.zeros(B, np.int32), step_stop=np.zeros(B, np.int32),
               xy_err=np.zeros(B), P_final=np.zeros((B, 225)), K_final=np.zeros((B, 60)), R_final=np.zeros((B, 16)),
               stop_cmd=np.zeros(B), trace0=None)
    for b in range(B):
        r = rg.gp_callback(mean[b], sigma[b], arr["P"][b], arr["Q"][b], arr["STM"][b], arr["Hvec"][b], pos[b])
        out["triggered"][b] = r["triggered"]
        out["i_stop"][b] = r["i_stop"]
        out["step_stop"][b] = r["step_stop"]
        out["xy_err"][b] = r["xy_err"]
        out["P_final"][b] = r["P"].ravel()
        out["K_final"][b] = r["K"].ravel()
        out["R_final"][b] = r["R"].ravel()
        out["stop_cmd"][b] = r["stop_cmd"]
        if out["trace0"] is None and r["n_steps"] > 300:
            out["trace0"] = r["xy_trace"]
            out["trace0_window"] = np.int32(b)
    enu_in = syn.INIT_LLH[None, :] + rng.normal(0, [1e-5, 1e-5, 5.0], (16, 3))
    enu_out = np.stack([rg.llh_to_enu(*p) for p in enu_in])
    return dict(mean=mean, sigma=sigma, pos=pos, **arr, **out, enu_in=enu_in, enu_out=enu_out,
                init_llh=np.array(rg.INIT_LLH), init_ecef=np.array(rg.INIT_ECEF))


if __name__ == "__main__":
    d = make()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stop_ref_golden.npz")
    np.savez_compressed(path, **d)
    print(path, os.path.getsize(path), "bytes; triggered", int(d["triggered"].sum()), "of", d["triggered"].size)
