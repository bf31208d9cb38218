#!/usr/bin/env python
"""Device-timed numbers for the BASELINE.json configs that bench.py does not carry on its line.

    python tools/bench_configs.py [sweep|mc|single|callback|recorder|all] [scale]

  sweep  - configs[2]: Kernel Selection sweep, C = 64 candidates (8 kernel families x 8 hyper-parameter draws) x
           B = 4096 windows, N = 256, LML + gradient (cngp_lml_grad_batch).
  mc     - configs[3]: Monte-Carlo slip windows, N = 128, M = 600, predictive sigma + ZUPT look-ahead.  One GPU's
           shard of the 2^20 windows (131072 = 2^20 / 8), processed in slabs of 16384 windows; under torchrun every
           rank takes its own shard (window ids offset by rank), no data-path collective.
  single - configs[0]: one N = 100 window, SE kernel, predict + look-ahead, host buffers (latency of one callback).

`scale` (default 1.0) shrinks the window counts for a quick look.  Inputs are resident in HBM, timing is CUDA events
on the launching stream after one warm-up pass; one JSON line per config.  Flop figures are SURVEY.md 8d's."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from corenav_gp_b200 import synthetic as syn  # noqa: E402
from corenav_gp_b200.api import GpContext  # noqa: E402

FAMILIES = ["rbf", "matern32", "matern52", "ratquad", "rbf+stdperiodic", "rbf*brownian", "matern32+linear", "rbf+bias"]


def n_params(ctx_kernel):
    from corenav_gp_b200.api import parse_kernel
    return parse_kernel(ctx_kernel).n_params


def draw_theta(kernel, n, rng):
    """log-uniform draws: variances in [1e-3, 1], length scales / periods in [1, 100], noise in [1e-4, 1e-1]."""
    P = n_params(kernel)
    th = np.exp(rng.uniform(np.log(1e-3), np.log(1.0), (n, P + 1)))
    k = kernel
    # positions of length-scale-like parameters per family (order of corenav_gp_b200.api.parse_kernel)
    ls_pos = {"rbf": [1], "matern32": [1], "matern52": [1], "ratquad": [1], "rbf+stdperiodic": [1, 3, 4],
              "rbf*brownian": [1], "matern32+linear": [1], "rbf+bias": [1]}[k]
    for p in ls_pos:
        th[:, p] = np.exp(rng.uniform(np.log(1.0), np.log(100.0), n))
    if k == "ratquad":
        th[:, 2] = np.exp(rng.uniform(np.log(0.5), np.log(5.0), n))     # power
    th[:, P] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), n))
    return th


def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def sweep(ctx, scale):
    B, N = max(64, int(4096 * scale)), 256
    x, y = syn.slip_windows(0, B, N)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    rng = np.random.default_rng(11)
    total_ms, flop, per_family, bad = 0.0, 0.0, {}, 0
    for fam in FAMILIES:
        th = torch.from_numpy(draw_theta(fam, 8, rng)).cuda()
        P = th.shape[1]
        ms, (lml, grad, status) = timed(lambda: ctx.lml_grad(fam, th, dx, dy))
        bad += int((status < 0).sum().item())
        fin = torch.isfinite(lml[status >= 0]).all().item() and torch.isfinite(grad[status >= 0]).all().item()
        per_family[fam] = {"ms": ms, "finite": bool(fin)}
        total_ms += ms
        flop += 8 * B * (N ** 3 + 2 * N * N * (1 + P))
    return {"config": "configs[2] Kernel Selection sweep", "candidates": 64, "windows": B, "N": N,
            "ms_total": total_ms, "lml_grad_per_s": 64 * B / (total_ms * 1e-3), "tflops_fp64": flop / (total_ms * 1e-3) * 1e-12,
            "not_positive_definite": bad, "per_family": per_family}


def mc(ctx, scale, rank, world):
    shard, N, M, slab = max(1024, int(131072 * scale)), 128, 600, 16384
    slab = min(slab, shard)
    first = rank * shard
    th = torch.from_numpy(syn.theta_for("rbf+stdperiodic")).cuda()
    look = syn.lookahead_context(0.5)
    shared = {k: torch.from_numpy(np.ascontiguousarray(look[k])).cuda() for k in ("STM", "Hvec", "pos")}
    slabs = []
    for s0 in range(0, shard, slab):
        n = min(slab, shard - s0)
        x, y = syn.slip_windows(first + s0, n, N)
        mcx = syn.monte_carlo_contexts(first + s0, n)     # per-window P0 and Q: a stated fraction never triggers
        slabs.append((torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(mcx["P"]).cuda(),
                      torch.from_numpy(mcx["Q"]).cuda()))
    xs = torch.from_numpy(syn.test_grid(syn.slip_windows(first, 1, N)[0][0], M)).cuda()
    res = {"trig": 0, "steps": 0}

    def one_pass():
        trig = steps = 0
        for dx, dy, dP, dQ in slabs:
            mean, var, _, status = ctx.predict("rbf+stdperiodic", th, dx, dy, xs, want_lml=False)
            sigma = 2.0 * torch.sqrt(var)                      # gp_slip_node.py:61 (torch elementwise: plumbing)
            out = ctx.zupt_lookahead(mean, sigma, dP, dQ, shared["STM"], shared["Hvec"], shared["pos"])
            trig += int(out["triggered"].sum().item())
            steps += int(out["step_stop"].sum().item())
        res["trig"], res["steps"] = trig, steps

    ctx.set_profiling(True)
    for k in range(6):
        ctx.profile_read(k, reset=True)
    ms, _ = timed(one_pass)
    prof = {name: ctx.profile_read(k) for k, name in ((0, "fit"), (1, "var"), (3, "lookahead"))}
    ctx.set_profiling(False)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        c = torch.tensor([res["trig"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(c)
        res["trig"] = int(c[0])
    gp_flop = shard * (N ** 3 / 3 + 2 * N * N + 2 * N * M + N * N * M + 2 * N * M)
    return {"config": "configs[3] Monte-Carlo slip windows (per-GPU shard of 2^20)", "n_gpus": world,
            "windows_per_gpu": shard, "N": N, "M": M, "ms": ms, "windows_per_s": shard * world / (ms * 1e-3),
            "gp_tflops_fp64_per_gpu": gp_flop / (ms * 1e-3) * 1e-12, "triggered": res["trig"],
            "trigger_rate": res["trig"] / float(shard * world),
            "lookahead_steps_rank0": res["steps"], "mean_steps_per_window_rank0": res["steps"] / float(shard),
            "fixture": "syn.monte_carlo_contexts: per-window P0 and Q (filter quality u ~ U[0,1] scales the dynamic "
                       "states and Q by 10^(-4.5u)), horizontal sigma U[0.2,0.8] m; windows that never trigger run all "
                       "2995 steps",
            "rank0_kernel_ms_two_passes": {k: v[0] for k, v in prof.items()}}


def single(ctx):
    N, M = 100, 600
    x, y = syn.slip_windows(0, 1, N)
    xs = syn.test_grid(x[0], M)
    th = syn.theta_for("rbf")
    look = syn.lookahead_context(0.5)

    def cb():
        mean, var, lml, st = ctx.predict("rbf", th, x, y, xs)
        return ctx.zupt_lookahead(mean, 2 * np.sqrt(var), look["P"], look["Q"], look["STM"], look["Hvec"], look["pos"])
    cb()
    t0 = time.perf_counter()
    for _ in range(20):
        out = cb()
    dt = (time.perf_counter() - t0) / 20
    ctx.set_profiling(True)
    for k in range(6):
        ctx.profile_read(k, reset=True)
    for _ in range(5):
        cb()
    prof = {name: ctx.profile_read(k)[0] / 5 for k, name in ((0, "fit"), (1, "var"), (3, "lookahead"))}
    ctx.set_profiling(False)
    return {"config": "configs[0] single window N=100, SE, predict + look-ahead, host buffers", "ms_per_callback": dt * 1e3,
            "kernel_ms": prof, "triggered": int(out["triggered"][0]), "i_stop": int(out["i_stop"][0]),
            "step_stop": int(out["step_stop"][0])}


def callback_fit(ctx):
    """Rows a1-a7 end to end, as the reference runs them: the node callback with the hyper-parameter fit (L-BFGS-B from
    all-ones on softplus parameters, gp_slip_node.py:31-36), n = 149 samples (134 training points), 600-step horizon."""
    out = {}
    for B in (1, 64, 4096):
        rows = []
        for b in range(B):
            t = 21.0 + np.arange(149)
            rng = np.random.default_rng(100 + b)
            rows.append((t, 0.02 + 0.05 * np.sin(2 * np.pi * np.arange(149) / 37.0) + 0.03 * rng.standard_normal(149)))
        t = np.stack([r[0] for r in rows]); s = np.stack([r[1] for r in rows])
        ctx.gp_slip("rbf*brownian", t, s)
        t0 = time.perf_counter()
        mean, sigma, status = ctx.gp_slip("rbf*brownian", t, s)
        dt = time.perf_counter() - t0
        ctx.set_profiling(True)
        for k in range(6):
            ctx.profile_read(k, reset=True)
        ctx.gp_slip("rbf*brownian", t, s)
        prof = {name: ctx.profile_read(k) for k, name in ((0, "fit"), (1, "var"), (2, "grad"), (5, "misc"))}
        ctx.set_profiling(False)
        out[f"B={B}"] = {"ms_per_call": dt * 1e3, "ms_per_window": dt * 1e3 / B, "windows_per_s": B / dt,
                         "ok": bool((status >= 0).all()),
                         "m": int(mean.shape[1]),
                         "kernel_ms_and_launches": {k: [round(v[0], 4), int(v[1])] for k, v in prof.items()}}
    return {"config": "node callback with hyper-parameter fit (rbf*brownian, n=149, horizon 600), host buffers", **out}


def recorder(ctx, scale):
    """Row N1: slip extraction + recorder, 65536 drives x 512 updates resident in HBM (104 algorithmic bytes/update)."""
    B, T = max(256, int(65536 * scale)), 512
    chunk = 4096
    base = syn.drives(0, chunk, T=T)
    reps = B // chunk
    dev = {k: torch.from_numpy(np.tile(v, (reps,) + (1,) * (v.ndim - 1))).cuda() for k, v in base.items()}
    B = reps * chunk
    ctx.set_profiling(True)
    ctx.profile_read(5, reset=True)
    ms, out = timed(lambda: ctx.slip_record(dev["joint"], dev["att"], dev["vel"], dev["cmd"], dev["stop_cmd"],
                                            max_windows=2, cap=149), reps=3)
    kms, kn = ctx.profile_read(5)
    ctx.set_profiling(False)
    ms_call, ms = ms, kms / max(1, kn)
    byts = B * T * 104.0 + B * 2 * 149 * 16.0
    return {"config": "N1 slip extraction + GP_Input recorder", "drives": B, "updates_per_drive": T, "ms": ms,
            "updates_per_s": B * T / (ms * 1e-3), "algorithmic_GBps": byts / (ms * 1e-3) * 1e-9,
            "ms_whole_call": ms_call,
            "note": "ms = slip_record_kernel alone (CUDA events around the launch); the call also clears the outputs "
                    "(five memsets); HBM peak 6453 GB/s (MEASURED_PEAKS.json)",
            "windows_closed": int(out["n_windows"].sum().item())}


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = GpContext(device=local)
    lines = []
    if what in ("single", "all") and rank == 0:
        lines.append(single(ctx))
    if what in ("sweep", "all") and rank == 0:
        lines.append(sweep(ctx, scale))
    if what in ("callback", "all") and rank == 0:
        lines.append(callback_fit(ctx))
    if what in ("recorder", "all") and rank == 0:
        lines.append(recorder(ctx, scale))
    if what in ("mc", "all"):
        r = mc(ctx, scale, rank, world)
        if rank == 0:
            lines.append(r)
    for ln in lines:
        print(json.dumps(ln))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
