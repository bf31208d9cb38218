#!/usr/bin/env python
"""Time the large-window factorisation (BASELINE.json configs[4]) on one GPU or, under torchrun, block-cyclic over N GPUs.

    python tools/bench_large.py [N=32768] [reps=3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_large.py 32768 3

Prints one JSON line: seconds per factorisation (CUDA events, max over ranks), FP64 TFLOP/s on N^3/3, and the residual
||Ky alpha - y|| / ||y|| evaluated with cngp_large_matvec."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from corenav_gp_b200 import large, synthetic as syn  # noqa: E402
from corenav_gp_b200.api import GpContext  # noqa: E402


def run(ctx, N, reps, rank, world, lookahead=True):
    """Factor one N-point window `reps` + 1 times (first one untimed) on `world` ranks; returns the result dict on every
    rank (timings are the max over ranks)."""
    if world > 1:
        import torch.distributed as dist
    kname = "rbf+stdperiodic"
    th = np.array([0.01, 10.0, 0.0025, 37.0, 1.0, 1e-2])
    x, y = syn.slip_windows(5, 1, N)
    win = large.LargeWindow(ctx, kname, th, x[0], y[0], rank=rank, world=world)
    times = []
    host_ms = []
    out = None
    for it in range(reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.set_profiling(it == reps)
        e0.record()
        out = large.chol_large_distributed(win, rank, world, want_alpha=True, lookahead=lookahead,
                                           profile=(it == reps and bool(os.environ.get("CNGP_LARGE_PHASES"))))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t[0]))
            host_ms.append(out.get("host_enqueue_ms"))
    prof_ms, prof_n = ctx.profile_read(4)
    ctx.set_profiling(False)
    r = ctx.large_matvec(kname, th, win.x, out["alpha"].contiguous())
    res = float(torch.linalg.norm(r - win.y) / torch.linalg.norm(win.y))
    best = min(times)
    del win
    return {"N": N, "n_gpus": world, "lookahead": lookahead, "ms": times, "best_ms": best,
            "median_ms": float(np.median(times)), "chunk_rows": int(os.environ.get("CNGP_LARGE_CHUNK_ROWS", "0")),
            "tflops_n3_over_3": N ** 3 / 3.0 / (best * 1e-3) * 1e-12, "lml": out["lml"],
            "logdet": out["logdet"], "quad": out["quad"], "residual": res,
            "rank0_kernel_ms_last_rep": prof_ms, "rank0_launches_last_rep": prof_n,
            "host_enqueue_ms_per_rep": host_ms,
            "phase_ms_rank": {"rank": rank, **(out.get("phase_ms") or {})}}


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lookahead = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = GpContext(device=local)
    line = run(ctx, N, reps, rank, world, lookahead)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
