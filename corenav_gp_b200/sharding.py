"""Window sharding across the GPUs of one box.

Slip windows (and candidate x window pairs) are independent, so the hot path has NO data-path collective: rank r of
W takes the contiguous range shard_range(total, r, W), computes it on its own GPU through its own cngp context, and the
only communication is one all-gather of the results afterwards (C1) - NCCL over NVLink when the tensors live on the
GPUs, gloo in the CPU tests.  torch.distributed is plumbing here; nothing in this module computes.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of rank's contiguous share; the first total % world ranks get one extra window."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def all_gather_rows(local, total: int, rank: int, world: int):
    """All-gather per-window results sharded by shard_range along dim 0 -> the full [total, ...] array on every rank.

    `local` is a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array (gathered through CPU tensors)."""
    if world == 1:
        return local
    as_numpy = isinstance(local, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local)) if as_numpy else local.contiguous()
    counts = [shard_range(total, r, world)[1] for r in range(world)]
    assert t.shape[0] == counts[rank], (t.shape, counts, rank)
    cap = max(counts)
    if t.shape[0] < cap:                                   # pad so every rank contributes the same shape
        pad = torch.zeros((cap - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out.numpy() if as_numpy else out


def run_sharded(compute: Callable[[int, int], Dict[str, object]], total: int, rank: int, world: int) -> Dict[str, object]:
    """compute(first, count) -> dict of per-window arrays for windows [first, first+count); returns the dict of full
    arrays, identical on every rank and identical to compute(0, total) (the windows are independent)."""
    first, count = shard_range(total, rank, world)
    local = compute(first, count)
    return {k: all_gather_rows(v, total, rank, world) for k, v in local.items()}


def predict_sharded(ctx, kernel, theta, x_of: Callable[[int, int], tuple], xstar, total: int, rank: int, world: int):
    """BASELINE.json configs[1]/[3] across ranks: x_of(first, count) -> (x, y) for that range (e.g.
    synthetic.slip_windows); every rank gets mean, var, lml, status of all windows."""
    def compute(first, count):
        x, y = x_of(first, count)
        mean, var, lml, status = ctx.predict(kernel, theta, x, y, xstar)
        return dict(mean=mean, var=var, lml=lml, status=status)
    return run_sharded(compute, total, rank, world)
