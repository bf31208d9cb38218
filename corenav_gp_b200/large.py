"""One large slip window (BASELINE.json configs[4]: N = 32768) factored across the GPUs of one box.

Same inference as the batched windows (row a3 of SURVEY.md section 8: GPy ExactGaussianInference as reached from
core_navigation/script/gp_slip_node.py:35) - Ky = K + (noise + 1e-8) I, Cholesky, log det, y' Ky^-1 y, LML, alpha - for
a window far beyond what the reference ever runs.  1-D block-cyclic column distribution: block column c (256 columns)
lives on rank c % world.  Per block column k

    owner(k):   factor the diagonal block, form the panel                    cngp_large_factor_panel   (CUDA)
    all ranks:  broadcast of the panel (C2 - the only data-path collective)   torch.distributed / NCCL over NVLink
    all ranks:  A(:, own columns > k) -= panel panel^T                        cngp_large_update         (CUDA, FP64 DMMA)

with a look-ahead of one block column: the owner of k+1 updates that block column first, factors it and starts its
broadcast while every rank is still busy with the rest of update k, so the broadcast and the (latency-bound) diagonal
factorisation hide behind the trailing update.  y rides along as an extra matrix row, so after the last panel each rank
holds its slice of z = L^-1 y; one all-reduce gives log det, z'z and z, and alpha = L^-T z is a backward sweep with
one 2 KB broadcast per block column.

`engine` is anything with the LargeWindow methods below - the CUDA one here, or a test double (tests/ uses a numpy
engine to exercise this driver under world_size-2 gloo on CPU).  Nothing in this module computes.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np

from . import _lib as L
from .api import CngpError, GpContext, parse_kernel

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

LOG_2PI = math.log(2.0 * math.pi)


def make_plan(N: int, world: int = 1, rank: int = 0) -> L.LargePlan:
    p = L.LargePlan()
    rc = L.load().cngp_large_make_plan(N, world, rank, C.byref(p))
    if rc != 0:
        raise CngpError(f"cngp_large_make_plan(N={N}, world={world}, rank={rank}) failed (rc={rc})")
    return p


class LargeWindow:
    """This rank's share of one large window on its GPU: block columns, two panel buffers, inverted diagonal blocks."""

    def __init__(self, ctx: GpContext, kernel, theta, x, y, rank: int = 0, world: int = 1):
        self.ctx, self.lib = ctx, ctx.lib
        self.kernel = parse_kernel(kernel)
        self.theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        if self.theta.size != self.kernel.n_params + 1:
            raise CngpError(f"theta has {self.theta.size} entries, kernel needs {self.kernel.n_params + 1} (noise last)")
        dev = torch.device(f"cuda:{ctx.device}")
        self.x = torch.as_tensor(x, dtype=torch.float64).to(dev).contiguous()
        self.y = torch.as_tensor(y, dtype=torch.float64).to(dev).contiguous()
        self.N = int(self.x.numel())
        self.rank, self.world = rank, world
        self.plan = make_plan(self.N, world, rank)
        p = self.plan
        self.n_blockcols, self.nb, self.n_pad = int(p.n_blockcols), L.LARGE_NB, int(p.n_pad)
        f64 = dict(dtype=torch.float64, device=dev)
        self.A = torch.empty(max(1, p.local_doubles), **f64)
        self.panels = [torch.zeros(p.panel_doubles, **f64), torch.zeros(p.panel_doubles, **f64)]
        self.winv = torch.zeros(max(1, p.winv_doubles), **f64)
        self.logdet = torch.zeros(p.n_blockcols, **f64)
        self.status = torch.zeros(p.n_blockcols, dtype=torch.int32, device=dev)
        self.z = torch.zeros(p.n_pad, **f64)
        self.alpha = torch.zeros(p.n_pad, **f64)
        self.sums = torch.zeros(4, **f64)

    def _bind(self):
        self.ctx._bind_stream(True)

    def _chk(self, rc, what):
        self.ctx._check(rc, what)

    # ---- the engine interface used by chol_large_distributed ----
    def panel_buffer(self, k: int):
        return self.panels[k % 2]

    def panel_payload(self, k: int):
        """The contiguous part of panel k's buffer the other ranks need (rows below the diagonal block)."""
        rows = int(self.plan.row_tiles) - (k + 1) * (self.nb // 8)
        return self.panels[k % 2][: (self.nb // 8) * rows * 64]

    def assemble(self):
        self._bind()
        self._chk(self.lib.cngp_large_assemble(self.ctx.h, C.byref(self.plan), C.byref(self.kernel), self.theta.ctypes.data,
                                               self.x.data_ptr(), self.y.data_ptr(), self.A.data_ptr()),
                  "cngp_large_assemble")

    def factor_panel(self, k: int):
        self._bind()
        self._chk(self.lib.cngp_large_factor_panel(self.ctx.h, C.byref(self.plan), self.A.data_ptr(), k,
                                                   self.panels[k % 2].data_ptr(), self.winv.data_ptr(),
                                                   self.logdet.data_ptr(), self.status.data_ptr()),
                  "cngp_large_factor_panel")

    def update(self, k: int, c_lo: int, c_hi: int):
        self._bind()
        self._chk(self.lib.cngp_large_update(self.ctx.h, C.byref(self.plan), self.A.data_ptr(), k,
                                             self.panels[k % 2].data_ptr(), c_lo, c_hi), "cngp_large_update")

    def reduce(self):
        """-> (z [n_pad] with this rank's columns, sums [3] = logdet part, z'z part, first failing pivot or 0)."""
        self._bind()
        self._chk(self.lib.cngp_large_reduce(self.ctx.h, C.byref(self.plan), self.A.data_ptr(), self.logdet.data_ptr(),
                                             self.status.data_ptr(), self.z.data_ptr(), self.sums.data_ptr()),
                  "cngp_large_reduce")
        return self.z, self.sums[:3]

    def backsolve_step(self, j: int):
        self._bind()
        self._chk(self.lib.cngp_large_backsolve_step(self.ctx.h, C.byref(self.plan), self.A.data_ptr(),
                                                     self.winv.data_ptr(), j, self.z.data_ptr(), self.alpha.data_ptr()),
                  "cngp_large_backsolve_step")

    def alpha_block(self, j: int):
        return self.alpha[j * self.nb:(j + 1) * self.nb]


class TorchCollectives:
    """broadcast / all-reduce through torch.distributed (NCCL for CUDA tensors, gloo for CPU tensors)."""

    def __init__(self, group=None):
        self.group = group

    def broadcast_async(self, t, src: int):
        return dist.broadcast(t, src=src, group=self.group, async_op=True)

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_max(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)


class NoCollectives:
    """world == 1."""

    class _Done:
        def wait(self):
            return True

    def broadcast_async(self, t, src: int):
        return self._Done()

    def all_reduce_sum(self, t):
        pass

    def all_reduce_max(self, t):
        pass


def chol_large_distributed(engine, rank: int, world: int, coll=None, want_alpha: bool = True,
                           lookahead: bool = True) -> Dict[str, object]:
    """Run the blocked factorisation of `engine`'s window over `world` ranks (SPMD: every rank calls this).

    Returns dict(logdet, quad, lml, pivot[, alpha]) - identical on every rank; pivot = 0, or the first (1-based) pivot
    at which the matrix was found not positive definite (then the other values are NaN)."""
    if coll is None:
        coll = TorchCollectives() if world > 1 else NoCollectives()
    nblk = engine.n_blockcols
    engine.assemble()
    if rank == 0 % world:
        engine.factor_panel(0)
    pending = coll.broadcast_async(engine.panel_payload(0), src=0)
    for k in range(nblk):
        pending.wait()                                           # panel k is here
        nxt = k + 1
        if nxt < nblk:
            if lookahead:
                if nxt % world == rank:                          # bring block column k+1 up to date first and factor it
                    engine.update(k, nxt, nxt + 1)
                    engine.factor_panel(nxt)
                pending = coll.broadcast_async(engine.panel_payload(nxt), src=nxt % world)   # overlaps the update below
                engine.update(k, nxt + 1, nblk)
            else:
                engine.update(k, nxt, nblk)
                if nxt % world == rank:
                    engine.factor_panel(nxt)
                pending = coll.broadcast_async(engine.panel_payload(nxt), src=nxt % world)
    z, sums = engine.reduce()
    # pivot: smallest non-zero over ranks  ->  max of (BIG - pivot)
    BIG = 1.0e15
    enc = sums.clone()
    enc[2] = torch.where(sums[2] > 0, BIG - sums[2], torch.zeros_like(sums[2]))
    both = enc[:2].clone()
    coll.all_reduce_sum(both)
    piv = enc[2:3].clone()
    coll.all_reduce_max(piv)
    coll.all_reduce_sum(z)
    logdet, quad = float(both[0]), float(both[1])
    pivot = int(round(BIG - float(piv[0]))) if float(piv[0]) > 0 else 0
    res: Dict[str, object] = dict(pivot=pivot)
    if pivot:
        nan = float("nan")
        res.update(logdet=nan, quad=nan, lml=nan)
        return res
    res.update(logdet=logdet, quad=quad, lml=0.5 * (-engine.N * LOG_2PI - logdet - quad))
    if want_alpha:
        for j in range(nblk - 1, -1, -1):
            if j % world == rank:
                engine.backsolve_step(j)
            coll.broadcast_async(engine.alpha_block(j), src=j % world).wait()
        res["alpha"] = engine.alpha[: engine.N]
    return res


def chol_large(ctx: GpContext, kernel, theta, x, y, rank: int = 0, world: int = 1, coll=None, want_alpha: bool = True,
               lookahead: bool = True) -> Dict[str, object]:
    """Convenience: build this rank's LargeWindow and run the distributed factorisation."""
    win = LargeWindow(ctx, kernel, theta, x, y, rank=rank, world=world)
    out = chol_large_distributed(win, rank, world, coll=coll, want_alpha=want_alpha, lookahead=lookahead)
    out["window"] = win
    return out
