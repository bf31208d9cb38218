"""The multi-rank driver of the large-window factorisation (corenav_gp_b200/large.py, SURVEY.md section 8e, cfg 5) under
world_size-2 gloo on CPU.  The per-rank engine is a TEST DOUBLE (dense torch-CPU block columns, same interface as
large.LargeWindow) - what is exercised here is the driver's schedule: owners, look-ahead order, double-buffered panel
broadcasts, the reductions and the backward sweep.  The result must equal a direct LAPACK Cholesky (the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NB = 256


class DenseEngine:
    """Block-cyclic dense engine on CPU tensors.  Block column c = rows [c NB, n) x NB columns, owned by c % world."""

    def __init__(self, K, y, rank, world):
        n = K.shape[0]
        assert n % NB == 0
        self.N, self.n_pad, self.nb = n, n, NB
        self.rank, self.world = rank, world
        self.n_blockcols = n // NB
        self.K, self.y = K, y
        self.cols = {}
        self.diag_inv = {}
        self.logdet_parts = {}
        self.panels = [torch.zeros(n + 1, NB, dtype=torch.float64), torch.zeros(n + 1, NB, dtype=torch.float64)]
        self.alpha = torch.zeros(n, dtype=torch.float64)
        self.calls = []

    def assemble(self):
        for c in range(self.rank, self.n_blockcols, self.world):
            blk = torch.cat([self.K[:, c * NB:(c + 1) * NB], self.y[None, c * NB:(c + 1) * NB]], 0)   # y as extra row
            self.cols[c] = blk.clone()

    def panel_payload(self, k):
        return self.panels[k % 2][(k + 1) * NB:]

    def factor_panel(self, k):
        assert k % self.world == self.rank
        self.calls.append(("factor", k))
        blk = self.cols[k]
        Lkk = torch.linalg.cholesky(blk[k * NB:(k + 1) * NB])
        self.logdet_parts[k] = 2.0 * torch.log(torch.diagonal(Lkk)).sum()
        Winv = torch.linalg.inv(Lkk)
        self.diag_inv[k] = Winv
        below = blk[(k + 1) * NB:] @ Winv.T
        blk[(k + 1) * NB:] = below
        self.panels[k % 2][(k + 1) * NB:] = below

    def update(self, k, c_lo, c_hi):
        self.calls.append(("update", k, c_lo, c_hi))
        P = self.panels[k % 2]
        for c in range(max(c_lo, k + 1), min(c_hi, self.n_blockcols)):
            if c % self.world == self.rank:
                self.cols[c][c * NB:] -= P[c * NB:] @ P[c * NB:(c + 1) * NB].T

    def reduce(self):
        z = torch.zeros(self.n_pad, dtype=torch.float64)
        for c, blk in self.cols.items():
            z[c * NB:(c + 1) * NB] = blk[-1]
        self.z = z
        sums = torch.tensor([float(sum(self.logdet_parts.values())), float((z * z).sum()), 0.0], dtype=torch.float64)
        return z, sums

    def backsolve_step(self, j):
        blk = self.cols[j]
        t = self.z[j * NB:(j + 1) * NB] - blk[(j + 1) * NB:-1].T @ self.alpha[(j + 1) * NB:]
        self.alpha[j * NB:(j + 1) * NB] = self.diag_inv[j].T @ t

    # the lazy sweep (large.py prefers it when the engine has it)
    def backsolve_begin(self):
        self.s_acc = torch.zeros(self.n_pad, dtype=torch.float64)

    def backsolve_finish(self, j):
        self.calls.append(("finish", j))
        t = self.z[j * NB:(j + 1) * NB] - self.s_acc[j * NB:(j + 1) * NB]
        self.alpha[j * NB:(j + 1) * NB] = self.diag_inv[j].T @ t

    def backsolve_apply(self, j):
        for c, blk in self.cols.items():
            if c < j:
                self.s_acc[c * NB:(c + 1) * NB] += blk[j * NB:(j + 1) * NB].T @ self.alpha[j * NB:(j + 1) * NB]

    def alpha_block(self, j):
        return self.alpha[j * NB:(j + 1) * NB]


def _problem(n):
    rng = np.random.default_rng(3)
    x = np.arange(n, dtype=np.float64) + 20.0
    d = x[:, None] - x[None, :]
    K = 0.01 * np.exp(-0.5 * d * d / 100.0) + 0.0025 * np.exp(-0.5 * np.sin(np.pi * d / 37.0) ** 2) + 1e-2 * np.eye(n)
    y = 0.05 * np.sin(2 * np.pi * x / 37.0) + 0.03 * rng.standard_normal(n)
    return torch.from_numpy(K), torch.from_numpy(y)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, lookahead, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corenav_gp_b200 import large
    K, y = _problem(n)
    eng = DenseEngine(K, y, rank, world)
    out = large.chol_large_distributed(eng, rank, world, coll=large.TorchCollectives(), lookahead=lookahead)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), logdet=out["logdet"], quad=out["quad"], lml=out["lml"],
             alpha=out["alpha"].numpy(), n_factor=len([c for c in eng.calls if c[0] == "factor"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lookahead", [True, False])
def test_two_rank_gloo_driver_matches_lapack(tmp_path, lookahead):
    world, n = 2, 5 * NB
    mp.spawn(_worker, args=(world, _free_port(), n, lookahead, str(tmp_path)), nprocs=world, join=True)
    K, y = _problem(n)
    from scipy.linalg import cho_factor, cho_solve
    c = cho_factor(K.numpy(), lower=True)
    alpha = cho_solve(c, y.numpy())
    logdet = 2.0 * np.sum(np.log(np.diag(c[0])))
    quad = float(y.numpy() @ alpha)
    lml = 0.5 * (-n * np.log(2 * np.pi) - logdet - quad)
    n_factor = 0
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert abs(got["logdet"] - logdet) < 1e-9 * abs(logdet)
        assert abs(got["quad"] - quad) < 1e-9 * abs(quad)
        assert abs(got["lml"] - lml) < 1e-9 * abs(lml)
        assert np.max(np.abs(got["alpha"] - alpha)) < 1e-9 * np.max(np.abs(alpha))
        n_factor += int(got["n_factor"])
    assert n_factor == n // NB            # every block column factored exactly once, by its owner


def test_single_rank_driver_no_collectives():
    from corenav_gp_b200 import large
    n = 3 * NB
    K, y = _problem(n)
    eng = DenseEngine(K, y, 0, 1)
    out = large.chol_large_distributed(eng, 0, 1)
    alpha = np.linalg.solve(K.numpy(), y.numpy())
    assert np.max(np.abs(out["alpha"].numpy() - alpha)) < 1e-9 * np.max(np.abs(alpha))
    # look-ahead order: block column k+1 is updated and factored before the rest of update k
    i_f1 = eng.calls.index(("factor", 1))
    assert eng.calls[i_f1 - 1] == ("update", 0, 1, 2) and eng.calls[i_f1 + 1] == ("update", 0, 2, 3)


def test_plan_block_cyclic_counts():
    from corenav_gp_b200 import large
    for N, world in ((32768, 8), (1300, 8), (2049, 3), (255, 1)):
        tot = 0
        for r in range(world):
            p = large.make_plan(N, world, r)
            assert p.n_pad % NB == 0 and p.n_pad >= N and p.n_pad - N < NB
            assert p.row_tiles == p.n_pad // 8 + 16
            tot += p.n_local_blockcols
            assert p.local_doubles == p.n_local_blockcols * 32 * p.row_tiles * 64
        assert tot == p.n_blockcols
