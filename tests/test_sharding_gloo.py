"""Host-side multi-process logic on CPU: two gloo ranks shard the windows, each computes its
contiguous range, one all-gather reassembles - and the result equals the single-process result.  The per-shard compute
is the ORACLE here (no GPU in this container); on the GPU box the same run_sharded drives GpContext.predict."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from corenav_gp_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    for total in (0, 1, 7, 4096, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, N, M, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from corenav_gp_b200 import synthetic as syn
    from oracle import gp_oracle as go
    e = go.KernelExpr("rbf")
    th = syn.theta_for("rbf")

    def compute(first, count):
        x, y = syn.slip_windows(first, count, N)
        mean = np.empty((count, M))
        var = np.empty((count, M))
        lml = np.empty(count)
        for b in range(count):
            xs = syn.test_grid(x[b], M)
            inf = go.inference(e, th[:-1], th[-1], x[b], y[b])
            mean[b], var[b] = go.predict(e, th[:-1], th[-1], x[b], y[b], xs, inf)
            lml[b] = inf.lml
        return dict(mean=mean, var=var, lml=lml, status=np.zeros(count, dtype=np.int32))

    out = sharding.run_sharded(compute, total, rank, world)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), **out)
    if rank == 0:
        np.savez(os.path.join(outdir, "single.npz"), **compute(0, total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 8])
def test_two_rank_gloo_equals_single_process(tmp_path, total):
    world, N, M = 2, 24, 9
    mp.spawn(_worker, args=(world, _free_port(), total, N, M, str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npz")
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        for k in ("mean", "var", "lml", "status"):
            assert got[k].shape[0] == total
            assert np.array_equal(got[k], single[k]), (r, k)
