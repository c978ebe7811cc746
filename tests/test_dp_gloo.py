"""World-size-2 gloo test of the data-parallel host logic (no GPU): sharding + one flat all-reduce reproduces the
global-batch gradient of a mean loss; the RAdam rectification constants match the oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zeggs_b200 import dp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    X = torch.from_numpy(rs.randn(16, 7).astype(np.float32))          # 16 independent "windows"
    w = torch.from_numpy(rs.randn(7).astype(np.float32)).requires_grad_(True)
    lo, hi = dp.shard_range(16, rank, world)
    loss = (X[lo:hi] @ w).abs().mean()                                 # an L1 mean like train.py:340-421
    (g,) = torch.autograd.grad(loss, w)
    flat = g.clone()
    dp.allreduce_sum_(flat)
    g_dp = flat / world                                                # = optimizer.grad_scale
    g2 = dp.mean_of_means_is_global_mean(g, hi - lo)
    w2 = w.detach().clone().requires_grad_(True)
    (g_full,) = torch.autograd.grad((X @ w2).abs().mean(), w2)
    q.put((rank, float((g_dp - g_full).abs().max()), float((g2 - g_full).abs().max()), (lo, hi)))
    dist.destroy_process_group()


def test_dp_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][3] == (0, 8) and res[1][3] == (8, 16)
    for _, e1, e2, _ in res:
        assert e1 <= 1e-6 and e2 <= 1e-6


def test_shard_range_covers_everything():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
