"""world_size-2 gloo tests of the data-parallel host logic (no GPU): sharding, id broadcast, max-over-ranks, and the
averaging identity the NCCL all-reduce relies on, checked with the oracle on a sharded batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from catgen import dist as cgd


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        payload = bytes(range(128)) if rank == 0 else None
        got = cgd.broadcast_bytes(payload, 0, dist)
        mx = cgd.max_over_ranks(10.0 + rank, dist)
        x = np.arange(8 * 3, dtype=np.float32).reshape(8, 3)
        mine = cgd.shard(x, rank, world)
        q.put((rank, got == bytes(range(128)), mx, mine.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert all(r[1] for r in res), "ncclUniqueId-sized payload arrives intact on every rank"
    assert all(r[2] == 11.0 for r in res), "timing is the max over ranks"
    rows = res[0][3] + res[1][3]
    assert rows == np.arange(24, dtype=np.float32).reshape(8, 3).tolist(), "shards are disjoint, ordered and cover the batch"


def test_shard_rejects_uneven_batches():
    with pytest.raises(ValueError):
        cgd.shard_bounds(10, 0, 4)


def test_average_of_local_mean_gradients_is_the_global_gradient():
    """The identity behind 'all-reduce sum then scale by 1/R': with a per-sample network (D in eval mode, no batch norm)
    and BCE averaged over the LOCAL batch, the mean of the per-shard gradients equals the full-batch gradient."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(0)
    B, R = 8, 2
    d = po.Model(po.D32_ST3, 3, 100, seed=2)
    p = d.params; p += rng.standard_normal(p.size).astype(np.float32) * 0.01
    x = rng.uniform(0, 1, (B, 3, 32, 32)).astype(np.float32)
    t = (rng.random(B) > 0.5).astype(np.float32)
    L = po.lib()

    def grad(xs, ts):
        n = xs.shape[0]
        sig, _ = d.D_forward(xs, None)
        df = np.empty(n, np.float32); L.og_bce_bwd(po.P(sig), po.P(np.ascontiguousarray(ts)), po.P(df), n)
        d.zero_grads(); d.D_backward(df)
        return d.grads.copy()

    full = grad(x, t)
    parts = [grad(cgd.shard(x, r, R), cgd.shard(t, r, R)) for r in range(R)]
    avg = cgd.average_gradients_reference(parts)
    assert np.abs(avg - full).max() <= 2e-5 * np.abs(full).max() + 1e-9
