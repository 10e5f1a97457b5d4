"""Host-side plumbing of the data-parallel mode (the reference has no multi-GPU code, so the
semantics are defined here).  One process per GPU; torch.distributed carries the rendezvous, the barrier and the
max-over-ranks timing; the gradient all-reduce itself is NCCL inside libcatgen (cg_dist_allreduce_grads), bound to
the same communicator id this module broadcasts.

  * the minibatch shards by consecutive rows: rank r of R gets rows [r*B/R, (r+1)*B/R) of inputs / targets / noise;
  * parameters are replicated, so every rank applies the identical Adam update after the averaged gradient;
  * G's batch-norm statistics are per-rank (local batch) -- with B/R per rank this is NOT the single-device batch-B
    statistic; parity for DP is defined against an oracle run with the same sharding.
"""
import numpy as np


def shard_bounds(n, rank, world):
    """Rows [lo, hi) of a global batch of n owned by `rank`; n must divide evenly (adversarial.train batches are even)."""
    if n % world:
        raise ValueError("global batch %d is not divisible by %d ranks" % (n, world))
    per = n // world
    return rank * per, (rank + 1) * per


def shard(a, rank, world, axis=0):
    lo, hi = shard_bounds(a.shape[axis], rank, world)
    idx = [slice(None)] * a.ndim
    idx[axis] = slice(lo, hi)
    return np.ascontiguousarray(a[tuple(idx)])


def broadcast_bytes(payload, src, dist, device="cpu"):
    """Broadcast a fixed-size bytes object (the 128-byte ncclUniqueId) from `src` to every rank."""
    import torch
    n = len(payload) if payload is not None else 0
    ln = torch.tensor([n], dtype=torch.int64, device=device)
    dist.broadcast(ln, src)
    buf = torch.zeros(int(ln.item()), dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def max_over_ranks(value, dist, device="cpu"):
    """Device time of a multi-GPU step is the MAX over ranks (never a wall clock)."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def average_gradients_reference(grads_per_rank):
    """What cg_dist_allreduce_grads computes, stated on the host for tests: sum over ranks scaled by 1/R (each rank's
    BCE is a mean over its LOCAL batch, so the average of the local-mean gradients is the global-mean gradient)."""
    return np.mean(np.stack(grads_per_rank), axis=0)
