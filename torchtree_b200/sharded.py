"""Pattern sharding across GPUs (one process per GPU, torch.distributed).

Site patterns are conditionally independent given the tree, so rank g owns the
patterns [g*N/G, (g+1)*N/G) for all nodes and evaluates them with its own engine;
the only exchange is a sum all-reduce of the scalar log-likelihood (forward) and
of the gradient vector (backward) -- SURVEY 8(e).  The two autograd functions
below are the usual pair: identity/all-reduce around the sharded region, so that
every rank ends up with the full value and the full gradient.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(pattern_count: int, rank: int, world_size: int):
    """Contiguous, balanced slice [lo, hi) of the patterns owned by `rank`."""
    per = (pattern_count + world_size - 1) // world_size
    lo = min(pattern_count, rank * per)
    hi = min(pattern_count, lo + per)
    return lo, hi


class _CopyToShards(torch.autograd.Function):
    """Forward: identity (parameters are replicated).  Backward: sum of the
    per-shard gradients over all ranks."""

    @staticmethod
    def forward(ctx, group, *tensors):
        ctx.group = group
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        live = [g for g in grads if g is not None]
        if live and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            flat = torch.cat([g.reshape(-1) for g in live])
            dist.all_reduce(flat, group=ctx.group)
            out, off = [], 0
            for g in grads:
                if g is None:
                    out.append(None)
                else:
                    n = g.numel()
                    out.append(flat[off:off + n].view_as(g))
                    off += n
            return (None,) + tuple(out)
        return (None,) + tuple(grads)


class _SumOverShards(torch.autograd.Function):
    """Forward: all-reduce (sum) of the per-shard log-likelihoods.  Backward: identity."""

    @staticmethod
    def forward(ctx, group, lnl):
        out = lnl.clone()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        return None, grad


def sharded_log_likelihood(local_fn, tensors, group=None):
    """lnL of the whole alignment from per-shard evaluations.

    `local_fn(*tensors) -> lnL[D]` evaluates this rank's pattern shard (e.g.
    `lambda *a: log_likelihood_eigen(engine, *a)` with an engine built on
    `shard_range(...)`).  Returns the full lnL on every rank; after `.backward()`
    every rank holds the full gradient of every tensor in `tensors`.
    """
    copies = _CopyToShards.apply(group, *tensors)
    return _SumOverShards.apply(group, local_fn(*copies))
