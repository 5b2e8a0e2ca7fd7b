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


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum all-reduce of `t` over the ranks of `group`.  torchtree keeps its Parameters
    (and therefore lnL and the gradients) on the host while NCCL only moves device memory: a host
    tensor goes through one device staging buffer -- host -> device, ncclAllReduce over
    NVLink, device -> host (the payload is the scalar lnL or the ~16 KB packed gradient)."""
    if _world(group) <= 1:
        return t
    backend = dist.get_backend(group)
    if t.is_cuda or "nccl" not in backend:
        dist.all_reduce(t, group=group)
        return t
    staged = t.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    dist.all_reduce(staged, group=group)
    t.copy_(staged)
    return t


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
        if live and _world(ctx.group) > 1:
            flat = torch.cat([g.reshape(-1) for g in live])
            all_reduce_sum(flat, ctx.group)
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
        return all_reduce_sum(lnl.clone(), group)

    @staticmethod
    def backward(ctx, grad):
        return None, grad


def sharded_log_likelihood(local_fn, tensors, group=None):
    """lnL of the whole alignment from per-shard evaluations.

    `local_fn(*tensors) -> lnL[D]` evaluates this rank's pattern shard (e.g.
    `lambda *a: log_likelihood_eigen(engine, *a)` with an engine built on
    `shard_range(...)`; None if the rank owns no patterns).  Returns the full lnL on every
    rank; after `.backward()` every rank holds the full gradient of every tensor in `tensors`.
    """
    copies = _CopyToShards.apply(group, *tensors)
    if local_fn is None:
        # this rank owns no patterns (more ranks than patterns): contributes 0, joins the collectives
        draws = max(int(t.shape[0]) for t in copies)
        local = sum((t.sum() * 0.0 for t in copies)) + torch.zeros(
            draws, dtype=copies[0].dtype, device=copies[0].device)
    else:
        local = local_fn(*copies)
    return _SumOverShards.apply(group, local)


# ---- pattern sharding, engine-aware fast path --------------------------------------------------
class _ShardedEngineLikelihood(torch.autograd.Function):
    """`sharded_log_likelihood(lambda *a: log_likelihood_eigen(engine, *a), tensors)` with the
    collectives moved onto the device: the shard's lnL and its packed gradient vector
    (`ttb2_grad_eigen_packed`: written in one piece by the engine's output kernels) are
    all-reduced over NCCL where they are produced and cross to the host once, already summed --
    instead of device -> host -> device -> NCCL -> host per collective.  Host or device tensors;
    every rank ends with the full value and the full gradients."""

    @staticmethod
    def forward(ctx, engine, group, general, bls, rates, props, q, freqs):
        from .function import reversible_eigensystem

        dev = torch.device("cuda", engine.device)
        src = (bls, rates, props, q, freqs)
        if all(not t.is_cuda for t in src):
            # host tensors (the stock torchtree set-up): one packed host -> device copy, then views
            flat = torch.cat([t.detach().reshape(-1).to(torch.float64) for t in src]).to(dev)
            ins, off = [], 0
            for t in src:
                ins.append(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()
        else:
            ins = [t.detach().to(dev, torch.float64).contiguous() for t in src]

        def run():
            S = engine.S
            if general:
                return engine.loglik_expm(*ins)
            if S > 8 and max(ins[3].shape[0], ins[4].shape[0]) < 6:
                # one large generator: LAPACK on the host, as the torch extension does
                evec, ivec, evals = reversible_eigensystem(q.detach().double().cpu(),
                                                           freqs.detach().double().cpu())
                return engine.loglik_eigen(ins[0], ins[1], ins[2], evec.to(dev), ivec.to(dev),
                                           evals.to(dev), ins[4])
            return engine.loglik_q(*ins)

        lnl = run()
        ctx.engine, ctx.group, ctx.run = engine, group, run
        ctx.meta = [(t.shape, t.dtype, t.device) for t in src]
        ctx.q_draws = q.shape[0]
        ctx.views = None
        if lnl.numel() == 1 and any(ctx.needs_input_grad[3:]):
            # One draw and a gradient is wanted: run the pre-order sweep right behind the post-order
            # sweep (d lnL / d theta for grad_lnL = 1; backward() scales it -- the gradient is
            # linear in the one incoming scalar) and all-reduce the packed vector, which starts
            # with lnL: ONE collective and ONE device -> host copy per evaluation instead of two of
            # each with a host round trip between the sweeps.
            packed = engine.grad_eigen_packed(None)
            if _world(group) > 1:
                dist.all_reduce(packed, group=group)
            if all(device.type == "cpu" for _, _, device in ctx.meta):
                packed = packed.cpu()
            ctx.views = engine.unpack(packed)
            return ctx.views["lnL"].to(bls.device, bls.dtype).clone()
        if _world(group) > 1:
            dist.all_reduce(lnl, group=group)
        ctx.serial = engine.eval_serial
        return lnl.to(bls.device, bls.dtype)

    @staticmethod
    def backward(ctx, grad):
        engine = ctx.engine
        if ctx.views is not None:   # gradient for grad_lnL = 1 came with the forward pass
            out = []
            for key, (shape, dtype, device) in zip(("branch_lengths", "site_rates", "props", "q",
                                                    "freqs"), ctx.meta):
                g = ctx.views[key]
                if key == "q" and g.shape[0] != ctx.q_draws:
                    g = g.sum(0, keepdim=True)
                scale = grad.detach().reshape(-1)[0].to(g.device, g.dtype)
                out.append((g * scale).to(device, dtype).reshape(shape))
            return (None, None, None) + tuple(out)
        if engine.eval_serial != ctx.serial:   # another forward ran on this engine: recompute
            ctx.run()
            ctx.serial = engine.eval_serial
        dev = torch.device("cuda", engine.device)
        packed = engine.grad_eigen_packed(grad.detach().to(dev, torch.float64).reshape(-1))
        if _world(ctx.group) > 1:
            dist.all_reduce(packed, group=ctx.group)
        if all(device.type == "cpu" for _, _, device in ctx.meta):
            packed = packed.cpu()   # one device -> host copy of the summed vector
        views = engine.unpack(packed)
        out = []
        for key, (shape, dtype, device) in zip(("branch_lengths", "site_rates", "props", "q",
                                                "freqs"), ctx.meta):
            g = views[key]
            if key == "q" and g.shape[0] != ctx.q_draws:
                g = g.sum(0, keepdim=True)
            out.append(g.to(device, dtype).reshape(shape))
        return (None, None, None) + tuple(out)


def sharded_engine_log_likelihood(engine, tensors, group=None, general=False):
    """Pattern-sharded lnL [D] of `engine`'s shard for `tensors` = (branch_lengths [D,B],
    site_rates, site_props, q [.,S,S], freqs) -- the arguments of `log_likelihood_eigen`
    (`general=True`: of `log_likelihood_expm`) -- summed over the ranks of `group`; after
    `.backward()` every rank holds the full gradient of every tensor."""
    return _ShardedEngineLikelihood.apply(engine, group, bool(general), *tensors)


# ---- draw sharding (BASELINE config 3: a batch of ADVI / HMC draws per step) -------------------
# Draws are independent evaluations of the same data with different parameters: rank g owns the
# draws [g*D/G, (g+1)*D/G) of every per-draw tensor (leading extent D) and the whole of every
# shared tensor (leading extent 1).  The exchanges are, per step, one all-reduce of the
# zero-padded lnL vector (D doubles) and one of the packed gradients (SURVEY 8(e) "Draw sharding").

class _ScatterDraws(torch.autograd.Function):
    """Forward: this rank's slice of the per-draw tensors, shared tensors untouched.
    Backward: the full gradients on every rank (slices placed into zeros, shared ones as they
    are, then one sum all-reduce of everything)."""

    @staticmethod
    def forward(ctx, group, lo, hi, draws, *tensors):
        ctx.group, ctx.lo, ctx.hi = group, lo, hi
        ctx.shapes = [tuple(t.shape) for t in tensors]
        ctx.per_draw = [t.shape[0] == draws and draws > 1 for t in tensors]
        return tuple(t[lo:hi] if p else t.view_as(t) for t, p in zip(tensors, ctx.per_draw))

    @staticmethod
    def backward(ctx, *grads):
        full = []
        for g, shape, per_draw in zip(grads, ctx.shapes, ctx.per_draw):
            if g is None:
                full.append(None)
            elif per_draw:
                z = g.new_zeros(shape)
                z[ctx.lo:ctx.hi] = g
                full.append(z)
            else:
                full.append(g.contiguous())
        live = [g for g in full if g is not None]
        if live and _world(ctx.group) > 1:
            flat = torch.cat([g.reshape(-1) for g in live])
            all_reduce_sum(flat, ctx.group)
            off = 0
            for i, g in enumerate(full):
                if g is not None:
                    full[i] = flat[off:off + g.numel()].view_as(g)
                    off += g.numel()
        return (None, None, None, None) + tuple(full)


class _GatherDraws(torch.autograd.Function):
    """Forward: lnL of all draws on every rank.  Backward: this rank's slice of the gradient."""

    @staticmethod
    def forward(ctx, group, lo, hi, draws, lnl):
        ctx.lo, ctx.hi = lo, hi
        out = lnl.new_zeros((draws,) + tuple(lnl.shape[1:]))
        out[lo:hi] = lnl
        return all_reduce_sum(out, group)

    @staticmethod
    def backward(ctx, grad):
        return None, None, None, None, grad[ctx.lo:ctx.hi]


def draw_sharded_log_likelihood(local_fn, tensors, draws, group=None):
    """lnL [D] of a batch of draws evaluated `D / world_size` draws per rank.

    `tensors`: every tensor has leading extent D (per-draw) or 1 (shared by all draws);
    `local_fn(*local_tensors) -> lnL[d_local]` evaluates this rank's draws with an engine that
    holds the whole alignment (`max_draws >= ceil(D / world_size)`).  Every rank receives lnL of
    all draws and, after `.backward()`, the full gradient of every tensor.
    """
    world = _world(group)
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(draws, rank, world)
    local = _ScatterDraws.apply(group, lo, hi, draws, *tensors)
    if hi > lo:
        lnl_local = local_fn(*local)
    else:  # more ranks than draws: this rank contributes nothing but still joins the collectives
        lnl_local = sum((t.sum() * 0.0 for t in local), torch.zeros(0, dtype=tensors[0].dtype))
    return _GatherDraws.apply(group, lo, hi, draws, lnl_local)
