"""Differentiable entry points: custom autograd Functions over the engine.

The reference differentiates the peeling loop with the autograd tape
(SURVEY 3.4); here `backward` calls the engine's analytic pre-order pass, so no
tape over the tree is ever recorded.  Inputs are ordinary tensors (CPU tensors
in the stock torchtree set-up, where Parameters live on the host) and the
returned gradients are ordinary tensors of the same shapes.
"""
from __future__ import annotations

import torch

from .engine import Engine


def reversible_eigensystem(q_norm: torch.Tensor, freqs: torch.Tensor):
    """Eigen-system of a reversible generator through the sqrt(pi)
    symmetrisation, as SymmetricSubstitutionModel.p_t does
    (torchtree/evolution/substitution_model/abstract.py:57-66), without
    recording a graph: (V, V^-1, lambda) with P(t) = V exp(lambda t) V^-1.

    q_norm [...,S,S]; freqs [...,S] (broadcast against q_norm's batch dims)."""
    root = freqs.sqrt()
    sym = root[..., :, None] * q_norm / root[..., None, :]
    evals, u = torch.linalg.eigh(sym)
    evec = u / root[..., :, None]
    ivec = u.transpose(-1, -2) * root[..., None, :]
    return evec, ivec, evals


class _EigenLikelihood(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: Engine, bls, rates, props, q_norm, freqs):
        with torch.no_grad():
            evec, ivec, evals = reversible_eigensystem(q_norm, freqs)
            lnl = engine.loglik_eigen(bls, rates, props, evec, ivec, evals, freqs)
        ctx.engine = engine
        ctx.stamp = engine._stamp = getattr(engine, "_stamp", 0) + 1
        ctx.save_for_backward(bls, rates, props, evec, ivec, evals, freqs)
        return lnl

    @staticmethod
    def backward(ctx, grad_lnl):
        engine = ctx.engine
        bls, rates, props, evec, ivec, evals, freqs = ctx.saved_tensors
        if engine._stamp != ctx.stamp:
            # another forward ran on this engine since ours: its buffers were
            # overwritten, recompute (SURVEY 8(b) autograd contract)
            engine.loglik_eigen(bls, rates, props, evec, ivec, evals, freqs)
            engine._stamp += 1
            ctx.stamp = engine._stamp
        g = engine.grad_eigen(grad_lnl.contiguous())
        return None, g["branch_lengths"], g["site_rates"], g["props"], g["q"], g["freqs"]


class _MatsLikelihood(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: Engine, mats, freqs, props):
        with torch.no_grad():
            lnl = engine.loglik_mats(mats, freqs, props)
        ctx.engine = engine
        ctx.stamp = engine._stamp = getattr(engine, "_stamp", 0) + 1
        ctx.save_for_backward(mats, freqs, props)
        return lnl

    @staticmethod
    def backward(ctx, grad_lnl):
        engine = ctx.engine
        mats, freqs, props = ctx.saved_tensors
        if engine._stamp != ctx.stamp:
            engine.loglik_mats(mats, freqs, props)
            engine._stamp += 1
            ctx.stamp = engine._stamp
        d_mats, d_freqs, d_props = engine.grad_mats(
            grad_lnl.contiguous(), want_mats=ctx.needs_input_grad[1])
        return None, d_mats, d_freqs, d_props


def _lead(x: torch.Tensor, tail: int) -> torch.Tensor:
    """Give `x` exactly one leading draws dimension."""
    if x.dim() == tail:
        return x.unsqueeze(0)
    if x.dim() == tail + 1:
        return x
    return x.reshape((-1,) + tuple(x.shape[x.dim() - tail:]))


def log_likelihood_eigen(engine: Engine, branch_lengths, site_rates, site_props, q_norm, freqs):
    """lnL [D] of a reversible model; differentiable w.r.t. every tensor argument.

    branch_lengths [D,B] (x clock rate, zero-padded for unrooted trees),
    site_rates / site_props [1 or D,K], q_norm [1 or D,S,S] (normalised
    generator), freqs [1 or D,S].  d/d q_norm treats all S*S entries as
    independent; d/d freqs is the root term only -- chain the rest through the
    caller's Q builder (which is where the reference's graph does it too).
    """
    q_norm, freqs = _lead(q_norm, 2), _lead(freqs, 1)
    if freqs.shape[0] > q_norm.shape[0]:
        # one eigen-system per frequency draw: give Q the same leading dimension
        q_norm = q_norm.expand(freqs.shape[0], -1, -1)
    return _EigenLikelihood.apply(
        engine, _lead(branch_lengths, 1), _lead(site_rates, 1), _lead(site_props, 1),
        q_norm, freqs)


def log_likelihood_mats(engine: Engine, mats, freqs, site_props):
    """lnL [D] from transition matrices [D,B,K,S,S] computed by the caller
    (any SubstitutionModel.p_t); differentiable w.r.t. mats, freqs, site_props."""
    return _MatsLikelihood.apply(engine, _lead(mats, 4), _lead(freqs, 1), _lead(site_props, 1))
