"""Differentiable entry points: the custom autograd Functions live in the torch C++
extension (csrc/torch_ext.cpp, `torch::autograd::Function` over the C ABI); this module
only normalises shapes and hands tensors over.

The reference differentiates the peeling loop with the autograd tape
(SURVEY 3.4); here `backward` calls the engine's analytic pre-order pass, so no
tape over the tree is ever recorded.  Inputs are ordinary tensors (CPU tensors
in the stock torchtree set-up, where Parameters live on the host) and the
returned gradients are ordinary tensors of the same shapes.
"""
from __future__ import annotations

import torch

from ._lib import EngineError
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


def _ext():
    """The torch C++ extension (csrc/torch_ext.cpp -> _ttb200_torch.so) that holds the
    autograd Functions; it must have been built (`python -m torchtree_b200.build`)."""
    try:
        from . import _ttb200_torch
    except ImportError as exc:  # no fallback: say how to get the product path
        raise EngineError(
            "torchtree_b200: the torch extension _ttb200_torch.so is missing or cannot be "
            "loaded (%s). Build it with `python -m torchtree_b200.build`. There is no "
            "Python or CPU fallback." % exc) from exc
    return _ttb200_torch


def _handle(engine: Engine):
    """The extension's owner object of the engine (`EngineRef`, csrc/torch_ext.cpp): every
    autograd node made from a forward keeps a reference to it, so the engine outlives the
    Python `Engine` for as long as a backward may still need its buffers."""
    h = getattr(engine, "_h", None)
    if h is None or not h.value:
        raise EngineError("engine is closed")
    return engine.share(_ext())


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
    return _ext().log_likelihood_eigen(
        _handle(engine), _lead(branch_lengths, 1), _lead(site_rates, 1), _lead(site_props, 1),
        q_norm, freqs)


def log_likelihood_expm(engine: Engine, branch_lengths, site_rates, site_props, q, freqs):
    """lnL [D] of a GENERAL generator (not necessarily reversible), differentiable w.r.t. every
    tensor argument: P = exp(Q r t) for every branch x category x draw by scaling and squaring on
    the device, gradient through the exact Frechet adjoint (csrc/expm.cu) -- the device version of
    NonSymmetricSubstitutionModel.p_t = torch.matrix_exp(Q t) (abstract.py:89-94) and its tape.

    Shapes as `log_likelihood_eigen`; `q` [1 or D,S,S] is the generator as the model normalises it;
    d/d freqs is the root term only."""
    return _ext().log_likelihood_expm(
        _handle(engine), _lead(branch_lengths, 1), _lead(site_rates, 1), _lead(site_props, 1),
        _lead(q, 2), _lead(freqs, 1))


def log_likelihood_mats(engine: Engine, mats, freqs, site_props):
    """lnL [D] from transition matrices [D,B,K,S,S] computed by the caller
    (any SubstitutionModel.p_t); differentiable w.r.t. mats, freqs, site_props."""
    return _ext().log_likelihood_mats(
        _handle(engine), _lead(mats, 4), _lead(freqs, 1), _lead(site_props, 1))
