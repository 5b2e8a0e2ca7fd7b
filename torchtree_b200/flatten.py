"""Host-side glue between torchtree's model objects and the engine.

`evaluate_models` is the body of the reference's `TreeLikelihoodModel._call`
(torchtree/evolution/tree_likelihood.py:313-356) up to the point where the
reference calls `subst_model.p_t` and the peeling functions; from there the
engine takes over.  It is duck-typed (no torchtree import) so that it can be
exercised on a GPU box where torchtree is not installed, with stand-ins that
expose the same attributes as the reference classes:

  tree_model.branch_lengths()            tree_model.py:250, :407-424
  site_model.rates() / .probabilities()  site_model.py:197-207
  clock_model.rates                      branch_model.py:43-47
  subst_model.q() / .frequencies / .norm(Q) / .p_t(t)
                                         substitution_model/abstract.py:46-94
"""
from __future__ import annotations

import torch

from .function import log_likelihood_eigen, log_likelihood_expm, log_likelihood_mats


def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def substitution_route(subst_model) -> str:
    """Which engine entry point serves this substitution model.

    "eigen": reversible models whose P(t) the reference obtains from the
             symmetrised eigen-decomposition (SymmetricSubstitutionModel.p_t,
             abstract.py:57-76: HKY, GTR, MG94, GeneralSymmetric...), the
             closed-form JC69 family (nucleotide.py:102-113, general.py:51-69)
             and the empirical amino-acid models (general.py:295-331);
    "expm":  NonSymmetricSubstitutionModel (p_t = matrix_exp(Q t), abstract.py:89-94;
             GeneralNonSymmetricSubstitutionModel, the discrete-trait models): the
             matrix exponential and its adjoint run on the device (csrc/expm.cu);
    "mats":  anything else (a user-defined model): the model's own p_t()
             supplies the matrices and autograd carries the gradient from
             d lnL / d P back into its parameters.
    """
    names = _mro_names(subst_model)
    if "NonSymmetricSubstitutionModel" in names:
        return "expm"
    if names & {"SymmetricSubstitutionModel", "JC69", "GeneralJC69",
                "EmpiricalSubstitutionModel"}:
        return "eigen"
    return "mats"


def normalised_generator(subst_model):
    """Q / (-sum_i pi_i Q_ii) as the reference normalises it
    (abstract.py:49-50, :58-59; general.py:303-305); differentiable w.r.t.
    the model's parameters."""
    names = _mro_names(subst_model)
    freqs = subst_model.frequencies
    if "GeneralJC69" in names:
        S = subst_model.state_count
        q = torch.full((S, S), 1.0 / (S - 1), dtype=freqs.dtype, device=freqs.device)
        q = q - torch.diag_embed(q.sum(-1))
        return q, freqs
    q = subst_model.q()
    if "JC69" in names:
        return q, freqs  # already normalised (nucleotide.py:115-126)
    norm = -(torch.diagonal(q, dim1=-2, dim2=-1) * freqs).sum(-1)
    return q / norm.unsqueeze(-1).unsqueeze(-1), freqs


def scaled_branch_lengths(tree_model, clock_model, sample_shape):
    """bls [..., 2T-2]: zero-padded (unrooted, no clock) or rate * length
    (tree_likelihood.py:323-344)."""
    branch_lengths = tree_model.branch_lengths()
    if clock_model is None:
        if branch_lengths.dim() == 1:
            branch_lengths = branch_lengths.expand(sample_shape + (-1,))
        pad = torch.zeros(sample_shape + (1,), dtype=branch_lengths.dtype,
                          device=branch_lengths.device)
        return torch.cat((branch_lengths, pad), -1)
    if branch_lengths.dim() == 1:
        return clock_model.rates * branch_lengths.expand(sample_shape + (1, -1))
    return clock_model.rates * branch_lengths


def _flat(x, sample_shape, tail):
    """`x` = [batch..., *tail] -> [1 or D, *tail].  A tensor without batch dimensions (or with
    extents of 1 only) stays shared by all draws; any other batch shape -- also a partial one
    such as [m, K] or [n, 1, K] under sample_shape [n, m] -- is broadcast to `sample_shape`
    first, as the reference's tensor arithmetic would (tree_likelihood.py:313-356)."""
    tail = tuple(tail)
    batch = tuple(x.shape[:x.dim() - len(tail)])
    if all(int(b) == 1 for b in batch):
        return x.reshape((1,) + tail)
    return x.expand(tuple(sample_shape) + tail).reshape((-1,) + tail)


def _nccl(group) -> bool:
    import torch.distributed as dist

    return dist.is_available() and dist.is_initialized() and "nccl" in dist.get_backend(group)


def evaluate_models(engine, tree_model, site_model, subst_model, clock_model, sample_shape,
                    shard=None):
    """lnL with the reference's output contract: shape sample_shape + (1,)
    (SURVEY F9), differentiable w.r.t. every parameter the sub-models carry.

    `shard` = None, or ("patterns" | "draws", process group): this process holds one shard of
    a multi-GPU evaluation (torchtree_b200.sharded); `engine` is then the engine of this rank's
    patterns (or of all patterns, for draw sharding), or None when the rank owns nothing."""
    sample_shape = torch.Size(sample_shape)
    D = 1
    for n in sample_shape:
        D *= int(n)
    bls = scaled_branch_lengths(tree_model, clock_model, sample_shape)
    B = bls.shape[-1]
    bls = _flat(bls, sample_shape, (B,))
    if bls.shape[0] != D:
        bls = bls.expand(D, B)
    rates = site_model.rates()
    K = rates.shape[-1]
    rates = _flat(rates, sample_shape, (K,))
    props = _flat(site_model.probabilities(), sample_shape, (K,))
    route = substitution_route(subst_model)
    if route in ("eigen", "expm"):
        q, freqs = normalised_generator(subst_model)
        S = freqs.shape[-1]
        tensors = (bls, rates, props, _flat(q, sample_shape, (S, S)),
                   _flat(freqs, sample_shape, (S,)))
        op = log_likelihood_eigen if route == "eigen" else log_likelihood_expm
        local = lambda *a: op(engine, *a)  # noqa: E731
    else:
        freqs = subst_model.frequencies
        S = freqs.shape[-1]
        t = bls.unsqueeze(-1) * rates.unsqueeze(-2)  # [D,B,K]
        # the model may carry its own sample shape: hand p_t the reference's layout
        mats = subst_model.p_t(t.reshape(sample_shape + (B, K)))
        mats = mats.expand(sample_shape + (B, K, S, S)).reshape(D, B, K, S, S)
        tensors = (mats, _flat(freqs, sample_shape, (S,)), props)
        local = lambda *a: log_likelihood_mats(engine, *a)  # noqa: E731
    if shard is None:
        lnl = local(*tensors)
    else:
        from .sharded import (draw_sharded_log_likelihood, sharded_engine_log_likelihood,
                              sharded_log_likelihood)

        kind, group = shard
        if kind == "draws":
            lnl = draw_sharded_log_likelihood(local, tensors, D, group)
        elif engine is not None and route in ("eigen", "expm") and _nccl(group):
            # collectives on the device, next to the engine's packed outputs
            lnl = sharded_engine_log_likelihood(engine, tensors, group, general=route == "expm")
        else:
            lnl = sharded_log_likelihood(local if engine is not None else None, tensors, group)
    return lnl.reshape(sample_shape + (1,))
