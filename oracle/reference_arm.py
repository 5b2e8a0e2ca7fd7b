"""The REAL reference (4ment/torchtree, vendored unmodified into `baseline/_ref` by
tools/vendor_reference.py) evaluating a flattened synthetic problem on the CPU --
TEST / MEASUREMENT INFRASTRUCTURE, NOT PRODUCT.

This is the CPU arm of bench.py (`--impl reference`, and `cpu_baseline` with
`kind: "reference"`): the reference's own model objects and functions on the same problem
object the CUDA engine gets, following the template of the reference's own benchmark
(benchmarks/benchmark.py:273-304) and the body of `TreeLikelihoodModel._call`
(torchtree/evolution/tree_likelihood.py:313-356):

    rates  = WeibullSiteModel.rates()                          site_model.py:173-207
    mats   = GTR.p_t(bls[...,B,1] * rates[...,1,K])            substitution_model/abstract.py:57-76
    lnL    = calculate_treelikelihood_discrete_rescaled(...)   tree_likelihood.py:186-221
    lnL.backward()                                             autograd tape

Nothing of this repository's engine, kernels or oracle restatement is on that path.  Only
bench.py and tests import this module.  `dendropy` (parsing only; not installable here) is
the stand-in under oracle/dendropy_shim -- it is needed only to get past the import of
`torchtree.evolution.tree_model`.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_VENDORED = os.path.join(REPO, "baseline", "_ref")
_SHIM = os.path.join(REPO, "oracle", "dendropy_shim")


def available() -> bool:
    return os.path.isfile(os.path.join(_VENDORED, "torchtree", "evolution", "tree_likelihood.py"))


def _activate():
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python tools/vendor_reference.py` "
                           "where /root/reference exists")
    for p in (_SHIM, _VENDORED):
        if p not in sys.path:
            sys.path.insert(0, p)


class ReferenceProblem:
    """The reference's objects for one `torchtree_b200.synthetic.Problem` (one draw)."""

    def __init__(self, prob):
        _activate()
        torch.set_default_dtype(torch.float64)
        from torchtree import Parameter
        from torchtree.evolution.site_model import ConstantSiteModel, WeibullSiteModel
        from torchtree.evolution.substitution_model.general import (
            GeneralSymmetricSubstitutionModel,
        )
        from torchtree.evolution.substitution_model.nucleotide import GTR
        from torchtree.evolution.tree_likelihood import (
            calculate_treelikelihood_discrete_rescaled,
        )

        if prob.draws != 1 or "exchangeabilities" not in prob.model_params:
            raise ValueError("reference arm: one draw of a GeneralSymmetric/GTR problem expected")
        S, K = prob.state_count, prob.category_count
        self.prob = prob
        self._peel = calculate_treelikelihood_discrete_rescaled
        self.rates6 = Parameter("rates", torch.tensor(prob.model_params["exchangeabilities"][0]))
        self.freqs = Parameter("freqs", torch.tensor(prob.freqs[0]))
        if S == 4:
            self.subst = GTR("gtr", self.rates6, self.freqs)
        else:
            from torchtree.evolution.datatype import GeneralDataType

            n_ex = S * (S - 1) // 2
            self.subst = GeneralSymmetricSubstitutionModel(
                "sym", GeneralDataType("dt", [str(i) for i in range(S)]),
                Parameter(None, torch.arange(n_ex)), self.rates6, self.freqs)
        if K > 1:
            self.shape = Parameter("shape", torch.tensor(
                [float(prob.model_params["weibull_shape"][0])]))
            self.site = WeibullSiteModel("site", self.shape, K)
        else:
            self.shape = None
            self.site = ConstantSiteModel("site")
        # unrooted parameterisation: 2T-3 free lengths, the likelihood pads the zero
        # (tree_likelihood.py:323-337)
        self.blens = Parameter("blens", torch.tensor(prob.branch_lengths[0, :-1]))
        table = np.concatenate([np.eye(S), np.ones((1, S))], 0) if prob.code_partials is None \
            else np.asarray(prob.code_partials)
        codes = prob.tip_states.astype(np.int64)
        if prob.code_partials is None:
            codes = np.minimum(codes, S)
        self.tips = [torch.from_numpy(np.ascontiguousarray(table[c].T)) for c in codes]
        self.weights = torch.tensor(prob.weights)
        self.post = [tuple(int(x) for x in row) for row in prob.postorder]
        self.leaves = [p for p in (self.blens, self.shape, self.rates6, self.freqs)
                       if p is not None]

    def evaluate(self, want_grad: bool = True) -> dict:
        """One logL (+ gradient by the reference's autograd tape) evaluation."""
        for p in self.leaves:
            p.tensor.grad = None
            p.requires_grad = want_grad  # also fires parameter_changed (parameter.py:63-66)
        K = self.prob.category_count
        rates = self.site.rates()
        rates = rates.expand((1, -1)) if rates.dim() == 1 else rates.reshape((1, -1))
        probs = self.site.probabilities().unsqueeze(-1).unsqueeze(-1)
        bl = self.blens.tensor
        bls = torch.cat((bl, torch.zeros((1,), dtype=bl.dtype)), -1)
        mats = self.subst.p_t(bls.reshape((-1, 1)) * rates)
        frequencies = self.subst.frequencies.reshape((1, -1))
        partials = list(self.tips) + [None] * (len(self.tips) - 1)
        lnl = self._peel(partials, self.weights, self.post, mats, frequencies, probs)
        out = {"lnL": lnl.detach().reshape(-1).numpy().copy()}
        if want_grad:
            lnl.sum().backward()
            g = self.blens.grad.numpy()
            out["branch_lengths"] = np.concatenate([g, [0.0]])[None, :]
            if self.shape is not None:
                out["weibull_shape"] = self.shape.grad.numpy().copy()
            out["gtr_rates"] = self.rates6.grad.numpy().copy()
            out["gtr_freqs"] = self.freqs.grad.numpy().copy()
        assert K == rates.shape[-1]
        return out


def estimated_tape_bytes(prob) -> float:
    """Host memory the reference's autograd tape holds for one logL+gradient evaluation:
    per internal node the two matmul results, their product, the scaled vector and the
    division's saved operands ([K,S,N] doubles each) -- ~6 tensors per node."""
    return 6.0 * (prob.tip_count - 1) * prob.category_count * prob.state_count * \
        prob.pattern_count * 8.0 + prob.tip_count * prob.state_count * prob.pattern_count * 8.0
