"""Device constant-population coalescent.

`ConstantCoalescentModel` is the reference class (torchtree/evolution/coalescent.py:62-86) with
`_call` replaced: instead of the argsort / gather / cumsum graph of `ConstantCoalescent.log_prob`
(:112-134) and its autograd tape, one kernel per call sorts every draw's node heights, counts
lineages and returns the log-density together with its closed-form partial derivatives
(`ttb2_coalescent_constant`, csrc/coalescent.cu; autograd Function in csrc/torch_ext.cpp).
`install(coalescent=True)` (tree_likelihood.py) / `python -m torchtree_b200.cli --b200-coalescent`
make existing configs resolve to it.  This module needs torchtree importable;
`constant_coalescent_log_prob` does not.
"""
from __future__ import annotations

import torch


def constant_coalescent_log_prob(node_heights: torch.Tensor, theta: torch.Tensor,
                                 device: int = 0) -> torch.Tensor:
    """log p(node_heights | theta), differentiable w.r.t. both, with the reference's shapes:
    node_heights [..., 2T-1] (tips first), theta [..., 1] -> [..., 1]."""
    from .function import _ext

    batch = torch.broadcast_shapes(node_heights.shape[:-1], theta.shape[:-1])
    n = node_heights.shape[-1]
    h = node_heights.expand(batch + (n,)).reshape(-1, n)
    th = theta.reshape(-1) if theta.numel() == 1 else theta.expand(batch + (1,)).reshape(-1)
    return _ext().constant_coalescent(int(device), h, th).reshape(batch + (1,))


def _model_class():
    from torchtree.evolution.coalescent import ConstantCoalescentModel as _Reference

    class ConstantCoalescentModel(_Reference):
        """Drop-in for torchtree's ConstantCoalescentModel (same constructor and JSON)."""

        device_index = 0

        def _call(self, *args, **kwargs) -> torch.Tensor:
            return constant_coalescent_log_prob(self.tree_model.node_heights, self.theta.tensor,
                                                self.device_index)

    ConstantCoalescentModel.__module__ = __name__
    ConstantCoalescentModel.__qualname__ = "ConstantCoalescentModel"
    return ConstantCoalescentModel


def __getattr__(name):
    # the class is built on first use: importing this module must not require torchtree
    if name == "ConstantCoalescentModel":
        cls = _model_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
