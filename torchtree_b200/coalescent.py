"""Device constant-population coalescent.

`ConstantCoalescentModel` is the reference class (torchtree/evolution/coalescent.py:62-86) with
`_call` replaced: instead of the argsort / gather / cumsum graph of `ConstantCoalescent.log_prob`
(:112-134) and its autograd tape, one kernel per call sorts every draw's node heights, counts
lineages and returns the log-density together with its closed-form partial derivatives
(`ttb2_coalescent_constant`, csrc/coalescent.cu; autograd Function in csrc/torch_ext.cpp).
`install(coalescent=True)` (tree_likelihood.py) / `python -m torchtree_b200.cli --b200-coalescent`
make existing configs resolve to it.  The piecewise-constant models -- skyride
(`PiecewiseConstantCoalescentModel`, :311-456) and skygrid (`PiecewiseConstantCoalescentGridModel`,
:459-549, :681-727) -- get the same treatment through `ttb2_coalescent_piecewise`.  This module needs torchtree importable;
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


def piecewise_coalescent_log_prob(node_heights: torch.Tensor, theta: torch.Tensor, grid=None,
                                  device: int = 0) -> torch.Tensor:
    """log p(node_heights | theta) of the piecewise-constant coalescents, differentiable w.r.t.
    heights and theta, with the reference's shapes: node_heights [..., 2T-1], theta [..., M] ->
    [..., 1].  `grid` None: skyride, M = T-1 (coalescent.py:311-396); `grid` [G] ascending: skygrid,
    M = G+1 (:459-549)."""
    from .function import _ext

    batch = torch.broadcast_shapes(node_heights.shape[:-1], theta.shape[:-1])
    n, M = node_heights.shape[-1], theta.shape[-1]
    h = node_heights.expand(batch + (n,)).reshape(-1, n)
    th = theta.reshape(1, M) if theta.numel() == M else theta.expand(batch + (M,)).reshape(-1, M)
    g = torch.empty(0, dtype=torch.float64) if grid is None else grid.reshape(-1)
    return _ext().piecewise_coalescent(int(device), h, th, g).reshape(batch + (1,))


def _piecewise_model_class():
    from torchtree.evolution.coalescent import PiecewiseConstantCoalescentModel as _Reference

    class PiecewiseConstantCoalescentModel(_Reference):
        """Drop-in for torchtree's skyride model (same constructor and JSON)."""

        device_index = 0

        def _call(self, *args, **kwargs) -> torch.Tensor:
            return piecewise_coalescent_log_prob(self.tree_model.node_heights, self.theta.tensor,
                                                 None, self.device_index)

    PiecewiseConstantCoalescentModel.__module__ = __name__
    PiecewiseConstantCoalescentModel.__qualname__ = "PiecewiseConstantCoalescentModel"
    return PiecewiseConstantCoalescentModel


def _grid_model_class():
    from torchtree.evolution.coalescent import PiecewiseConstantCoalescentGridModel as _Reference

    class PiecewiseConstantCoalescentGridModel(_Reference):
        """Drop-in for torchtree's skygrid model; the soft (temperature) variant stays with the
        reference class."""

        device_index = 0

        def _call(self, *args, **kwargs) -> torch.Tensor:
            if self.temperature is not None:
                return super()._call(*args, **kwargs)
            return piecewise_coalescent_log_prob(self.tree_model.node_heights, self.theta.tensor,
                                                 self.grid.tensor, self.device_index)

    PiecewiseConstantCoalescentGridModel.__module__ = __name__
    PiecewiseConstantCoalescentGridModel.__qualname__ = "PiecewiseConstantCoalescentGridModel"
    return PiecewiseConstantCoalescentGridModel


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
    builders = {"ConstantCoalescentModel": _model_class,
                "PiecewiseConstantCoalescentModel": _piecewise_model_class,
                "PiecewiseConstantCoalescentGridModel": _grid_model_class}
    if name in builders:
        cls = builders[name]()
        globals()[name] = cls
        return cls
    raise AttributeError(name)
