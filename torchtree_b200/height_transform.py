"""Device node-height transform.

`GeneralNodeHeightTransform` here has the constructor and methods of the reference
class (torchtree/evolution/tree_height_transform.py:7-98) -- `_call` (ratios + root
height -> internal node heights) runs on the GPU through `ttb2_heights_forward` /
`ttb2_heights_backward` instead of a Python loop of T-2 taped tensor updates;
`_inverse` and `log_abs_det_jacobian` are single vectorised torch expressions, as in
the reference.  `install()` (tree_likelihood.py) rebinds the name torchtree's
`ReparameterizedTimeTreeModel` looks up.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch.distributions import Transform

from . import _lib
from ._lib import EngineError


class NodeHeightPlan:
    """One topology + its lower bounds on the device (`ttb2_heights_create`)."""

    def __init__(self, tip_count: int, postorder, bounds, device: int = 0):
        self._lib = _lib.load()
        self.T = int(tip_count)
        self.I = self.T - 1
        post = np.ascontiguousarray(np.asarray(postorder, dtype=np.int32).reshape(-1, 3))
        if post.shape[0] != self.I:
            raise EngineError("postorder must have tip_count - 1 rows")
        b = np.ascontiguousarray(np.asarray(bounds, dtype=np.float64).reshape(-1))
        if b.shape[0] != self.I:
            raise EngineError("bounds must have tip_count - 1 entries")
        self.device = int(device)
        h = ctypes.c_void_p()
        _lib.check(self._lib.ttb2_heights_create(
            self.T, post.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
            self.device, ctypes.byref(h)), "ttb2_heights_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ttb2_heights_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _where(self, *tensors):
        on_gpu = [t.is_cuda for t in tensors]
        if all(on_gpu):
            torch.cuda.current_stream(tensors[0].device).synchronize()
            return 1
        if any(on_gpu):
            raise EngineError("all tensors of one call must be on the host or on the device")
        return 0

    @staticmethod
    def _flat(t, I):
        t = t.detach()
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        return t.reshape(-1, I).contiguous()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xf = self._flat(x, self.I)
        out = torch.empty_like(xf)
        _lib.check(self._lib.ttb2_heights_forward(
            self._h, xf.shape[0], ctypes.c_void_p(xf.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            self._where(xf, out)), "ttb2_heights_forward")
        return out.reshape(x.shape)

    def backward(self, x, heights, grad_heights) -> torch.Tensor:
        xf, hf, gf = (self._flat(t, self.I) for t in (x, heights, grad_heights))
        out = torch.empty_like(xf)
        _lib.check(self._lib.ttb2_heights_backward(
            self._h, xf.shape[0], ctypes.c_void_p(xf.data_ptr()), ctypes.c_void_p(hf.data_ptr()),
            ctypes.c_void_p(gf.data_ptr()), ctypes.c_void_p(out.data_ptr()),
            self._where(xf, hf, gf, out)), "ttb2_heights_backward")
        return out.reshape(x.shape)


def node_heights(x: torch.Tensor, plan: NodeHeightPlan) -> torch.Tensor:
    """Differentiable ratios/root-height -> internal node heights: the autograd Function
    is `NodeHeights` in the torch C++ extension (csrc/torch_ext.cpp)."""
    from .function import _ext

    if not getattr(plan, "_h", None):
        raise EngineError("node-height plan is closed")
    if x.shape[-1] != plan.I:
        raise EngineError("node_heights: last dimension must be tip_count - 1")
    return _ext().node_heights(int(plan._h.value), plan.device, x)


class GeneralNodeHeightTransform(Transform):
    """Drop-in for torchtree's class of the same name (tree_height_transform.py:7-98)."""
    bijective = True
    sign = +1

    def __init__(self, tree, cache_size=0, device: int = 0) -> None:
        super().__init__(cache_size=cache_size)
        self.tree = tree
        self.taxa_count = tree.taxa_count
        self.device_index = device
        self._plan = None
        self.update_bounds()

    # the reference exposes both; a topology change needs both
    def sort_indices(self):
        self.update_bounds()

    def update_bounds(self) -> None:
        T = self.taxa_count
        postorder = [tuple(int(v) for v in t) for t in self.tree.postorder]
        times = self.tree.sampling_times
        low = [float(t) for t in times] + [0.0] * (T - 1)
        parent = [-1] * (2 * T - 1)
        for node, left, right in postorder:
            low[node] = max(low[left], low[right])
            parent[left] = node
            parent[right] = node
        self._bounds = torch.cat((times, torch.tensor(low[T:], dtype=times.dtype)), -1)
        root = postorder[-1][0]
        self._root = root - T
        inner = [n for n in range(T, 2 * T - 1) if n != root]
        self._child_idx = torch.tensor([n - T for n in inner], dtype=torch.long)
        self._parent_idx = torch.tensor([parent[n] - T for n in inner], dtype=torch.long)
        self._postorder = postorder
        if self._plan is not None:
            self._plan.close()
        self._plan = None

    def _get_plan(self) -> NodeHeightPlan:
        if self._plan is None:
            bounds = self._bounds[self.taxa_count:].detach().cpu().numpy()
            self._plan = NodeHeightPlan(self.taxa_count, self._postorder, bounds,
                                        self.device_index)
        return self._plan

    def _call(self, x: torch.Tensor) -> torch.Tensor:
        return node_heights(x, self._get_plan())

    def _inverse(self, y: torch.Tensor) -> torch.Tensor:
        b = self._bounds[self.taxa_count:][self._child_idx]
        x = y.clone()
        x[..., self._child_idx] = (y[..., self._child_idx] - b) / (y[..., self._parent_idx] - b)
        return x

    def log_abs_det_jacobian(self, x, y):
        b = self._bounds[self.taxa_count:][self._child_idx]
        return torch.log(y[..., self._parent_idx] - b).sum(-1)
