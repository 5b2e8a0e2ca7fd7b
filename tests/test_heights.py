"""Node-height transform: oracle vs golden vectors of the real
reference class (CPU), native CUDA path vs both (GPU, through the C ABI)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import heights as oh

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "heights", "*.npz")))
RTOL = 1e-12   # fp64 fma chains of depth <= tree height


def _load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-300)
    assert np.max(np.abs(a - b) / np.maximum(scale, np.max(np.abs(b)) * 1e-6)) < rtol


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z = _load(path)
    T = int(z["T"])
    bounds = oh.internal_bounds(T, z["postorder"], z["sampling_times"])
    assert np.array_equal(bounds.numpy(), z["bounds"])
    x = torch.tensor(z["x"], requires_grad=True)
    h = oh.heights_from_ratios(T, z["postorder"], bounds, x)
    _close(h.detach().numpy(), z["heights"])
    ld = oh.log_abs_det_jacobian(T, z["postorder"], bounds, h)
    _close(ld.detach().numpy(), z["logdet"])
    (gx,) = torch.autograd.grad((h * torch.tensor(z["g"])).sum(), x, retain_graph=True)
    _close(gx.numpy(), z["grad_x"], 1e-10)
    (gl,) = torch.autograd.grad(ld.sum(), x)
    _close(gl.numpy(), z["grad_x_logdet"], 1e-10)
    inv = oh.ratios_from_heights(T, z["postorder"], bounds, torch.tensor(z["heights"]))
    assert np.allclose(inv.numpy(), z["inverse"], rtol=0, atol=1e-9)


class _Tree:
    def __init__(self, T, postorder, sampling_times):
        self.taxa_count = T
        self.postorder = [tuple(int(v) for v in t) for t in postorder]
        self.sampling_times = torch.as_tensor(sampling_times)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("where", ["host", "device"])
def test_native_matches_reference_golden(path, where):
    from torchtree_b200.height_transform import GeneralNodeHeightTransform
    z = _load(path)
    T = int(z["T"])
    dev = "cuda:0" if where == "device" else "cpu"
    tr = GeneralNodeHeightTransform(_Tree(T, z["postorder"], z["sampling_times"]))
    assert np.array_equal(tr._bounds[T:].numpy(), z["bounds"])
    x = torch.tensor(z["x"], device=dev, requires_grad=True)
    if where == "device":
        tr._bounds = tr._bounds.to(dev)
        tr._child_idx, tr._parent_idx = tr._child_idx.to(dev), tr._parent_idx.to(dev)
    h = tr(x)
    _close(h.detach().cpu().numpy(), z["heights"])
    ld = tr.log_abs_det_jacobian(x, h)
    _close(ld.detach().cpu().numpy(), z["logdet"])
    (gx,) = torch.autograd.grad((h * torch.tensor(z["g"], device=dev)).sum(), x, retain_graph=True)
    _close(gx.cpu().numpy(), z["grad_x"], 1e-10)
    (gl,) = torch.autograd.grad(ld.sum(), x)       # through y: exercises the native backward
    _close(gl.cpu().numpy(), z["grad_x_logdet"], 1e-10)
    inv = tr.inv(h.detach())
    assert np.allclose(inv.cpu().numpy(), z["inverse"], rtol=0, atol=1e-9)
    # unbatched input, as the reference's TimeTreeModel passes it
    h1 = tr(torch.tensor(z["x"][0], device=dev))
    _close(h1.cpu().numpy(), z["heights"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("topology,T,D", [("random", 1000, 128), ("caterpillar", 300, 5),
                                          ("balanced", 256, 7), ("random", 2, 3)])
def test_native_matches_oracle(topology, T, D):
    from torchtree_b200.height_transform import NodeHeightPlan, node_heights
    from torchtree_b200.synthetic import random_postorder
    rng = np.random.default_rng(T + D)
    post = random_postorder(T, rng, topology)
    times = rng.uniform(0, 5, T) * (rng.random(T) < 0.5)
    bounds = oh.internal_bounds(T, post, times)
    x = torch.rand(D, T - 1, dtype=torch.float64) * 0.9 + 0.05
    x[:, int(post[-1][0]) - T] = 6.0 + torch.rand(D, dtype=torch.float64)
    g = torch.randn(D, T - 1, dtype=torch.float64)
    xo = x.clone().requires_grad_(True)
    ho = oh.heights_from_ratios(T, post, bounds, xo)
    (go,) = torch.autograd.grad((ho * g).sum(), xo)
    plan = NodeHeightPlan(T, post, bounds.numpy())
    xn = x.clone().requires_grad_(True)
    hn = node_heights(xn, plan)
    (gn,) = torch.autograd.grad((hn * g).sum(), xn)
    _close(hn.detach().numpy(), ho.detach().numpy())
    _close(gn.numpy(), go.numpy(), 1e-10)
    # bit-wise reproducible
    assert torch.equal(node_heights(x, plan), hn.detach())
    plan.close()


@pytest.mark.gpu
def test_invalid_topology_is_rejected():
    from torchtree_b200._lib import EngineError
    from torchtree_b200.height_transform import NodeHeightPlan
    with pytest.raises(EngineError):
        NodeHeightPlan(3, [(4, 3, 2), (3, 0, 1)], [0.0, 0.0])    # child used before it is defined
    with pytest.raises(EngineError):
        NodeHeightPlan(3, [(3, 0, 1)], [0.0, 0.0])               # wrong number of rows
