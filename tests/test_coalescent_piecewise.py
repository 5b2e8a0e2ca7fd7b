"""Piecewise-constant coalescents on the device ("piecewise variants"):
skyride (PiecewiseConstantCoalescent, coalescent.py:311-396) and skygrid
(PiecewiseConstantCoalescentGrid, :459-549).  The oracle against golden vectors from the real
reference distributions (CPU), the native kernel against both (GPU), and the drop-in model
classes against the reference classes."""
import glob
import os

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "coalescent_piecewise", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _compare(rec, lp, d_heights, d_theta):
    T = (rec["heights"].shape[-1] + 1) // 2
    np.testing.assert_allclose(lp, rec["log_prob"], rtol=1e-12)
    tied = len(np.unique(rec["heights"][0, :T])) < T   # tied tip times: argsort order is arbitrary
    got, want = (d_heights[..., T:], rec["d_heights"][..., T:]) if tied \
        else (d_heights, rec["d_heights"])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    np.testing.assert_allclose(d_theta, rec["d_theta"], rtol=1e-9,
                               atol=1e-9 * np.abs(rec["d_theta"]).max())


def _oracle(rec, h, theta):
    from oracle.coalescent import piecewise_grid_log_prob, piecewise_log_prob

    if rec["grid"].size:
        return piecewise_grid_log_prob(h, theta, torch.tensor(rec["grid"]))
    return piecewise_log_prob(h, theta)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_matches_reference_golden(path):
    rec = np.load(path)
    h = torch.tensor(rec["heights"], requires_grad=True)
    theta = torch.tensor(rec["theta"], requires_grad=True)
    lp = _oracle(rec, h, theta)
    (lp * torch.tensor(rec["grad_out"])).sum().backward()
    _compare(rec, lp.detach().numpy(), h.grad.numpy(), theta.grad.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("where", ["host", "device"])
def test_native_matches_reference_golden(path, where):
    from torchtree_b200.coalescent import piecewise_coalescent_log_prob

    rec = np.load(path)
    dev = "cuda" if where == "device" else "cpu"
    h = torch.tensor(rec["heights"], device=dev, requires_grad=True)
    theta = torch.tensor(rec["theta"], device=dev, requires_grad=True)
    grid = torch.tensor(rec["grid"], device=dev) if rec["grid"].size else None
    lp = piecewise_coalescent_log_prob(h, theta, grid)
    assert lp.shape == rec["log_prob"].shape and lp.device.type == dev
    (lp * torch.tensor(rec["grad_out"], device=dev)).sum().backward()
    _compare(rec, lp.detach().cpu().numpy(), h.grad.cpu().numpy(), theta.grad.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("T,D,G,shared", [(500, 16, 0, False), (500, 16, 100, True), (2, 2, 0, True),
                                          (3, 1, 4, False), (5000, 2, 0, True), (5000, 2, 200, False)])
def test_native_matches_oracle(T, D, G, shared):
    from oracle.coalescent import piecewise_grid_log_prob, piecewise_log_prob
    from torchtree_b200.coalescent import piecewise_coalescent_log_prob

    rng = np.random.default_rng(T + D + G)
    tips = rng.uniform(0.0, 10.0, (D, T))
    inner = tips.max(-1, keepdims=True) + np.cumsum(rng.exponential(0.3, (D, T - 1)), -1)
    hv = np.concatenate([tips, inner], -1)
    M = G + 1 if G else T - 1
    tv = rng.uniform(1.0, 20.0, (1 if shared else D, M))
    grid = torch.tensor(np.linspace(0.0, 0.7 * hv.max(), G + 1)[1:]) if G else None
    outs = []
    for native in (False, True):
        h = torch.tensor(hv, requires_grad=True)
        theta = torch.tensor(tv, requires_grad=True)
        if native:
            lp = piecewise_coalescent_log_prob(h, theta, grid)
        else:
            lp = piecewise_grid_log_prob(h, theta, grid) if G else piecewise_log_prob(h, theta)
        lp.sum().backward()
        outs.append((lp.detach().numpy(), h.grad.numpy(), theta.grad.numpy()))
    (lo, gho, gto), (ln, ghn, gtn) = outs
    np.testing.assert_allclose(ln, lo, rtol=1e-12)
    np.testing.assert_allclose(ghn, gho, rtol=1e-9, atol=1e-9 * np.abs(gho).max())
    np.testing.assert_allclose(gtn, gto, rtol=1e-9, atol=1e-9 * np.abs(gto).max())
    again = piecewise_coalescent_log_prob(torch.tensor(hv), torch.tensor(tv), grid)
    assert np.array_equal(again.numpy(), ln)   # bit-wise reproducible
    nan = torch.tensor(hv).clone()
    nan[0, -1] = float("nan")
    assert torch.isnan(piecewise_coalescent_log_prob(nan, torch.tensor(tv), grid)[0]).all()
