"""Constant-population coalescent on the device: the oracle against golden
vectors from the real reference distribution (CPU), the native kernel against both (GPU)."""
import glob
import os

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(REPO, "tests", "golden", "coalescent", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _internal(a, T):
    return a[..., T:]


def _compare(rec, lp, d_heights, d_theta):
    """Gradients w.r.t. tip times that are tied (isochronous sampling) depend on the arbitrary
    order argsort gives the ties; tips are data, so only internal nodes are compared there."""
    T = (rec["heights"].shape[-1] + 1) // 2
    np.testing.assert_allclose(lp, rec["log_prob"], rtol=1e-12)
    tied = len(np.unique(rec["heights"][0, :T])) < T
    got, want = (_internal(d_heights, T), _internal(rec["d_heights"], T)) if tied \
        else (d_heights, rec["d_heights"])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    np.testing.assert_allclose(d_theta, rec["d_theta"], rtol=1e-10)


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_matches_reference_golden(path):
    from oracle.coalescent import constant_log_prob

    rec = np.load(path)
    h = torch.tensor(rec["heights"], requires_grad=True)
    theta = torch.tensor(rec["theta"], requires_grad=True)
    lp = constant_log_prob(h, theta)
    (lp * torch.tensor(rec["grad_out"])).sum().backward()
    _compare(rec, lp.detach().numpy(), h.grad.numpy(), theta.grad.numpy())


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from torchtree_b200.coalescent import constant_coalescent_log_prob

    rec = np.load(GOLDEN[0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        constant_coalescent_log_prob(torch.tensor(rec["heights"]), torch.tensor(rec["theta"]))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
@pytest.mark.parametrize("where", ["host", "device"])
def test_native_matches_reference_golden(path, where):
    from torchtree_b200.coalescent import constant_coalescent_log_prob

    rec = np.load(path)
    dev = "cuda" if where == "device" else "cpu"
    h = torch.tensor(rec["heights"], device=dev, requires_grad=True)
    theta = torch.tensor(rec["theta"], device=dev, requires_grad=True)
    lp = constant_coalescent_log_prob(h, theta)
    assert lp.shape == rec["log_prob"].shape and lp.device.type == dev
    (lp * torch.tensor(rec["grad_out"], device=dev)).sum().backward()
    _compare(rec, lp.detach().cpu().numpy(), h.grad.cpu().numpy(), theta.grad.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("T,D,shared", [(500, 128, True), (2, 3, False), (3, 1, True),
                                        (1000, 16, False), (4096, 2, True),
                                        (4097, 2, False), (10000, 3, True)])
def test_native_matches_oracle(T, D, shared):
    from oracle.coalescent import constant_log_prob
    from torchtree_b200.coalescent import constant_coalescent_log_prob

    rng = np.random.default_rng(T + D)
    tips = rng.uniform(0.0, 10.0, (D, T))
    inner = tips.max(-1, keepdims=True) + np.cumsum(rng.exponential(0.3, (D, T - 1)), -1)
    hv = np.concatenate([tips, inner], -1)
    tv = rng.uniform(1.0, 20.0, (1 if shared else D, 1))
    outs = []
    for fn in (constant_log_prob, constant_coalescent_log_prob):
        h = torch.tensor(hv, requires_grad=True)
        theta = torch.tensor(tv, requires_grad=True)
        lp = fn(h, theta)
        lp.sum().backward()
        outs.append((lp.detach().numpy(), h.grad.numpy(), theta.grad.numpy()))
    (lo, gho, gto), (ln, ghn, gtn) = outs
    np.testing.assert_allclose(ln, lo, rtol=1e-12)
    np.testing.assert_allclose(ghn, gho, rtol=1e-9, atol=1e-9 * np.abs(gho).max())
    np.testing.assert_allclose(gtn, gto, rtol=1e-10)
    # bit-wise reproducible
    again = constant_coalescent_log_prob(torch.tensor(hv), torch.tensor(tv))
    assert np.array_equal(again.numpy(), ln)


@pytest.mark.gpu
def test_unbatched_heights_with_batched_theta_and_limits():
    from oracle.coalescent import constant_log_prob
    from torchtree_b200.coalescent import constant_coalescent_log_prob

    rec = np.load(GOLDEN[0])
    h = torch.tensor(rec["heights"][0], requires_grad=True)          # [2T-1]
    theta = torch.tensor([[2.0], [3.5], [7.0]], dtype=torch.float64, requires_grad=True)  # [3,1]
    lp = constant_coalescent_log_prob(h, theta)
    assert lp.shape == (3, 1)
    lp.sum().backward()
    h2 = torch.tensor(rec["heights"][0], requires_grad=True)
    t2 = theta.detach().clone().requires_grad_(True)
    constant_log_prob(h2, t2).sum().backward()
    np.testing.assert_allclose(h.grad.numpy(), h2.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(theta.grad.numpy(), t2.grad.numpy(), rtol=1e-10)
    # beyond 4096 tips the sort runs on a global-memory scratch area: device tensors too
    big = torch.rand(2, 2 * 5000 - 1, dtype=torch.float64)
    on_host = constant_coalescent_log_prob(big, torch.ones(1, 1, dtype=torch.float64))
    on_dev = constant_coalescent_log_prob(big.cuda(), torch.ones(1, 1, dtype=torch.float64).cuda())
    assert torch.equal(on_host, on_dev.cpu())
    nan = torch.tensor(rec["heights"][:1]).clone()
    nan[0, -1] = float("nan")
    nan.requires_grad_(True)
    out = constant_coalescent_log_prob(nan, torch.ones(1, 1, dtype=torch.float64))
    assert torch.isnan(out).all()
    out.sum().backward()
    assert torch.isnan(nan.grad).all()
