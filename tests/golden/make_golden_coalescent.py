"""Golden vectors for the constant-population coalescent, from the REAL reference distribution
(torchtree/evolution/coalescent.py:88-134): value and autograd gradients.

    PYTHONPATH=oracle/dendropy_shim:/root/reference python tests/golden/make_golden_coalescent.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "dendropy_shim"))
sys.path.insert(0, "/root/reference")

from torchtree.evolution.coalescent import (ConstantCoalescent, PiecewiseConstantCoalescent,  # noqa: E402
                                            PiecewiseConstantCoalescentGrid)

torch.set_default_dtype(torch.float64)


def heights(rng, T, D, heterochronous):
    """Valid node heights of D random time trees on T tips: tips first, every internal node above
    both of its children (random-join topology, heights accumulated upwards)."""
    out = np.zeros((D, 2 * T - 1))
    for d in range(D):
        tips = rng.uniform(0.0, 5.0, T) if heterochronous else np.zeros(T)
        h = list(tips)
        active = list(range(T))
        for node in range(T, 2 * T - 1):
            i, j = rng.choice(len(active), 2, replace=False)
            a, b = active[i], active[j]
            h.append(max(h[a], h[b]) + rng.exponential(1.0))
            active = [x for k, x in enumerate(active) if k not in (i, j)] + [node]
        out[d] = h
    return out


def case(name, T, D, heterochronous, shared_theta, seed):
    rng = np.random.default_rng(seed)
    h = torch.tensor(heights(rng, T, D, heterochronous), requires_grad=True)
    theta = torch.tensor(rng.uniform(2.0, 9.0, (1 if shared_theta else D, 1)), requires_grad=True)
    lp = ConstantCoalescent(theta).log_prob(h)
    w = torch.tensor(rng.uniform(-1.0, 2.0, (D, 1)))
    (lp * w).sum().backward()
    np.savez(os.path.join(HERE, "coalescent", name + ".npz"), heights=h.detach().numpy(),
             theta=theta.detach().numpy(), log_prob=lp.detach().numpy(), grad_out=w.numpy(),
             d_heights=h.grad.numpy(), d_theta=theta.grad.numpy())
    print(name, lp.detach().numpy().reshape(-1)[:3])


def piecewise_case(name, T, D, heterochronous, shared_theta, seed, grid_points=0):
    """Skyride (grid_points = 0: T-1 population sizes, coalescent.py:311-396) or skygrid
    (coalescent.py:459-549) from the real reference distribution."""
    rng = np.random.default_rng(seed)
    hv = heights(rng, T, D, heterochronous)
    h = torch.tensor(hv, requires_grad=True)
    M = grid_points + 1 if grid_points else T - 1
    theta = torch.tensor(rng.uniform(1.0, 9.0, (1 if shared_theta else D, M)), requires_grad=True)
    if shared_theta:
        theta_in = theta.reshape(M) if D == 1 else theta.expand(D, M)
    else:
        theta_in = theta
    if grid_points:
        grid = torch.tensor(np.linspace(0.0, 0.8 * hv.max(), grid_points + 1)[1:])
        lp = PiecewiseConstantCoalescentGrid(theta_in, grid).log_prob(h)
    else:
        grid = torch.zeros(0)
        lp = PiecewiseConstantCoalescent(theta_in).log_prob(h)
    w = torch.tensor(rng.uniform(-1.0, 2.0, (D, 1)))
    (lp * w).sum().backward()
    np.savez(os.path.join(HERE, "coalescent_piecewise", name + ".npz"), heights=hv,
             theta=theta.detach().numpy(), grid=grid.numpy(), log_prob=lp.detach().numpy(),
             grad_out=w.numpy(), d_heights=h.grad.numpy(), d_theta=theta.grad.numpy())
    print(name, lp.detach().numpy().reshape(-1)[:3])


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "coalescent_piecewise"), exist_ok=True)
    piecewise_case("skyride_T12_D3", 12, 3, True, False, 11)
    piecewise_case("skyride_T60_D4_shared", 60, 4, True, True, 12)
    piecewise_case("skyride_iso_T33_D2", 33, 2, False, False, 13)
    piecewise_case("skygrid_T12_D3_G5", 12, 3, True, False, 21, grid_points=5)
    piecewise_case("skygrid_T60_D4_G20_shared", 60, 4, True, True, 22, grid_points=20)
    piecewise_case("skygrid_T200_D2_G50", 200, 2, True, False, 23, grid_points=50)
    case("hetero_T12_D3", 12, 3, True, False, 1)
    case("hetero_T60_D5_shared", 60, 5, True, True, 2)
    case("iso_T33_D4", 33, 4, False, False, 3)      # all tips at time 0: ties among the tips
    case("hetero_T500_D2", 500, 2, True, True, 4)
