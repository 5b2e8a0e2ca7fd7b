#!/usr/bin/env python
"""Ground truth for the substitution-parameter gradients, independent of every fp64 code path:
the tree likelihood of a small GTR + Weibull problem evaluated with mpmath at 50 digits
(P = expm(Q r t) by mpmath's own matrix exponential, Felsenstein pruning in exact order) and
differentiated with mpmath's high-order numerical differentiation at that precision.

Why: the parity tests compare the engine's GTR-parameter gradients with the reference's
autograd at 1e-7 instead of north_star's 1e-8, on the grounds that the *reference's* `eigh`
backward divides by eigenvalue gaps.  This fixture lets a GPU test show which
side the slack belongs to: two points, one generic and one where two eigenvalues of the
symmetrised generator are 1e-7 apart, each with the truth and with what the real reference
(imported here from /root/reference or baseline/_ref) returns.

    python tests/golden/make_truth_mpmath.py      # writes tests/golden/truth/truth_mpmath.npz
"""
import os
import sys

import mpmath as mp
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import refenv  # noqa: E402

mp.mp.dps = 50
T, N, K = 6, 7, 3
POST = [(6, 0, 1), (7, 2, 3), (8, 6, 7), (9, 4, 5), (10, 8, 9)]       # (node, left, right)
TIPS = np.array([[0, 1, 2, 3, 0, 1, 4],                                 # 4 = gap
                 [0, 1, 2, 3, 1, 1, 2],
                 [0, 2, 2, 3, 0, 3, 2],
                 [1, 1, 2, 0, 0, 1, 2],
                 [0, 1, 3, 3, 2, 1, 0],
                 [0, 1, 2, 3, 0, 4, 2]], dtype=np.uint8)
WEIGHTS = np.array([3.0, 1.0, 2.0, 1.0, 1.0, 2.0, 1.0])
B = 2 * T - 2


def lnl_mp(theta):
    """theta = [bl (2T-3) | shape | rates6 | freqs4] as mpf."""
    bl = list(theta[:2 * T - 3]) + [mp.mpf(0)]
    shape = theta[2 * T - 3]
    r6 = theta[2 * T - 2:2 * T + 4]
    pi = theta[2 * T + 4:2 * T + 8]
    quant = [(2 * mp.mpf(i) + 1) / (2 * K) for i in range(K)]
    rates = [mp.power(-mp.log(1 - q), 1 / shape) for q in quant]
    mean = sum(rates) / K
    rates = [r / mean for r in rates]
    R = mp.zeros(4, 4)
    idx = 0
    for i in range(4):
        for j in range(i + 1, 4):
            R[i, j] = R[j, i] = r6[idx]
            idx += 1
    Q = mp.zeros(4, 4)
    for i in range(4):
        for j in range(4):
            if i != j:
                Q[i, j] = R[i, j] * pi[j]
        Q[i, i] = -sum(Q[i, j] for j in range(4) if j != i)
    norm = -sum(pi[i] * Q[i, i] for i in range(4))
    Q = Q / norm
    total = mp.mpf(0)
    mats = [[mp.expm(Q * (bl[b] * rates[k])) for k in range(K)] for b in range(B)]
    for site in range(N):
        like = mp.mpf(0)
        for k in range(K):
            part = {}
            for t in range(T):
                c = TIPS[t, site]
                part[t] = [mp.mpf(1)] * 4 if c >= 4 else [mp.mpf(1 if s == c else 0) for s in range(4)]
            for node, l, r in POST:
                ul = [sum(mats[l][k][s, x] * part[l][x] for x in range(4)) for s in range(4)]
                ur = [sum(mats[r][k][s, x] * part[r][x] for x in range(4)) for s in range(4)]
                part[node] = [ul[s] * ur[s] for s in range(4)]
            like += sum(pi[s] * part[POST[-1][0]][s] for s in range(4)) / K
        total += mp.mpf(WEIGHTS[site]) * mp.log(like)
    return total


def truth(theta0):
    theta0 = [mp.mpf(float(x)) for x in theta0]
    value = lnl_mp(theta0)
    grad = []
    for i in range(len(theta0)):
        def f(x, i=i):
            th = list(theta0)
            th[i] = x
            return lnl_mp(th)
        grad.append(mp.diff(f, theta0[i], h=mp.mpf(10) ** -15))
    return float(value), np.array([float(g) for g in grad])


def reference(theta0):
    """What the real reference returns at the same point (TreeLikelihood functions + GTR +
    WeibullSiteModel, autograd)."""
    refenv.activate()
    torch.set_default_dtype(torch.float64)
    from torchtree import Parameter
    from torchtree.evolution.site_model import WeibullSiteModel
    from torchtree.evolution.substitution_model.nucleotide import GTR
    from torchtree.evolution.tree_likelihood import calculate_treelikelihood_discrete

    nb = 2 * T - 3
    bl = Parameter("bl", torch.tensor(theta0[:nb]))
    shape = Parameter("shape", torch.tensor(theta0[nb:nb + 1]))
    rates = Parameter("rates", torch.tensor(theta0[nb + 1:nb + 7]))
    freqs = Parameter("freqs", torch.tensor(theta0[nb + 7:nb + 11]))
    for p in (bl, shape, rates, freqs):
        p.requires_grad = True
    gtr = GTR("gtr", rates, freqs)
    site = WeibullSiteModel("site", shape, K)
    r = site.rates().reshape(1, -1)
    bls = torch.cat((bl.tensor, torch.zeros(1)))
    mats = gtr.p_t(bls.reshape(-1, 1) * r)
    table = np.concatenate([np.eye(4), np.ones((1, 4))], 0)
    partials = [torch.tensor(table[c].T.copy()) for c in TIPS] + [None] * (T - 1)
    lnl = calculate_treelikelihood_discrete(
        partials, torch.tensor(WEIGHTS), POST, mats, gtr.frequencies.reshape(1, -1),
        site.probabilities().unsqueeze(-1).unsqueeze(-1))
    lnl.sum().backward()
    grad = torch.cat([bl.grad, shape.grad, rates.grad, freqs.grad]).numpy()
    return float(lnl.sum()), grad


def main():
    rng = np.random.default_rng(42)
    nb = 2 * T - 3
    cases = {}
    generic = np.concatenate([rng.uniform(0.02, 0.3, nb), [0.7],
                              [0.9, 3.1, 0.6, 1.3, 4.2, 1.0], [0.33, 0.19, 0.22, 0.26]])
    # near-degenerate spectrum: K80-like rates (two equal eigenvalues at equal frequencies),
    # split by 1e-7 through the frequencies
    near = np.concatenate([rng.uniform(0.02, 0.3, nb), [0.7],
                           [1.0, 2.5, 1.0, 1.0, 2.5, 1.0],
                           [0.25 + 1e-7, 0.25 - 1e-7, 0.25 + 2e-7, 0.25 - 2e-7]])
    for name, theta in (("generic", generic), ("near_degenerate", near)):
        v, g = truth(theta)
        vr, gr = reference(theta)
        cases[name + "_theta"] = theta
        cases[name + "_lnL"] = np.array(v)
        cases[name + "_grad"] = g
        cases[name + "_ref_lnL"] = np.array(vr)
        cases[name + "_ref_grad"] = gr
        err = np.abs(gr - g) / np.maximum(np.abs(g), 1e-8 * np.abs(g).max())
        print(name, "lnL", v, "reference rel err of lnL %.2e" % (abs(vr - v) / abs(v)),
              "max rel err of the reference gradient: bl %.2e shape %.2e rates %.2e freqs %.2e"
              % (err[:nb].max(), err[nb], err[nb + 1:nb + 7].max(), err[nb + 7:].max()))
    np.savez(os.path.join(HERE, "truth", "truth_mpmath.npz"), T=T, N=N, K=K, postorder=np.array(POST),
             tips=TIPS, weights=WEIGHTS, **cases)


if __name__ == "__main__":
    main()
