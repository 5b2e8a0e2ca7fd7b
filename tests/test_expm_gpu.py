"""Device matrix-exponential route (csrc/expm.cu, ttb2_loglik_expm):
general (non-reversible) generators through the C ABI against the oracle's
`torch.matrix_exp` route (what NonSymmetricSubstitutionModel.p_t computes,
substitution_model/abstract.py:89-94) -- lnL at 1e-10, every gradient at 1e-8, for state
counts on both sides of the shared-memory opt-in limit, long branches (many squarings),
zero-length branches (P = I exactly), batches of draws with shared and per-draw generators,
NaN propagation and bit-wise reproducibility."""
import numpy as np
import pytest
import torch

from helpers import assert_grad_close, assert_lnl_close
from oracle import treelik as orc
from torchtree_b200 import Engine, log_likelihood_expm
from torchtree_b200.synthetic import make_problem

pytestmark = pytest.mark.gpu


def general_generator(S, rng, draws=1):
    """Random non-reversible generators, normalised as the reference does (abstract.py:90-91):
    -sum_i pi_i Q_ii = 1."""
    q = rng.gamma(2.0, 1.0, size=(draws, S, S))
    pi = rng.dirichlet(np.full(S, 5.0), size=draws)
    for d in range(draws):
        np.fill_diagonal(q[d], 0.0)
        np.fill_diagonal(q[d], -q[d].sum(1))
        q[d] /= -(np.diag(q[d]) * pi[d]).sum()
    return q, pi


def _problem(T, N, S, K, D=1, seed=1, per_draw=False, mean_branch=0.026):
    prob = make_problem(T, N, S, K, draws=D, seed=seed, gap_fraction=0.03, mean_branch=mean_branch)
    rng = np.random.default_rng(seed + 100)
    q, pi = general_generator(S, rng, D if per_draw else 1)
    prob.q_matrix, prob.freqs = q, pi
    return prob


def _run(prob):
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, prob.state_count,
                 prob.category_count, max_draws=prob.draws)
    lnl = eng.loglik_expm(prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix,
                          prob.freqs)
    g = eng.grad_eigen()
    return eng, lnl, g


@pytest.mark.parametrize("S,T,N,K", [(2, 5, 17, 1), (4, 12, 100, 4), (7, 9, 33, 2), (20, 8, 40, 4),
                                     (33, 6, 20, 1), (61, 5, 16, 2), (64, 4, 9, 1)])
def test_expm_route_matches_matrix_exp_oracle(S, T, N, K):
    prob = _problem(T, N, S, K, seed=S)
    eng, lnl, g = _run(prob)
    want = orc.evaluate(prob, want_grad=True, route="expm")
    assert_lnl_close(lnl.numpy(), want["lnL"])
    for key, ref in (("branch_lengths", "branch_lengths"), ("site_rates", "site_rates"),
                     ("props", "site_props"), ("freqs", "freqs"), ("q", "q_matrix")):
        assert_grad_close(g[key].numpy(), want[ref], what="S=%d %s" % (S, key))
    eng.close()


def test_long_and_zero_branches():
    """Branch lengths from 0 (P = I exactly) to 40 expected substitutions (11 squarings)."""
    prob = _problem(10, 60, 5, 2, seed=3)
    prob.branch_lengths[0, :] = np.geomspace(1e-6, 40.0, prob.branch_count)
    prob.branch_lengths[0, 2] = 0.0
    prob.branch_lengths[0, -1] = 0.0
    eng, lnl, g = _run(prob)
    mats = eng.get_mats().numpy()
    assert np.array_equal(mats[0, 2, 0], np.eye(5))
    assert np.allclose(mats.sum(-1), 1.0, rtol=0, atol=1e-13)   # rows of exp(Q t) sum to one
    want = orc.evaluate(prob, want_grad=True, route="expm")
    assert_lnl_close(lnl.numpy(), want["lnL"])
    assert_grad_close(g["branch_lengths"].numpy(), want["branch_lengths"], what="bl")
    assert_grad_close(g["q"].numpy(), want["q_matrix"], what="q")
    eng.close()


@pytest.mark.parametrize("per_draw", [False, True])
def test_batches_of_draws(per_draw):
    prob = _problem(9, 50, 6, 3, D=4, seed=11, per_draw=per_draw)
    eng, lnl, g = _run(prob)
    want = orc.evaluate(prob, want_grad=True, route="expm")
    assert_lnl_close(lnl.numpy(), want["lnL"])
    for key, ref in (("branch_lengths", "branch_lengths"), ("q", "q_matrix"), ("freqs", "freqs")):
        assert g[key].shape == want[ref].shape
        assert_grad_close(g[key].numpy(), want[ref], what=key)
    # weighted draws through the torch extension, device tensors
    dev = torch.device("cuda", 0)
    leaves = [torch.tensor(x, device=dev, requires_grad=True) for x in
              (prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix, prob.freqs)]
    w = torch.linspace(0.5, 2.0, 4, dtype=torch.float64, device=dev)
    (log_likelihood_expm(eng, *leaves) * w).sum().backward()
    ref_leaves = [torch.tensor(x, requires_grad=True) for x in
                  (prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix, prob.freqs)]
    D = prob.draws
    t = ref_leaves[0].unsqueeze(-1) * ref_leaves[1].expand(D, -1).unsqueeze(-2)
    mats = orc.p_t_expm(ref_leaves[3].expand(D, -1, -1), t)
    tips = orc.tip_partials_from_states(prob.tip_states, 6, prob.code_partials)
    lo = orc.log_likelihood(tips, torch.tensor(prob.weights), prob.postorder, mats,
                            ref_leaves[4].expand(D, -1).unsqueeze(-2),
                            ref_leaves[2].expand(D, -1)[..., None, None]).squeeze(-1)
    (lo * w.cpu()).sum().backward()
    for a, b, name in zip(leaves, ref_leaves, ("bl", "rates", "props", "q", "freqs")):
        assert_grad_close(a.grad.cpu().numpy(), b.grad.numpy(), what=name)
    eng.close()


def test_nan_in_nan_out_and_reproducible():
    prob = _problem(8, 40, 5, 2, seed=5)
    eng, lnl, g = _run(prob)
    lnl2 = eng.loglik_expm(prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix,
                           prob.freqs)
    g2 = eng.grad_eigen()
    assert torch.equal(lnl, lnl2)
    for k in g:
        assert torch.equal(g[k], g2[k]), k
    prob.q_matrix[0, 1, 2] = np.nan
    lnl3 = eng.loglik_expm(prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix,
                           prob.freqs)
    assert torch.isnan(lnl3).all()
    eng.close()
