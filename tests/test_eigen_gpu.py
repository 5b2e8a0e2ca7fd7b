"""Device eigen-decomposition of reversible generators (ttb2_loglik_q, csrc/eigen.cu;
SURVEY 8(f) row f4) against torch.linalg.eigh -- what the reference's
SymmetricSubstitutionModel.p_t calls (substitution_model/abstract.py:57-66) -- and against
the engine's own host-eigen entry point."""
import numpy as np
import pytest
import torch

from helpers import assert_grad_close, assert_lnl_close

pytestmark = pytest.mark.gpu


def _engine(prob, flags=0):
    from torchtree_b200 import Engine

    return Engine(prob.tip_states, prob.weights, prob.postorder, prob.state_count,
                  prob.category_count, code_partials=prob.code_partials, max_draws=prob.draws,
                  flags=flags)


def _host_eigen(q, f):
    from torchtree_b200 import reversible_eigensystem

    return reversible_eigensystem(q, f.expand(q.shape[0], -1) if f.shape[0] != q.shape[0] else f)


@pytest.mark.parametrize("S", [2, 4, 5, 20, 21, 61, 64])
@pytest.mark.parametrize("per_draw_model", [False, True])
def test_device_eigensystem_matches_lapack(S, per_draw_model):
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(10, 40, S, 2, draws=3, seed=100 + S, per_draw_model=per_draw_model)
    eng = _engine(prob)
    args = [torch.tensor(a) for a in (prob.branch_lengths, prob.site_rates, prob.site_props)]
    q, f = torch.tensor(prob.q_matrix), torch.tensor(prob.freqs)
    evec_h, ivec_h, evals_h = _host_eigen(q, f)
    lnl_h = eng.loglik_eigen(*args, evec_h, ivec_h, evals_h, f).clone()
    g_h = {k: v.clone() for k, v in eng.grad_eigen().items()}

    lnl_d = eng.loglik_q(*args, q, f)
    evec, ivec, evals = eng.get_eigen()
    assert evec.shape == (q.shape[0], S, S)
    # ascending eigenvalues equal to LAPACK's; Q = V L V^-1 and V V^-1 = I to rounding
    assert (evals[:, 1:] >= evals[:, :-1]).all()
    scale = float(q.abs().max())
    assert float((evals - evals_h).abs().max()) <= 2e-14 * S * scale
    rec = evec @ torch.diag_embed(evals) @ ivec
    assert float((rec - q).abs().max()) <= 2e-14 * S * scale
    eye = torch.eye(S, dtype=torch.float64)
    assert float((evec @ ivec - eye).abs().max()) <= 2e-14 * S
    # the zero eigenvalue of a generator
    assert float(evals[:, -1].abs().max()) <= 1e-14 * S * scale

    assert_lnl_close(lnl_d.numpy(), lnl_h.numpy(), rtol=1e-12)
    g_d = eng.grad_eigen()
    for k in g_h:
        # d lnL/dQ goes through divided differences of the eigenvalues: it is as
        # accurate as the eigenvalue gaps allow, for either decomposition
        assert_grad_close(g_d[k].numpy(), g_h[k].numpy(), rtol=1e-7 if k == "q" else 1e-9, what=k)
    eng.close()


def test_degenerate_spectrum_jc69():
    """JC69: a three-fold eigenvalue.  Jacobi needs no gap; the gradient stays finite."""
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(16, 200, 4, 1, seed=9)
    eng = _engine(prob)
    q = torch.full((1, 4, 4), 1.0 / 3.0, dtype=torch.float64)
    q[0].fill_diagonal_(-1.0)
    f = torch.full((1, 4), 0.25, dtype=torch.float64)
    args = [torch.tensor(a) for a in (prob.branch_lengths, prob.site_rates, prob.site_props)]
    lnl = eng.loglik_q(*args, q, f)
    evec, ivec, evals = eng.get_eigen()
    assert torch.allclose(evals[0], torch.tensor([-4 / 3, -4 / 3, -4 / 3, 0.0], dtype=torch.float64),
                          atol=1e-15)
    # closed-form JC69 matrices (nucleotide.py:102-113) through the matrices entry point
    t = (args[0].unsqueeze(-1) * args[1].unsqueeze(-2))[..., None, None]
    e = torch.exp(-4.0 / 3.0 * t)
    mats = (0.25 - 0.25 * e).expand(-1, -1, -1, 4, 4).clone()
    idx = torch.arange(4)
    mats[..., idx, idx] = (0.25 + 0.75 * e).expand(-1, -1, -1, 4, 4)[..., idx, idx]
    lnl_m = eng.loglik_mats(mats, f, args[2])
    assert_lnl_close(lnl.numpy(), lnl_m.numpy(), rtol=1e-12)
    eng.loglik_q(*args, q, f)
    g = eng.grad_eigen()
    assert all(torch.isfinite(v).all() for v in g.values())
    eng.close()


def test_nan_generator_gives_nan_likelihood():
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(8, 50, 4, 2, seed=2)
    eng = _engine(prob)
    args = [torch.tensor(a) for a in (prob.branch_lengths, prob.site_rates, prob.site_props)]
    q, f = torch.tensor(prob.q_matrix).clone(), torch.tensor(prob.freqs)
    assert torch.isfinite(eng.loglik_q(*args, q, f)).all()
    q[0, 2, 1] = float("nan")
    assert torch.isnan(eng.loglik_q(*args, q, f)).all()
    eng.close()


@pytest.mark.parametrize("flags", [0, 32], ids=["graph", "nograph"])
def test_changing_generators_through_graph_replay(flags):
    """Small problems replay a captured kernel sequence: the eigen kernel is part of it and
    must see every new generator."""
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(20, 100, 4, 4, seed=11)
    eng = _engine(prob, flags=flags)
    args = [torch.tensor(a) for a in (prob.branch_lengths, prob.site_rates, prob.site_props)]
    f = torch.tensor(prob.freqs)
    rng = np.random.default_rng(0)
    for _ in range(4):
        r = rng.uniform(0.2, 2.0, (4, 4))
        r = torch.tensor(r + r.T)
        q = r * f[0][None, :]
        q.fill_diagonal_(0.0)
        q = q - torch.diag(q.sum(-1))
        q = (q / -(torch.diagonal(q) * f[0]).sum()).unsqueeze(0)
        lnl_d = eng.loglik_q(*args, q, f).clone()
        lnl_h = eng.loglik_eigen(*args, *_host_eigen(q, f), f)
        assert_lnl_close(lnl_d.numpy(), lnl_h.numpy(), rtol=1e-12)
    eng.close()
