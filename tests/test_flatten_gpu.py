"""The host-side glue (torchtree_b200.flatten.evaluate_models) driving the CUDA
engine from model objects, compared with what the real reference's
`TreeLikelihoodModel` produced for the same models (golden fixtures)."""
import numpy as np
import pytest
import torch

import standins as sm
from helpers import assert_grad_close, assert_lnl_close, load_golden

pytestmark = pytest.mark.gpu


def _engine(prob, draws=1):
    from torchtree_b200 import Engine

    return Engine(prob.tip_states, prob.weights, prob.postorder, prob.state_count,
                  prob.category_count, code_partials=prob.code_partials, max_draws=draws)


@pytest.mark.parametrize("name", ["fluA_gtr_w4_generic", "fluA_gtr_w4_ambig", "syn40_gtr_w4",
                                  "fluA_gtr_w4_batch3", "syn17_gtr_w3"])
@pytest.mark.parametrize("route", ["eigen", "expm", "mats"])
def test_unrooted_gtr_weibull_matches_reference_model(name, route):
    from torchtree_b200.flatten import evaluate_models, substitution_route

    prob, rec = load_golden(name)
    D = prob.draws
    blens = torch.tensor(rec["param_blens"], requires_grad=True)
    rates6 = torch.tensor(rec["param_gtr_rates"], requires_grad=True)
    freqs = torch.tensor(rec["param_gtr_freqs"], requires_grad=True)
    shape = torch.tensor(rec["param_shape"], requires_grad=True)
    tree = sm.UnRootedTreeModel(blens, prob.postorder)
    site = sm.WeibullSiteModel(shape, prob.category_count)
    subst = {"eigen": sm.GTR, "expm": sm.NonSymmetricSubstitutionModel,
             "mats": sm.UserDefinedSubstitutionModel}[route](rates6, freqs)
    assert substitution_route(subst) == route
    sample_shape = torch.Size([D]) if D > 1 else torch.Size([])
    eng = _engine(prob, D)
    lnl = evaluate_models(eng, tree, site, subst, None, sample_shape)
    assert lnl.shape == sample_shape + (1,)
    assert_lnl_close(lnl.detach().numpy().reshape(-1), rec["model_lnL"], what=name)
    lnl.sum().backward()
    assert_grad_close(blens.grad.numpy(), rec["dparam_blens"], what="d_blens")
    assert_grad_close(shape.grad.numpy(), rec["dparam_shape"], what="d_shape")
    assert_grad_close(rates6.grad.numpy(), rec["dparam_gtr_rates"], rtol=1e-7, what="d_rates6")
    assert_grad_close(freqs.grad.numpy(), rec["dparam_gtr_freqs"], rtol=1e-7, what="d_freqs")
    eng.close()


def test_hky_constant_site_model():
    from torchtree_b200.flatten import evaluate_models

    prob, rec = load_golden("fluA_hky_k1")
    blens = torch.tensor(rec["param_blens"], requires_grad=True)
    kappa = torch.tensor(rec["param_hky_kappa"], requires_grad=True)
    freqs = torch.tensor(rec["param_hky_freqs"], requires_grad=True)
    eng = _engine(prob)
    lnl = evaluate_models(eng, sm.UnRootedTreeModel(blens, prob.postorder),
                          sm.ConstantSiteModel(), sm.HKY(kappa, freqs), None, torch.Size([]))
    assert_lnl_close(lnl.detach().numpy(), rec["model_lnL"])
    lnl.sum().backward()
    assert_grad_close(blens.grad.numpy(), rec["dparam_blens"], what="d_blens")
    assert_grad_close(kappa.grad.numpy(), rec["dparam_hky_kappa"], rtol=1e-7, what="d_kappa")
    assert_grad_close(freqs.grad.numpy(), rec["dparam_hky_freqs"], rtol=1e-7, what="d_freqs")
    eng.close()


def test_time_tree_strict_clock_jc69_weibull():
    """test/test_tree_likelihood.py:268-342: -4618.2062529058, unbatched and as a
    batch of 3 identical draws."""
    from torchtree_b200.flatten import evaluate_models

    prob, rec = load_golden("fluA_jc69_w4_clock")
    clock_rate = 0.001
    lengths = torch.tensor(prob.branch_lengths[0] / clock_rate)
    shape = torch.tensor([0.1], dtype=torch.float64)
    eng = _engine(prob, 3)
    tree = sm.TimeTreeModel(lengths, prob.postorder)
    clock = sm.StrictClockModel(torch.tensor([clock_rate], dtype=torch.float64), prob.branch_count)
    lnl = evaluate_models(eng, tree, sm.WeibullSiteModel(shape, 4), sm.JC69(), clock, torch.Size([]))
    assert lnl.shape == (1,)
    assert abs(lnl.item() - (-4618.2062529058)) < 1e-6
    tree3 = sm.TimeTreeModel(lengths.repeat(3, 1), prob.postorder)
    clock3 = sm.StrictClockModel(torch.full((3, 1), clock_rate, dtype=torch.float64),
                                 prob.branch_count)
    lnl3 = evaluate_models(eng, tree3, sm.WeibullSiteModel(shape.repeat(3, 1), 4), sm.JC69(),
                           clock3, torch.Size([3]))
    assert lnl3.shape == (3, 1)
    assert torch.allclose(lnl3, torch.full((3, 1), -4618.2062529058, dtype=torch.float64),
                          rtol=1e-12, atol=0)
    eng.close()
