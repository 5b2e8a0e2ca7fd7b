"""Config 3 end to end on the device side: ratios/root height --(ttb2_heights)--> node
heights --> branch lengths x strict-clock rate --(engine, batch of draws)--> lnL, with
gradients back to the ratios, the root height and the clock rate; compared with the same
pipeline built from the CPU oracles (oracle/heights.py + oracle/treelik.py, autograd)."""
import numpy as np
import pytest
import torch

import standins as sm
from oracle import heights as oh
from oracle import treelik as orc

pytestmark = pytest.mark.gpu


def _time_tree_problem(T, N, D, seed):
    from torchtree_b200.synthetic import make_problem
    rng = np.random.default_rng(seed)
    prob = make_problem(T, N, 4, 1, seed=seed)          # only the tree, tips and weights are used
    post = prob.postorder
    times = torch.tensor(rng.uniform(0.0, 3.0, T) * (rng.random(T) < 0.6))
    bounds = oh.internal_bounds(T, post, times)
    x = torch.rand(D, T - 1, dtype=torch.float64) * 0.8 + 0.1
    root = int(post[-1][0]) - T
    x[:, root] = 8.0 + 4.0 * torch.rand(D, dtype=torch.float64)
    rate = torch.full((D, 1), 0.02, dtype=torch.float64) * (1.0 + 0.1 * torch.rand(D, 1, dtype=torch.float64))
    parent = oh.parents_from_postorder(T, post)
    nodes = [n for n in range(2 * T - 1) if n != int(post[-1][0])]     # branch b = node b
    return prob, post, times, bounds, x, rate, torch.tensor(nodes), torch.tensor([parent[n] for n in nodes])


def _branch_lengths(heights, times, child_idx, parent_idx):
    """tree_model.py:407-424: height[parent] - height[node] for every non-root node."""
    all_heights = torch.cat((times.expand(heights.shape[:-1] + (-1,)), heights), -1)
    return all_heights[..., parent_idx] - all_heights[..., child_idx]


@pytest.mark.parametrize("T,N,D", [(40, 300, 3), (12, 64, 1)])
def test_ratios_to_likelihood_gradients(T, N, D):
    from torchtree_b200 import Engine
    from torchtree_b200.flatten import evaluate_models
    from torchtree_b200.height_transform import NodeHeightPlan, node_heights

    prob, post, times, bounds, x, rate, child_idx, parent_idx = _time_tree_problem(T, N, D, 11 + T)
    B = 2 * T - 2

    # --- oracle pipeline (CPU, autograd) ---
    xo = x.clone().requires_grad_(True)
    ro = rate.clone().requires_grad_(True)
    ho = oh.heights_from_ratios(T, post, bounds, xo)
    blo = _branch_lengths(ho, times, child_idx, parent_idx) * ro
    tips = orc.tip_partials_from_states(prob.tip_states, 4, prob.code_partials)
    mats = orc.p_t_jc69(blo.reshape(D, B, 1))
    freqs = torch.full((1, 4), 0.25, dtype=torch.float64)
    lo = orc.log_likelihood(tips, torch.tensor(prob.weights), post, mats, freqs,
                            torch.ones(1, 1, 1, dtype=torch.float64))
    total_o = lo.sum() + oh.log_abs_det_jacobian(T, post, bounds, ho).sum()
    total_o.backward()

    # --- device pipeline ---
    eng = Engine(prob.tip_states, prob.weights, post, 4, 1, code_partials=prob.code_partials,
                 max_draws=D)
    plan = NodeHeightPlan(T, post, bounds.numpy())
    xn = x.clone().requires_grad_(True)
    rn = rate.clone().requires_grad_(True)
    hn = node_heights(xn, plan)
    bln = _branch_lengths(hn, times, child_idx, parent_idx)
    tree = sm.TimeTreeModel(bln, post)
    clock = sm.StrictClockModel(rn, B)
    ln = evaluate_models(eng, tree, sm.ConstantSiteModel(), sm.JC69(), clock, torch.Size([D]))
    assert ln.shape == (D, 1)
    ld = torch.log(hn[..., parent_idx[T:] - T] - bounds[child_idx[T:] - T]).sum(-1)
    (ln.sum() + ld.sum()).backward()

    assert np.allclose(ln.detach().numpy().reshape(-1), lo.detach().numpy().reshape(-1),
                       rtol=1e-10, atol=0)
    for got, want, what in ((xn.grad, xo.grad, "d ratios/root height"), (rn.grad, ro.grad, "d clock rate")):
        scale = want.abs().max()
        assert (got - want).abs().max() <= 1e-8 * scale, what
    eng.close()
    plan.close()
