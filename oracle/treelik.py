"""CPU ORACLE for the tree-likelihood hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A torch-CPU (fp64) restatement of the reference algorithm, function by function,
with the reference file:line each one follows.  The reference is pure Python on
torch ATen CPU ops with gradients from the autograd tape, so the
restatement uses the same substrate: batched matmuls for the conditional
likelihoods, `linalg.eigh` for P(t), and `.backward()` for the gradient.  That
keeps it usable as the `cpu_baseline` ("port") leg of bench.py.

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The product package
(`torchtree_b200`) never does; it fails loudly when the CUDA library is absent.

PARITY PINNING: this oracle is pinned against (a) the reference's own
known-answer tests (test/test_tree_likelihood.py:43-52 -> -83.329016,
:268-342 -> -4618.2062529058, test/test_substitution_model.py GTR/HKY matrices,
test/test_site_model.py Weibull rates) and (b) outputs of the real reference
imported from /root/reference at authoring time, committed as fixtures under
tests/golden/ by tests/golden/make_golden.py.  See tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch

torch_f64 = torch.float64


# ----------------------------------------------------------------------------
# transition matrices
# ----------------------------------------------------------------------------
def normalise_q(q_unnorm: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Q / (-sum_i pi_i Q_ii)  (substitution_model/abstract.py:49-50, :58-59)."""
    norm = -(torch.diagonal(q_unnorm, dim1=-2, dim2=-1) * freqs).sum(-1)
    return q_unnorm / norm[..., None, None]


def p_t_reversible(q_norm: torch.Tensor, freqs: torch.Tensor, t: torch.Tensor):
    """P(t) = (sqrt(pi)^-1 U) exp(e t) (U^-1 sqrt(pi)) with (e, U) = eigh of the
    symmetrised generator (substitution_model/abstract.py:57-76).

    q_norm [...,S,S], freqs [...,S], t [...,B,K]  ->  [...,B,K,S,S]
    """
    root = freqs.sqrt()
    sym = root[..., :, None] * q_norm / root[..., None, :]
    evals, u = torch.linalg.eigh(sym)
    left = u / root[..., :, None]  # sqrt_pi_inv @ v
    right = torch.linalg.inv(u) * root[..., None, :]  # v.inverse() @ sqrt_pi
    extra = t.dim() - evals.dim() + 1
    shape_v = evals.shape[:-1] + (1,) * extra
    decay = torch.exp(evals.reshape(shape_v + evals.shape[-1:]) * t.unsqueeze(-1))
    left = left.reshape(shape_v + left.shape[-2:])
    right = right.reshape(shape_v + right.shape[-2:])
    return (left * decay.unsqueeze(-2)) @ right


def p_t_jc69(t: torch.Tensor) -> torch.Tensor:
    """JC69 closed form (substitution_model/nucleotide.py:102-113)."""
    x = torch.exp(-4.0 / 3.0 * t)
    same = 0.25 + 0.75 * x
    diff = 0.25 - 0.25 * x
    eye = torch.eye(4, dtype=t.dtype)
    return same[..., None, None] * eye + diff[..., None, None] * (1.0 - eye)


def p_t_expm(q_norm: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """matrix_exp(Q t) route (substitution_model/abstract.py:89-94); also the
    truth for substitution-parameter gradients near degenerate spectra (F12)."""
    return torch.matrix_exp(q_norm[..., None, None, :, :] * t[..., None, None])


def gtr_q_unnorm(rates6: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Unnormalised GTR generator; order AC,AG,AT,CG,CT,GT
    (substitution_model/nucleotide.py:328-374)."""
    iu = torch.triu_indices(4, 4, 1)
    R = torch.zeros(rates6.shape[:-1] + (4, 4), dtype=rates6.dtype)
    R[..., iu[0], iu[1]] = rates6
    R = R + R.transpose(-1, -2)
    Q = R * freqs[..., None, :]
    return Q - torch.diag_embed(Q.sum(-1))


def hky_q_unnorm(kappa: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """HKY generator (substitution_model/nucleotide.py:238-264): transitions
    A<->G and C<->T scaled by kappa."""
    one = torch.ones_like(kappa[..., 0])
    k = kappa[..., 0]
    rates6 = torch.stack([one, k, one, one, k, one], -1)
    return gtr_q_unnorm(rates6, freqs)


def symmetric_q_unnorm(exch: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Reversible generator from S(S-1)/2 exchangeabilities (upper triangle,
    row-major) -- the EmpiricalSubstitutionModel / GeneralSymmetric layout
    (substitution_model/general.py:170-184, :345-355)."""
    S = freqs.shape[-1]
    iu = torch.triu_indices(S, S, 1)
    R = torch.zeros(exch.shape[:-1] + (S, S), dtype=exch.dtype)
    R[..., iu[0], iu[1]] = exch
    R = R + R.transpose(-1, -2)
    Q = R * freqs[..., None, :]
    return Q - torch.diag_embed(Q.sum(-1))


# ----------------------------------------------------------------------------
# site model
# ----------------------------------------------------------------------------
def weibull_site_model(shape, categories, invariant=None, mu=None):
    """Discretised Weibull rates / proportions
    (site_model.py:173-195 update_rates, :237-247 inverse_cdf)."""
    quant = (2.0 * torch.arange(categories, dtype=shape.dtype) + 1.0) / (
        2.0 * categories
    )
    r = torch.pow(-torch.log(1.0 - quant), 1.0 / shape)
    if invariant is not None:
        props = torch.cat(
            (
                invariant,
                ((1.0 - invariant) / categories).expand(
                    invariant.shape[:-1] + (categories,)
                ),
            ),
            -1,
        )
        r = torch.cat((torch.zeros_like(invariant), r), -1)
    else:
        props = torch.full((categories,), 1.0 / categories, dtype=shape.dtype)
    r = r / (r * props).sum(-1, keepdim=True)
    if mu is not None:
        r = r * mu
    return r, props


# ----------------------------------------------------------------------------
# peeling
# ----------------------------------------------------------------------------
def tip_partials_from_states(tip_states: np.ndarray, state_count: int, code_partials=None):
    """uint8 [T,N] codes -> list of T tensors [S,N] (one-hot; code>=S -> ones),
    the layout of site_pattern.compress_alignment (site_pattern.py:100-124)."""
    S = state_count
    if code_partials is None:
        table = np.concatenate([np.eye(S), np.ones((1, S))], 0)
        codes = np.minimum(tip_states.astype(np.int64), S)
    else:
        table = np.asarray(code_partials, dtype=np.float64)
        codes = tip_states.astype(np.int64)
    return [torch.from_numpy(np.ascontiguousarray(table[c].T)) for c in codes]


def log_likelihood(
    tips,
    weights: torch.Tensor,
    postorder,
    mats: torch.Tensor,
    freqs: torch.Tensor,
    props: torch.Tensor,
    rescale: bool = True,
    tip_states: bool = False,
):
    """Felsenstein pruning with rate categories.

    Restates the four reference variants behind two switches:
      rescale=False, tip_states=False -> calculate_treelikelihood_discrete
                                         (tree_likelihood.py:40-75)
      rescale=True,  tip_states=False -> ..._discrete_rescaled (:186-221)
      rescale=False, tip_states=True  -> ..._tip_states_discrete (:78-131)
      rescale=True,  tip_states=True  -> ..._tip_states_discrete_rescaled (:224-278)

    tips   : list of T tensors [S,N] (partials) or [N] int64 (states, gap = S)
    mats   : [...,B,K,S,S];  freqs [...,1,S];  props [...,K,1,1]
    returns: [...,1]
    """
    T = len(postorder) + 1
    store = list(tips) + [None] * (T - 1)
    if tip_states:
        ones = torch.ones(mats[..., :T, :, :, :].shape[:-1] + (1,), dtype=mats.dtype)
        gather_src = torch.cat((mats[..., :T, :, :, :], ones), -1)
    log_scale = None
    for node, left, right in postorder:
        node, left, right = int(node), int(left), int(right)
        sides = []
        for child in (left, right):
            if tip_states and child < T:
                sides.append(gather_src[..., child, :, :, store[child]])
            else:
                sides.append(mats[..., child, :, :, :] @ store[child])
        vec = sides[0] * sides[1]  # [...,K,S,N]
        if rescale:
            top = vec.flatten(-3, -2).max(-2, keepdim=True)[0]  # [...,1,N]
            vec = vec / top.unsqueeze(-2)
            log_scale = top.log() if log_scale is None else log_scale + top.log()
        store[node] = vec
    root = store[int(postorder[-1][0])]
    site = torch.log(freqs @ (props * root).sum(-3))  # [...,1,N]
    if rescale:
        site = site + log_scale
    return (site * weights).sum(-1)


def log_likelihood_no_categories(tips, weights, postorder, mats, freqs):
    """calculate_treelikelihood (tree_likelihood.py:14-37): mats [...,B,S,S]."""
    T = len(postorder) + 1
    store = list(tips) + [None] * (T - 1)
    for node, left, right in postorder:
        store[int(node)] = (mats[..., int(left), :, :] @ store[int(left)]) * (
            mats[..., int(right), :, :] @ store[int(right)]
        )
    return (torch.log(freqs @ store[int(postorder[-1][0])]) * weights).sum(-1)


# ----------------------------------------------------------------------------
# whole evaluation on a flattened Problem (torchtree_b200.synthetic.Problem)
# ----------------------------------------------------------------------------
def evaluate(problem, want_grad: bool = True, rescale: bool = True, tip_states=False,
             through_q: bool = True, route: str = "eigh"):
    """logL (and autograd gradient, like the reference: SURVEY 3.4) of a
    flattened problem.  Mirrors TreeLikelihoodModel._call
    (tree_likelihood.py:313-356): mats = p_t(bls[...,B,1] * rates[...,1,K]).

    route="eigh" is the reference's reversible route; its gradient w.r.t. the
    generator is only the part `eigh` sees (the lower triangle of the
    symmetrised matrix, returned symmetrised).  route="expm" evaluates
    P = matrix_exp(Q t) (abstract.py:89-94) whose autograd gives the full
    unconstrained d lnL / d Q -- the truth the engine's d_q is compared with.

    Returns dict with lnL [D] and gradients w.r.t. branch_lengths [D,B],
    site_rates, site_props, freqs (root term only when through_q, i.e. with Q
    held as an independent input) and q_matrix.
    """
    D = problem.draws
    bl = torch.tensor(problem.branch_lengths, dtype=torch_f64, requires_grad=want_grad)
    rates = torch.tensor(problem.site_rates, dtype=torch_f64, requires_grad=want_grad)
    props = torch.tensor(problem.site_props, dtype=torch_f64, requires_grad=want_grad)
    freqs = torch.tensor(problem.freqs, dtype=torch_f64, requires_grad=want_grad)
    q = torch.tensor(problem.q_matrix, dtype=torch_f64, requires_grad=want_grad)
    weights = torch.tensor(problem.weights, dtype=torch_f64)

    t = bl.unsqueeze(-1) * rates.expand(D, -1).unsqueeze(-2)  # [D,B,K]
    if route == "expm":
        mats = p_t_expm(q.expand(D, -1, -1), t)
    elif through_q:
        # Q is an independent input: P = expm(Q t) evaluated through the
        # reversible eigen route with a *detached* symmetrising pi, so that
        # d lnL / d freqs is the root term only and d lnL / d Q is the full
        # unconstrained derivative.
        mats = p_t_reversible(q.expand(D, -1, -1), freqs.detach().expand(D, -1), t)
    else:
        mats = p_t_reversible(q.expand(D, -1, -1), freqs.expand(D, -1), t)
    if tip_states:
        tips = [
            torch.from_numpy(np.minimum(row.astype(np.int64), problem.state_count))
            for row in problem.tip_states
        ]
    else:
        tips = tip_partials_from_states(
            problem.tip_states, problem.state_count, problem.code_partials
        )
    lnl = log_likelihood(
        tips,
        weights,
        problem.postorder,
        mats,
        freqs.expand(D, -1).unsqueeze(-2),
        props.expand(D, -1)[..., None, None],
        rescale=rescale,
        tip_states=tip_states,
    ).squeeze(-1)
    out = {"lnL": lnl.detach().numpy().copy()}
    if want_grad:
        lnl.sum().backward()
        out["branch_lengths"] = bl.grad.numpy().copy()
        out["site_rates"] = rates.grad.numpy().copy()
        out["site_props"] = props.grad.numpy().copy()
        out["freqs"] = freqs.grad.numpy().copy()
        out["q_matrix"] = q.grad.numpy().copy()
    return out


def evaluate_mats(problem, mats: np.ndarray, want_grad=True, rescale=True):
    """P-mode: matrices supplied directly [D,B,K,S,S]; gradient w.r.t. mats,
    freqs and props."""
    D = problem.draws
    m = torch.tensor(mats, dtype=torch_f64, requires_grad=want_grad)
    props = torch.tensor(problem.site_props, dtype=torch_f64, requires_grad=want_grad)
    freqs = torch.tensor(problem.freqs, dtype=torch_f64, requires_grad=want_grad)
    weights = torch.tensor(problem.weights, dtype=torch_f64)
    tips = tip_partials_from_states(
        problem.tip_states, problem.state_count, problem.code_partials
    )
    lnl = log_likelihood(
        tips, weights, problem.postorder, m,
        freqs.expand(D, -1).unsqueeze(-2), props.expand(D, -1)[..., None, None],
        rescale=rescale,
    ).squeeze(-1)
    out = {"lnL": lnl.detach().numpy().copy()}
    if want_grad:
        lnl.sum().backward()
        out["mats"] = m.grad.numpy().copy()
        out["site_props"] = props.grad.numpy().copy()
        out["freqs"] = freqs.grad.numpy().copy()
    return out


def transition_matrices(problem) -> np.ndarray:
    """[D,B,K,S,S] matrices of a flattened problem (reversible eigen route)."""
    D = problem.draws
    bl = torch.tensor(problem.branch_lengths, dtype=torch_f64)
    rates = torch.tensor(problem.site_rates, dtype=torch_f64).expand(D, -1)
    t = bl.unsqueeze(-1) * rates.unsqueeze(-2)
    q = torch.tensor(problem.q_matrix, dtype=torch_f64).expand(D, -1, -1)
    f = torch.tensor(problem.freqs, dtype=torch_f64).expand(D, -1)
    return p_t_reversible(q, f, t).numpy()
