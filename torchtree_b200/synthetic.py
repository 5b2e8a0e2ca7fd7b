"""Seeded synthetic phylogenetic problems (numpy only; shared by tests, oracle and bench).

Index conventions follow the reference (SURVEY Appendix A;
torchtree/evolution/tree_model.py:37-53, :187-195):

* tips are nodes ``0..T-1``; internal nodes ``T..2T-2`` numbered in post-order
  visit order; the root is ``2T-2``;
* ``postorder`` is a list of ``(node, left, right)`` triples for the internal
  nodes in post-order;
* branch ``b`` is the edge above node ``b`` (``B = 2T-2`` branches); for the
  unrooted parameterisation the last branch (node ``2T-3``, always a child of
  the root) has length zero (torchtree/evolution/tree_likelihood.py:323-337).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


def random_join_topology(tip_count: int, rng: np.random.Generator):
    """Random binary tree from successive random joins of the active set."""
    children = {}
    active = list(range(tip_count))
    next_id = tip_count
    while len(active) > 1:
        i, j = rng.choice(len(active), size=2, replace=False)
        a, b = active[i], active[j]
        children[next_id] = (a, b)
        active = [x for idx, x in enumerate(active) if idx != i and idx != j]
        active.append(next_id)
        next_id += 1
    return children, active[0]


def caterpillar_topology(tip_count: int):
    children = {}
    prev = 0
    next_id = tip_count
    for t in range(1, tip_count):
        children[next_id] = (prev, t)
        prev = next_id
        next_id += 1
    return children, prev


def balanced_topology(tip_count: int):
    children = {}
    layer = list(range(tip_count))
    next_id = tip_count
    while len(layer) > 1:
        nxt = []
        for a in range(0, len(layer) - 1, 2):
            children[next_id] = (layer[a], layer[a + 1])
            nxt.append(next_id)
            next_id += 1
        if len(layer) % 2 == 1:
            nxt.append(layer[-1])
        layer = nxt
    return children, layer[0]


def postorder_from_children(children: dict, root: int, tip_count: int) -> np.ndarray:
    """Relabel internal nodes in post-order visit order (reference convention)
    and return int32 triples ``[I,3]`` = (node, left, right)."""
    order = []
    stack = [(root, False)]
    while stack:
        node, expanded = stack.pop()
        if node < tip_count:
            continue
        if expanded:
            order.append(node)
        else:
            stack.append((node, True))
            l, r = children[node]
            stack.append((r, False))
            stack.append((l, False))
    relabel = {old: tip_count + i for i, old in enumerate(order)}

    def lab(x):
        return x if x < tip_count else relabel[x]

    triples = np.array(
        [(lab(n), lab(children[n][0]), lab(children[n][1])) for n in order],
        dtype=np.int32,
    )
    return triples


def random_postorder(tip_count: int, rng: np.random.Generator, topology: str = "random") -> np.ndarray:
    """Post-order triples of a synthetic topology ("random", "caterpillar", "balanced")."""
    if topology == "caterpillar":
        children, root = caterpillar_topology(tip_count)
    elif topology == "balanced":
        children, root = balanced_topology(tip_count)
    else:
        children, root = random_join_topology(tip_count, rng)
    return postorder_from_children(children, root, tip_count)


@dataclass
class Problem:
    """Flattened inputs of one tree-likelihood evaluation (what the engine,
    the oracle and the reference all consume)."""

    tip_count: int
    pattern_count: int
    state_count: int
    category_count: int
    postorder: np.ndarray  # int32 [I,3]
    tip_states: np.ndarray  # uint8 [T,N]; code >= S means gap / unknown
    weights: np.ndarray  # float64 [N]
    branch_lengths: np.ndarray  # float64 [D,B] (B = 2T-2, already x clock rate)
    site_rates: np.ndarray  # float64 [D or 1,K]
    site_props: np.ndarray  # float64 [D or 1,K]
    freqs: np.ndarray  # float64 [D or 1,S]
    q_matrix: Optional[np.ndarray] = None  # float64 [D or 1,S,S], normalised
    model: str = "GTR"
    model_params: dict = field(default_factory=dict)
    code_partials: Optional[np.ndarray] = None  # float64 [C,S] tip code -> partial

    @property
    def draws(self) -> int:
        return self.branch_lengths.shape[0]

    @property
    def branch_count(self) -> int:
        return 2 * self.tip_count - 2

    @property
    def units(self) -> int:
        """patterns x internal nodes x categories x draws (BASELINE.json metric unit)."""
        return (
            self.pattern_count
            * (self.tip_count - 1)
            * self.category_count
            * self.draws
        )


def weibull_rates(shape: float, categories: int) -> tuple:
    """Median-quantile discretised Weibull(scale 1) rates, mean-normalised
    (torchtree/evolution/site_model.py:173-195, :237-247)."""
    if categories == 1:
        return np.ones(1), np.ones(1)
    quant = (2.0 * np.arange(categories) + 1.0) / (2.0 * categories)
    rates = np.power(-np.log(1.0 - quant), 1.0 / shape)
    props = np.full(categories, 1.0 / categories)
    rates = rates / np.sum(rates * props)
    return rates, props


def gtr_q(rates6: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    """Normalised GTR rate matrix, upper-triangle order AC,AG,AT,CG,CT,GT
    (torchtree/evolution/substitution_model/nucleotide.py:328-374,
    abstract.py:49-50, :58-59)."""
    a, b, c, d, e, f = rates6
    R = np.array(
        [[0, a, b, c], [a, 0, d, e], [b, d, 0, f], [c, e, f, 0]], dtype=np.float64
    )
    Q = R * freqs[None, :]
    np.fill_diagonal(Q, -Q.sum(axis=1))
    norm = -np.sum(np.diag(Q) * freqs)
    return Q / norm


def symmetric_q(exch: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    """Normalised reversible rate matrix from S(S-1)/2 exchangeabilities
    (row-major upper triangle) and frequencies."""
    S = freqs.shape[0]
    R = np.zeros((S, S))
    iu = np.triu_indices(S, 1)
    R[iu] = exch
    R = R + R.T
    Q = R * freqs[None, :]
    np.fill_diagonal(Q, -Q.sum(axis=1))
    norm = -np.sum(np.diag(Q) * freqs)
    return Q / norm


def make_problem(
    tip_count: int,
    pattern_count: int,
    state_count: int = 4,
    category_count: int = 4,
    draws: int = 1,
    seed: int = 20260101,
    topology: str = "random",
    tips: str = "iid",
    gap_fraction: float = 0.0,
    weibull_shape: float = 0.5,
    per_draw_model: bool = False,
    mean_branch: float = 0.026,
) -> Problem:
    """Synthetic problem of SURVEY 8(d) shape: random-join tree, branch lengths
    ~ U(0.001, 0.051), integer weights U{1..4}, Dirichlet frequencies and
    exchangeabilities, Weibull site rates."""
    rng = np.random.default_rng(seed)
    T, N, S, K, D = tip_count, pattern_count, state_count, category_count, draws
    if topology == "random":
        children, root = random_join_topology(T, rng)
    elif topology == "caterpillar":
        children, root = caterpillar_topology(T)
    elif topology == "balanced":
        children, root = balanced_topology(T)
    else:
        raise ValueError(topology)
    post = postorder_from_children(children, root, T)
    B = 2 * T - 2
    lo, hi = mean_branch - 0.025, mean_branch + 0.025
    bl = rng.uniform(max(lo, 1e-4), hi, size=(D, B))
    bl[:, B - 1] = 0.0  # unrooted convention: padded zero branch (node 2T-3)

    nq = D if per_draw_model else 1
    freqs = rng.dirichlet(np.full(S, 5.0), size=nq)
    n_ex = S * (S - 1) // 2
    exch = rng.dirichlet(np.full(n_ex, 3.0), size=nq) * n_ex
    q = np.stack([symmetric_q(exch[i], freqs[i]) for i in range(nq)])

    nshape = D if per_draw_model else 1
    shapes = weibull_shape * (1.0 + 0.1 * np.arange(nshape))
    rp = [weibull_rates(s, K) for s in shapes]
    site_rates = np.stack([r for r, _ in rp])
    site_props = np.stack([p for _, p in rp])

    if tips == "iid":
        states = rng.integers(0, S, size=(T, N), dtype=np.int64)
    elif tips == "evolved":
        states = _evolve_tips(post, T, N, S, bl[0], q[0], freqs[0], rng)
    else:
        raise ValueError(tips)
    if gap_fraction > 0:
        mask = rng.random((T, N)) < gap_fraction
        states = np.where(mask, S, states)
    weights = rng.integers(1, 5, size=N).astype(np.float64)

    return Problem(
        tip_count=T,
        pattern_count=N,
        state_count=S,
        category_count=K,
        postorder=post,
        tip_states=states.astype(np.uint8),
        weights=weights,
        branch_lengths=bl,
        site_rates=site_rates,
        site_props=site_props,
        freqs=freqs,
        q_matrix=q,
        model="GeneralSymmetric",
        model_params={"exchangeabilities": exch, "weibull_shape": shapes},
    )


def _evolve_tips(post, T, N, S, bl, q, freqs, rng):
    """Simulate tip states down the tree under exp(Q t) (single rate)."""
    from scipy.linalg import expm  # test/bench helper only

    node_count = 2 * T - 1
    states = np.zeros((node_count, N), dtype=np.int64)
    root = post[-1, 0]
    states[root] = rng.choice(S, size=N, p=freqs)
    for node, left, right in post[::-1]:
        for child in (left, right):
            P = expm(q * bl[child])
            P = np.clip(P, 0, None)
            P /= P.sum(axis=1, keepdims=True)
            cdf = np.cumsum(P, axis=1)
            u = rng.random(N)
            states[child] = (u[:, None] > cdf[states[node]]).sum(axis=1).clip(0, S - 1)
    return states[:T]


def make_time_tree(postorder: np.ndarray, tip_count: int, draws: int, seed: int = 3,
                   dated_fraction: float = 0.6):
    """Heterochronous time-tree inputs on a given topology (BASELINE config 3 shape;
    torchtree/evolution/tree_model.py:380-424, tree_height_transform.py:36-56):

    times   [T]      sampling times of the tips (a fraction is 0 = contemporaneous)
    bounds  [T-1]    lower bound of every internal node = the oldest tip below it
    x       [D,T-1]  ratios in (0.1, 0.9) and, at the root's slot, the root height
    child / parent [2T-2]  node and parent-node index of every branch (branch b = node b)
    """
    rng = np.random.default_rng(seed)
    T = tip_count
    times = rng.uniform(0.0, 3.0, T) * (rng.random(T) < dated_fraction)
    node_bound = np.concatenate([times, np.zeros(T - 1)])
    parent = np.full(2 * T - 1, -1, dtype=np.int64)
    for node, left, right in postorder:
        node_bound[node] = max(node_bound[left], node_bound[right])
        parent[left] = parent[right] = node
    root = int(postorder[-1][0])
    x = rng.random((draws, T - 1)) * 0.8 + 0.1
    x[:, root - T] = node_bound[root] + 5.0 + 4.0 * rng.random(draws)
    child = np.array([n for n in range(2 * T - 1) if n != root], dtype=np.int64)
    return dict(times=times, bounds=node_bound[T:].copy(), x=x, child=child,
                parent=parent[child], root=root)
