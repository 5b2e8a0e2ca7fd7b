#!/usr/bin/env python
"""Small logL + gradient evaluations of every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py

Three evaluations per problem (the codon path switches to the kept-u kernels from the second
evaluation on); 4, 20 and 61 states, tips with gaps, a draw batch for the 4-state path."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from torchtree_b200 import Engine, reversible_eigensystem  # noqa: E402
from torchtree_b200.synthetic import make_problem  # noqa: E402

for (T, N, S, K, D) in ((24, 300, 4, 4, 2), (14, 200, 20, 2, 1), (12, 160, 61, 2, 1)):
    prob = make_problem(T, N, S, K, draws=D, seed=S, gap_fraction=0.05)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, S, K, code_partials=prob.code_partials,
                 max_draws=D, flags=32)
    evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix), torch.tensor(prob.freqs))
    for _ in range(3):
        lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props, evec, ivec, evals,
                               prob.freqs)
        g = eng.grad_eigen()
    print(S, float(lnl.sum()), float(g["branch_lengths"].sum()))
    eng.close()
