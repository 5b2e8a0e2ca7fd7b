#!/usr/bin/env python
"""Device eigen-decomposition (ttb2_loglik_q, csrc/eigen.cu) vs the host route
(torch.linalg.eigh + ttb2_loglik_eigen): wall time per likelihood call through host tensors,
on small alignments where the decomposition is a visible share.  One JSON line per shape."""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from torchtree_b200 import Engine, reversible_eigensystem  # noqa: E402
from torchtree_b200.synthetic import make_problem  # noqa: E402


def bench(fn, n):
    for _ in range(5):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


for S, T, N, K, D in ((4, 69, 238, 4, 1), (4, 69, 238, 4, 128), (20, 50, 500, 4, 1),
                      (20, 50, 500, 4, 32), (61, 30, 200, 4, 1), (61, 30, 200, 4, 32)):
    prob = make_problem(T, N, S, K, draws=D, seed=1, per_draw_model=D > 1)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, S, K, max_draws=D)
    args = [torch.tensor(a) for a in (prob.branch_lengths, prob.site_rates, prob.site_props)]
    q, f = torch.tensor(prob.q_matrix), torch.tensor(prob.freqs)

    def host():
        e = reversible_eigensystem(q, f)
        return eng.loglik_eigen(*args, *e, f)

    def device():
        return eng.loglik_q(*args, q, f)

    a, b = host().clone(), device().clone()
    rel = float(((a - b).abs() / a.abs()).max())
    n = 200 if D == 1 else 50
    t_eigh = bench(lambda: reversible_eigensystem(q, f), n)
    line = {"states": S, "draws": D, "taxa": T, "patterns": N,
            "host_eigh_ms": t_eigh, "loglik_host_eigen_ms": bench(host, n),
            "loglik_device_eigen_ms": bench(device, n), "lnL_rel_diff": rel,
            "torch_threads": torch.get_num_threads()}
    print(json.dumps(line), flush=True)
    eng.close()
