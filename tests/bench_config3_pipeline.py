"""Wall time of one config-3 ADVI-style step on the device side: 128 draws of
(ratios, root height, clock rate) -> node heights (ttb2_heights) -> branch lengths ->
engine logL for all draws -> backward to every input.  Runs on a GPU box.

    python tests/bench_config3_pipeline.py [taxa] [patterns] [draws]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import standins as sm  # noqa: E402
from test_config3_pipeline_gpu import _branch_lengths, _time_tree_problem  # noqa: E402

from torchtree_b200 import Engine  # noqa: E402
from torchtree_b200.flatten import evaluate_models  # noqa: E402
from torchtree_b200.height_transform import NodeHeightPlan, node_heights  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    prob, post, times, bounds, x, rate, child_idx, parent_idx = _time_tree_problem(T, N, D, 3)
    eng = Engine(prob.tip_states, prob.weights, post, 4, 1, code_partials=prob.code_partials,
                 max_draws=D, flags=1)
    plan = NodeHeightPlan(T, post, bounds.numpy())
    x.requires_grad_(True)
    rate.requires_grad_(True)
    parts = {}

    def step():
        x.grad = rate.grad = None
        t0 = time.perf_counter()
        h = node_heights(x, plan)
        bl = _branch_lengths(h, times, child_idx, parent_idx)
        t1 = time.perf_counter()
        lnl = evaluate_models(eng, sm.TimeTreeModel(bl, post), sm.ConstantSiteModel(), sm.JC69(),
                              sm.StrictClockModel(rate, 2 * T - 2), torch.Size([D]))
        t2 = time.perf_counter()
        lnl.mean().backward()
        t3 = time.perf_counter()
        parts.update(heights_ms=(t1 - t0) * 1e3, likelihood_ms=(t2 - t1) * 1e3,
                     backward_ms=(t3 - t2) * 1e3)
        return t3 - t0

    for _ in range(3):
        step()
    best = min(step() for _ in range(10))
    print(json.dumps({"taxa": T, "patterns": N, "draws": D, "step_ms": round(best * 1e3, 3),
                      "last_split_ms": {k: round(v, 3) for k, v in parts.items()},
                      "units_per_s": N * (T - 1) * D / best}))


if __name__ == "__main__":
    main()
