"""Wall time of one logL + backward through the model-object glue (flatten.evaluate_models ->
autograd Function -> C ABI) on the fluA fixture, next to the bare engine call: how much of a
small evaluation is host-side Python.  Runs on a GPU box.  Lives under tests/ because the
stand-in model classes (tests/standins.py) borrow model formulas from oracle/.

    python tests/bench_plugin_overhead.py [fixture] [--profile]
"""
import cProfile
import json
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # standins, helpers

import standins as sm  # noqa: E402
from helpers import load_golden  # noqa: E402

from torchtree_b200 import Engine  # noqa: E402
from torchtree_b200.flatten import evaluate_models  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "fluA_gtr_w4_generic"
    prob, rec = load_golden(name)
    D = prob.draws
    blens = torch.tensor(rec["param_blens"], requires_grad=True)
    rates6 = torch.tensor(rec["param_gtr_rates"], requires_grad=True)
    freqs = torch.tensor(rec["param_gtr_freqs"], requires_grad=True)
    shape = torch.tensor(rec["param_shape"], requires_grad=True)
    tree = sm.UnRootedTreeModel(blens, prob.postorder)
    site = sm.WeibullSiteModel(shape, prob.category_count)
    subst = sm.GTR(rates6, freqs)
    sample_shape = torch.Size([D]) if D > 1 else torch.Size([])
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, prob.state_count,
                 prob.category_count, code_partials=prob.code_partials, max_draws=D)

    def step():
        for p in (blens, rates6, freqs, shape):
            p.grad = None
        lnl = evaluate_models(eng, tree, site, subst, None, sample_shape)
        lnl.sum().backward()
        return lnl

    def step_nograd():
        with torch.no_grad():
            return evaluate_models(eng, tree, site, subst, None, sample_shape)

    def best(fn, n=200):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3

    out = {"fixture": name, "draws": D, "logL_plus_backward_ms": round(best(step), 4),
           "logL_only_no_grad_ms": round(best(step_nograd), 4)}
    print(json.dumps(out))
    if "--profile" in sys.argv:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            step()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(30)


if __name__ == "__main__":
    main()
