#!/usr/bin/env python
"""Minimal driver for ncu: a few logL+gradient evaluations of one BASELINE config through the
C ABI (device-resident inputs, no graphs so that every kernel is a separate launch).

    ncu --set full ... python tools/profile_eval.py --config 5 --evals 2
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

import bench  # noqa: E402
from torchtree_b200 import Engine, reversible_eigensystem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--evals", type=int, default=2)
ap.add_argument("--patterns", type=int, default=None)
a = ap.parse_args()
cfg = dict(bench.CONFIGS[a.config], index=a.config, topology="random")
if a.patterns:
    cfg["patterns"] = a.patterns
prob = bench.build_problem(cfg)
dev = torch.device("cuda", 0)
eng = Engine(prob.tip_states, prob.weights, prob.postorder, cfg["states"], cfg["categories"],
             max_draws=cfg["draws"], device=0, flags=1 | 32)
evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix), torch.tensor(prob.freqs))
args = [torch.tensor(x).to(dev) for x in (prob.branch_lengths, prob.site_rates, prob.site_props)] + \
       [t.contiguous().to(dev) for t in (evec, ivec, evals)] + [torch.tensor(prob.freqs).to(dev)]
for _ in range(a.evals):
    eng.loglik_eigen(*args)
    eng.grad_eigen_packed()
torch.cuda.synchronize()
print("done", eng.launch_count)
