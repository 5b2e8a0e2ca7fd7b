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
ap.add_argument("--trace", action="store_true",
                help="print the phase timeline of gm_bwd3_kernel (library built with -DTTB2_GM_TRACE)")
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

if a.trace:
    import ctypes

    import numpy as np

    from torchtree_b200 import _lib

    lib = _lib.load()
    buf = np.zeros(2 * 64 * 6, dtype=np.int64)
    rc = lib.ttb2_debug_gm_trace(buf.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    t = buf.reshape(2, 64, 6)
    t0 = t[t > 0].min()
    print("group trip  start    load      U  token      Q      G   (SM clocks; start relative to the first stamp)")
    for trip in range(12):
        for grp in range(2):
            r = t[grp, trip]
            if r[0] == 0:
                continue
            print("%5d %4d %7d %6d %6d %6d %6d %6d" % (grp, trip, r[0] - t0, r[1] - r[0], r[2] - r[1],
                                                      r[3] - r[2], r[4] - r[3], r[5] - r[4]))
