#!/usr/bin/env python
"""Run the other BASELINE.json configurations (3, 4, 5 shapes) on the GPU:
parity against the CPU oracle at a reduced pattern count, then logL+gradient
timing at the configuration's full size.  One JSON line per configuration.

    python tools/run_configs.py [--quick]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from torchtree_b200 import Engine, reversible_eigensystem  # noqa: E402
from torchtree_b200.synthetic import make_problem  # noqa: E402

PEAK_GBS = 6553.9


def eig(prob):
    return reversible_eigensystem(torch.tensor(prob.q_matrix), torch.tensor(prob.freqs))


def evaluate(eng, prob, e, need_q=True):
    lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props, *e, prob.freqs)
    g = eng.grad_eigen()
    return lnl, g


def parity(cfg):
    from oracle import treelik as orc

    prob = make_problem(cfg["T"], cfg["N_parity"], cfg["S"], cfg["K"], draws=cfg["D_parity"],
                        seed=cfg["seed"], per_draw_model=False)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, cfg["S"], cfg["K"],
                 max_draws=prob.draws)
    lnl, g = evaluate(eng, prob, eig(prob))
    want = orc.evaluate(prob, want_grad=True, through_q=True)
    rel = float(np.max(np.abs(lnl.numpy() - want["lnL"]) / np.abs(want["lnL"])))
    gb = g["branch_lengths"].numpy()
    gerr = float(np.max(np.abs(gb - want["branch_lengths"]) /
                        np.maximum(np.abs(want["branch_lengths"]), 1e-8 * np.abs(want["branch_lengths"]).max())))
    eng.close()
    return rel, gerr


def timing(cfg, steps):
    prob = make_problem(cfg["T"], cfg["N"], cfg["S"], cfg["K"], draws=cfg["D"], seed=cfg["seed"])
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, cfg["S"], cfg["K"],
                 max_draws=prob.draws, flags=1)
    dev = torch.device("cuda", 0)
    e = eig(prob)
    args = [torch.tensor(a).to(dev) if not isinstance(a, torch.Tensor) else a.to(dev)
            for a in (prob.branch_lengths, prob.site_rates, prob.site_props, *e, prob.freqs)]
    eng.set_stream(torch.cuda.current_stream().cuda_stream)

    def step():
        eng.loglik_eigen(*args)
        eng.grad_eigen()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.enable_timing(True)
    step()
    phases = eng.phase_ms()
    eng.enable_timing(False)
    bytes_per_unit = 5 * cfg["S"] * 8
    out = {
        "ms_per_eval": ms, "units": prob.units, "units_per_s": prob.units / (ms * 1e-3),
        "evals_per_s": 1e3 / ms,
        "hbm_frac_of_measured": prob.units * bytes_per_unit / (ms * 1e-3) / 1e9 / PEAK_GBS,
        "device_GB": eng.device_bytes / 1e9,
        "phases_ms": {k: round(v, 3) for k, v in phases.items()},
    }
    eng.close()
    return out


def config1_flu():
    """Config 1 shape on the real fluA data (golden fixture): 69 taxa, 238 patterns,
    GTR + 4 categories, unrooted; parity against the reference's own outputs."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from helpers import load_golden

    prob, rec = load_golden("fluA_gtr_w4_generic")
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 4,
                 code_partials=prob.code_partials, max_draws=1, flags=1)
    e = eig(prob)
    lnl, g = evaluate(eng, prob, e)
    rel = abs(lnl.item() - rec["lnL"][0]) / abs(rec["lnL"][0])
    gerr = float(np.max(np.abs(g["branch_lengths"].numpy() - rec["d_branch_lengths"]) /
                        np.maximum(np.abs(rec["d_branch_lengths"]), 1e-8 * np.abs(rec["d_branch_lengths"]).max())))
    host = [torch.tensor(a) if not isinstance(a, torch.Tensor) else a
            for a in (prob.branch_lengths, prob.site_rates, prob.site_props, *e, prob.freqs)]
    for _ in range(20):
        eng.loglik_eigen(*host); eng.grad_eigen()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        eng.loglik_eigen(*host); eng.grad_eigen()
    ms = (time.perf_counter() - t0) / n * 1e3
    line = {"config": "config1_fluA_gtr_w4", "note": "real fluA data (69 taxa, 238 patterns), host "
            "tensors in/out through the C ABI (launch-bound: %d kernel launches per evaluation)"
            % (eng.launch_count // 221),
            "shape": {"T": 69, "N": 238, "S": 4, "K": 4, "D": 1},
            "parity": {"lnL_rel_err_vs_reference": rel, "branch_grad_max_rel_err": gerr,
                       "ok": bool(rel <= 1e-10 and gerr <= 1e-8)},
            "timing": {"ms_per_eval_e2e": ms, "evals_per_s": 1e3 / ms,
                       "reference_cpu_ms_per_eval": 25.0,
                       "reference_note": "SURVEY section 6: 0.025 s logL+grad on 8 CPU threads"}}
    eng.close()
    print(json.dumps(line), flush=True)


CONFIGS = {
    "config3_jc69_clock_D128": dict(T=500, N=10_000, S=4, K=1, D=128, D_parity=3, N_parity=300,
                                    seed=3, note="500-taxon tree, K=1, 128 draws per step "
                                    "(draw batch; tree prior and height transform stay in torch)"),
    "config4_aa_LG_like": dict(T=200, N=50_000, S=20, K=4, D=1, D_parity=1, N_parity=200, seed=4,
                               note="20-state reversible model + 4 categories (warp-autonomous DMMA kernels, "
                               "csrc/kernels_gwarp.cu)"),
    "config5_codon_61": dict(T=100, N=20_000, S=61, K=4, D=1, D_parity=1, N_parity=64, seed=5,
                             note="61-state reversible model + 4 categories (DMMA tile kernels, "
                             "csrc/kernels_gmma.cu)"),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    if not a.only or a.only == "config1":
        config1_flu()
    for name, cfg in CONFIGS.items():
        if a.only and a.only != name:
            continue
        t0 = time.time()
        rel, gerr = parity(cfg)
        line = {"config": name, "note": cfg["note"],
                "shape": {k: cfg[k] for k in ("T", "N", "S", "K", "D")},
                "parity": {"patterns": cfg["N_parity"], "lnL_rel_err": rel,
                           "branch_grad_max_rel_err": gerr,
                           "ok": bool(rel <= 1e-10 and gerr <= 1e-8)}}
        if a.quick:
            cfg = dict(cfg, N=max(1000, cfg["N"] // 10))
            line["shape"]["N"] = cfg["N"]
        line["timing"] = timing(cfg, 3)
        line["wall_s"] = time.time() - t0
        print(json.dumps(line), flush=True)
