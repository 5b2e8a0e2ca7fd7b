#!/usr/bin/env python
"""Benchmark of the tree-likelihood hot path (logL + full gradient), BASELINE.json configs.

    python bench.py [--config 2] --gpus N --steps K --warmup W       # CUDA engine
    python bench.py --impl reference [--config 2] --gpus N ...        # the reference on the host cores

Default workload = BASELINE.json configs[1] ("config 2"): synthetic 1,000 taxa x 100,000 site
patterns, GTR + 4 Weibull rate categories, unrooted, one draw; a "step" is one logL + gradient
evaluation.  With N > 1 (torchrun, one process per GPU) the site patterns are sharded across the
ranks (strong scaling: the problem is fixed) and the packed {lnL, gradient} vector is all-reduced
over NCCL.  `--config 3|4|5` select the other BASELINE shapes (config 3 shards its batch of draws);
the default line also carries them as `other_configs`, measured in the same run at N = 1.

Metric: patterns x internal-nodes x categories (x draws) processed per second, fp64.
`value` has the inputs resident in HBM; `e2e` goes through the public API (the torch extension's
autograd Function, `sharded_log_likelihood` around it for N > 1) with pinned host tensors in and
host gradients out.  `--impl reference` times the UNMODIFIED reference (baseline/_ref, vendored by
tools/vendor_reference.py) on the same problem object, full size, on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "fp64 logL+grad patterns*nodes*cats/s"
UNIT = "patterns*nodes*cats/s"
FP64_TENSOR_PEAK_TFLOPS = 37.1  # DMMA, measured on the B200 box: profiles/r01_fp64_peak.jsonl

# BASELINE.json configs (index = BASELINE.json's 1-based position).  `n_parity` = patterns of the
# sample on which the engine is checked against the CPU oracle inside the run.
CONFIGS = {
    2: dict(name="config2_gtr_w4", taxa=1000, patterns=100_000, states=4, categories=4, draws=1,
            shard="patterns", seed=20260101, n_parity=8000,
            workload="BASELINE.json configs[1]: synthetic 1000 taxa x 100000 site patterns, GTR + 4 "
                     "Weibull rate categories, unrooted, one logL+gradient evaluation per step"),
    3: dict(name="config3_jc69_clock_D128", taxa=500, patterns=10_000, states=4, categories=1,
            draws=128, shard="draws", seed=3, n_parity=300,
            workload="BASELINE.json configs[2]: synthetic 500-taxon time tree x 10000 site patterns, "
                     "JC69 strict clock + constant coalescent, 128 variational draws per step"),
    4: dict(name="config4_aa_lg_w4", taxa=200, patterns=50_000, states=20, categories=4, draws=1,
            shard="patterns", seed=4, n_parity=200,
            workload="BASELINE.json configs[3]: synthetic 200 taxa x 50000 patterns, 20-state reversible "
                     "(LG-like) model + 4 rate categories, one logL+gradient evaluation per step"),
    5: dict(name="config5_codon61_w4", taxa=100, patterns=20_000, states=61, categories=4, draws=1,
            shard="patterns", seed=5, n_parity=64,
            workload="BASELINE.json configs[4]: synthetic 100 taxa x 20000 codon patterns, 61-state "
                     "reversible model + 4 rate categories, one logL+gradient evaluation per step"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--patterns", type=int, default=None)
    ap.add_argument("--categories", type=int, default=None)
    ap.add_argument("--draws", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--cpu-patterns", type=int, default=None,
                    help="patterns of the bounded CPU-baseline sample (default: 25000 for config 2)")
    ap.add_argument("--topology", default="random", choices=["random", "caterpillar", "balanced"],
                    help="tree shape (the headline workload is the random-join tree)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=600.0,
                    help="--impl reference: wall-clock budget for warm-up + timed steps")
    ap.add_argument("--engine-flags", type=int, default=0,
                    help="extra TTB2_FLAG_* bits for experiments (32 = no CUDA graphs)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    for key in ("taxa", "patterns", "categories", "draws", "seed"):
        if getattr(args, key) is not None:
            cfg[key] = getattr(args, key)
    cfg["topology"] = args.topology
    cfg["index"] = args.config
    args.cfg = cfg
    return args


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region: NVML in a thread (cheap
    queries; no subprocess competing for the driver lock), nvidia-smi as the fall-back."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
               ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period=0.1):
        self.index, self.period = index, period
        self.mhz, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.thread = None
        self.how = None

    def _nvml_loop(self, nv, handle):
        while not self._stop.is_set():
            try:
                self.mhz.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", str(int(self.period * 1000))],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        names = [n for n, _ in self.REASONS]
        while not self._stop.is_set():
            line = proc.stdout.readline()
            if not line:
                break
            r = [x.strip() for x in line.split(",")]
            try:
                self.mhz.append(float(r[0]))
                self.mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
        proc.terminate()

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.index
            if visible:
                try:
                    index = int(visible.split(",")[self.index])
                except Exception:
                    pass
            handle = nv.nvmlDeviceGetHandleByIndex(index)
            nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
            self.how = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
        except Exception:
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler unavailable"]}
        time.sleep(0.12)
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.mhz)) if self.mhz else None,
                "sm_max_mhz": float(max(self.mx)) if self.mx else None,
                "samples": len(self.mhz), "reasons": sorted(self.reasons), "how": self.how}


# ------------------------------------------------------------------------------------------
# problems
# ------------------------------------------------------------------------------------------
def build_problem(cfg, lo=0, hi=None, patterns=None, draws=None):
    from torchtree_b200.synthetic import make_problem

    D = draws if draws is not None else cfg["draws"]
    prob = make_problem(cfg["taxa"], patterns or cfg["patterns"], cfg["states"], cfg["categories"],
                        draws=D, seed=cfg["seed"], topology=cfg.get("topology", "random"))
    if cfg["index"] == 3:  # JC69: equal frequencies and exchangeabilities
        S = cfg["states"]
        prob.freqs = np.full((1, S), 1.0 / S)
        q = np.full((S, S), 1.0 / (S - 1))
        np.fill_diagonal(q, -1.0)
        prob.q_matrix = q[None]
        prob.model = "JC69"
    if hi is not None:
        prob.tip_states = np.ascontiguousarray(prob.tip_states[:, lo:hi])
        prob.weights = np.ascontiguousarray(prob.weights[lo:hi])
        prob.pattern_count = hi - lo
    return prob


def workload_config(cfg, world):
    per_gpu_patterns = cfg["patterns"] if cfg["shard"] == "draws" else \
        (cfg["patterns"] + world - 1) // world
    per_gpu_draws = (cfg["draws"] + world - 1) // world if cfg["shard"] == "draws" else cfg["draws"]
    gb = 2 * per_gpu_patterns * (cfg["taxa"] - 1) * cfg["categories"] * cfg["states"] * 8 * \
        per_gpu_draws / 1e9
    return {
        "workload": cfg["workload"], "taxa": cfg["taxa"], "patterns": cfg["patterns"],
        "states": cfg["states"], "categories": cfg["categories"], "draws": cfg["draws"],
        "topology": cfg.get("topology", "random"),
        "sharding": "%s/%d" % (cfg["shard"], world),
        "l2": "working set (%.1f GB of conditional-likelihood vectors per GPU) >> 126 MB L2; no "
              "explicit flush" % gb,
    }


# ------------------------------------------------------------------------------------------
# what one evaluation must move / compute, counted from the topology (per pattern, category, draw)
# ------------------------------------------------------------------------------------------
def traffic_model(postorder, T, S, cherries: bool, kept_u: bool = False):
    """Bytes per (pattern, category) that the engine's algorithm has to move through HBM in the
    two sweeps, and the fp64 flops of its GEMM-shaped work, from the tree alone.

    One conditional-likelihood vector = S doubles.  Post-order: every stored node is written once
    and read once by its parent; tips are 1-byte codes; with cherry tabulation (4-state path) a
    node whose two children are tips is a table entry, never stored.  Pre-order: a parent reads
    its own q^ and the stored vectors of its internal children and writes q^ of every internal
    child (a tabulated cherry still receives q^).  kept_u (the 61-state path from an engine's
    second evaluation on): the post-order sweep also writes u_c = P_c p~_c of every internal child
    and the pre-order sweep reads it back instead of repeating the product."""
    V = S * 8
    post = np.asarray(postorder)
    root = int(post[-1][0])
    is_cherry = {int(n) for n, l, r in post if l < T and r < T and cherries and int(n) != root}
    stored = lambda n: n >= T and n not in is_cherry  # noqa: E731
    post_bytes = pre_bytes = 0
    flops_post = flops_pre = 0
    for n, l, r in post:
        n, l, r = int(n), int(l), int(r)
        if n not in is_cherry:
            post_bytes += V                                  # write p~_n
            post_bytes += V * (stored(l) + stored(r))        # read stored children
        pre_bytes += V                                       # read q^_n
        pre_bytes += V * (stored(l) + stored(r))             # read stored children
        pre_bytes += V * ((l >= T) + (r >= T))               # write q^ of internal children
        # GEMM-shaped work: u = P p~ per internal child (tips are table look-ups) in both sweeps,
        # q^_c = P^T m per internal child, G_c += m (x) p~ for both children
        inner = (l >= T) + (r >= T)
        flops_post += 2 * S * S * inner
        flops_pre += 2 * S * S * inner * (1 if kept_u else 2) + 2 * S * S * 2
        if kept_u:
            post_bytes += V * inner
            pre_bytes += V * inner
    I = len(post)
    return {"post_bytes": post_bytes / I, "pre_bytes": pre_bytes / I,
            "flops_post": flops_post / I, "flops_pre": flops_pre / I,
            "cherry_nodes": len(is_cherry)}


def ncu_traffic_for_build():
    """DRAM bytes of the pre-order sweep from the ncu capture of THIS build, if one is committed
    (profiles/r02_ncu_traffic.json is keyed by the library's source digest)."""
    try:
        with open(os.path.join(REPO, "profiles", "r02_ncu_traffic.json")) as fp:
            rec = json.load(fp)
        with open(os.path.join(REPO, "torchtree_b200", "lib", "libttb200.stamp")) as fp:
            stamp = fp.read().strip()
        return rec if rec.get("lib_stamp") == stamp else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# the engine arm
# ------------------------------------------------------------------------------------------
def gtr_generator(rates, freqs):
    """Normalised reversible generator from exchangeabilities (row-major upper triangle) and
    frequencies, in torch (differentiable): what GTR.q / GeneralSymmetric.q + the normalisation
    of SymmetricSubstitutionModel.p_t do (nucleotide.py:328-374, abstract.py:49-59)."""
    S = freqs.shape[-1]
    iu = torch.triu_indices(S, S, 1)
    R = torch.zeros((S, S), dtype=rates.dtype)
    R = R.index_put((iu[0], iu[1]), rates)
    R = R + R.transpose(-1, -2)
    Q = R * freqs.unsqueeze(-2)
    Q = Q - torch.diag_embed(Q.sum(-1))
    norm = -(torch.diagonal(Q) * freqs).sum()
    return Q / norm


def weibull_rates(shape, K):
    """site_model.py:173-195, :237-247 in torch (differentiable w.r.t. the shape)."""
    quant = (2.0 * torch.arange(K, dtype=shape.dtype) + 1.0) / (2.0 * K)
    r = torch.pow(-torch.log(1.0 - quant), 1.0 / shape)
    props = torch.full((K,), 1.0 / K, dtype=shape.dtype)
    return r / (r * props).sum(-1, keepdim=True), props


def parity_vs_reference(cfg, patterns, device):
    """Parity gate against the REAL reference on a bounded sample (same tree and model, fewer
    patterns) and, in the same breath, the CPU baseline timing: the reference's own objects
    (oracle/reference_arm.py) vs the engine through the product API with the model parameters as
    autograd leaves.  lnL 1e-10, branch / Weibull-shape gradients 1e-8, GTR parameters 1e-7 (the
    reference's eigh backward)."""
    from oracle import reference_arm as ra
    from torchtree_b200 import Engine, log_likelihood_eigen

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    prob = build_problem(cfg, patterns=patterns)
    rp = ra.ReferenceProblem(prob)
    ref = rp.evaluate()  # warm-up, and the values compared below
    t0 = time.perf_counter()
    rp.evaluate()
    best = time.perf_counter() - t0

    K, S = cfg["categories"], cfg["states"]
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, S, K, max_draws=1, device=device)
    bl = torch.tensor(prob.branch_lengths[0, :-1], requires_grad=True)
    shape = torch.tensor([float(prob.model_params["weibull_shape"][0])], requires_grad=True)
    rates6 = torch.tensor(prob.model_params["exchangeabilities"][0], requires_grad=True)
    freqs = torch.tensor(prob.freqs[0], requires_grad=True)
    site_rates, props = weibull_rates(shape, K)
    bls = torch.cat((bl, torch.zeros(1, dtype=bl.dtype)))
    lnl = log_likelihood_eigen(eng, bls[None], site_rates[None], props[None],
                               gtr_generator(rates6, freqs)[None], freqs[None])
    lnl.sum().backward()
    eng.close()

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    parity = {
        "against": "real reference (baseline/_ref), %d-pattern sample" % patterns,
        "lnL_rel_err": rel(lnl.detach().numpy(), ref["lnL"]),
        "branch_grad_rel_err": rel(bl.grad.numpy(), ref["branch_lengths"][0, :-1]),
        "weibull_shape_grad_rel_err": rel(shape.grad.numpy(), ref["weibull_shape"]),
        "gtr_rates_grad_rel_err": rel(rates6.grad.numpy(), ref["gtr_rates"]),
        "gtr_freqs_grad_rel_err": rel(freqs.grad.numpy(), ref["gtr_freqs"]),
    }
    parity["ok"] = bool(parity["lnL_rel_err"] <= 1e-10 and parity["branch_grad_rel_err"] <= 1e-8
                        and parity["weibull_shape_grad_rel_err"] <= 1e-8
                        and parity["gtr_rates_grad_rel_err"] <= 1e-7
                        and parity["gtr_freqs_grad_rel_err"] <= 1e-7)
    return {
        "parity_on_sample": parity, "value": prob.units / best, "unit": UNIT, "cores": threads,
        "kind": "reference",
        "sample": "%d taxa x %d patterns x K=%d (bounded sample of the %d-pattern workload): the "
                  "vendored reference's GTR.p_t + WeibullSiteModel + "
                  "calculate_treelikelihood_discrete_rescaled + .backward(), second of 2 evaluations "
                  "(%.2f s/eval); `bench.py --impl reference` times the full size"
                  % (cfg["taxa"], patterns, K, cfg["patterns"], best),
    }


def parity_vs_oracle(cfg, device):
    """Parity of one configuration against the CPU oracle port on a small sample (seconds)."""
    from oracle import treelik as orc
    from torchtree_b200 import Engine

    D = min(cfg["draws"], 3)
    prob = build_problem(cfg, patterns=cfg["n_parity"], draws=D)
    t0 = time.perf_counter()
    want = orc.evaluate(prob, want_grad=True, through_q=True)
    cpu_s = time.perf_counter() - t0
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, cfg["states"], cfg["categories"],
                 max_draws=D, device=device)
    lnl = eng.loglik_q(prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix,
                       prob.freqs)
    g = eng.grad_eigen()
    eng.close()

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    out = {"against": "oracle port (oracle/treelik.py), %d patterns x %d draws" % (cfg["n_parity"], D),
           "lnL_rel_err": rel(lnl.numpy(), want["lnL"]),
           "branch_grad_rel_err": rel(g["branch_lengths"].numpy(), want["branch_lengths"]),
           "site_rate_grad_rel_err": rel(g["site_rates"].numpy(), want["site_rates"]),
           "oracle_units_per_s": prob.units / cpu_s}
    out["ok"] = bool(out["lnL_rel_err"] <= 1e-10 and out["branch_grad_rel_err"] <= 1e-8
                     and out["site_rate_grad_rel_err"] <= 1e-8)
    return out


def run_engine(cfg, args, rank, local_rank, world, dist, with_clocks=True):
    """Times one configuration on this process group; returns the JSON fields (rank 0) or None."""
    from torchtree_b200 import Engine, log_likelihood_eigen
    from torchtree_b200.sharded import (draw_sharded_log_likelihood, shard_range,
                                        sharded_engine_log_likelihood)

    dev = torch.device("cuda", local_rank)
    T, N, S, K, D = cfg["taxa"], cfg["patterns"], cfg["states"], cfg["categories"], cfg["draws"]
    by_draws = cfg["shard"] == "draws"
    if by_draws:
        dlo, dhi = shard_range(D, rank, world)
        prob = build_problem(cfg)
        local_draws = dhi - dlo
    else:
        lo, hi = shard_range(N, rank, world)
        prob = build_problem(cfg, lo, hi)
        dlo, dhi, local_draws = 0, D, D
    units_total = N * (T - 1) * K * D
    units_rank = prob.pattern_count * (T - 1) * K * local_draws
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, S, K,
                 max_draws=max(1, local_draws), device=local_rank, flags=1 | args.engine_flags)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    # ---- inputs: full-size host tensors (what a user holds), this rank's slice on the device ----
    full = [torch.tensor(prob.branch_lengths), torch.tensor(prob.site_rates),
            torch.tensor(prob.site_props), torch.tensor(prob.q_matrix), torch.tensor(prob.freqs)]
    local = [t[dlo:dhi] if (by_draws and t.shape[0] == D and D > 1) else t for t in full]
    devin = [t.contiguous().to(dev) for t in local]
    lnl_d = torch.empty(max(1, local_draws), dtype=torch.float64, device=dev)
    if S > 8:
        # large state spaces: the eigen-system of the (single) generator is decomposed once on the
        # host and cached, as the torch extension does (csrc/torch_ext.cpp device_eigh)
        from torchtree_b200 import reversible_eigensystem

        evec, ivec, evals = reversible_eigensystem(local[3], local[4])
        devin = devin[:3] + [t.contiguous().to(dev) for t in (evec, ivec, evals)] + devin[4:]
        forward = eng.loglik_eigen
    else:
        forward = eng.loglik_q
    forward(*devin, out=lnl_d)
    n_local = eng.grad_eigen_packed().numel()
    # pattern sharding sums the packed vectors; draw sharding exchanges them (every rank's block
    # lands in its own slot of a zero-initialised vector, so the sum all-reduce is an all-gather)
    if by_draws and world > 1:
        block = torch.tensor([n_local], device=dev)
        dist.all_reduce(block, op=dist.ReduceOp.MAX)
        block = int(block.item())
        packed = torch.zeros(block * world, dtype=torch.float64, device=dev)
        my = packed[rank * block:rank * block + n_local]
    else:
        packed = torch.zeros(n_local, dtype=torch.float64, device=dev)
        my = packed

    def step_device():
        forward(*devin, out=lnl_d)
        eng.grad_eigen_packed(out=my)
        if world > 1:
            dist.all_reduce(packed)

    # ---- the call a user makes: the differentiable op of the torch extension, pinned host
    # tensors in, .backward() hands the gradients back as host tensors, lnL read on the host;
    # N > 1: the same op inside sharded_log_likelihood / draw_sharded_log_likelihood ----
    user_in = [t.clone().pin_memory().requires_grad_(True) for t in full]
    group = None

    def step_e2e():
        for t in user_in:
            t.grad = None
        fn = lambda *a: log_likelihood_eigen(eng, *a)  # noqa: E731
        if world == 1:
            lnl = fn(*user_in)
        elif by_draws:
            lnl = draw_sharded_log_likelihood(fn, user_in, D, group)
        else:   # what TreeLikelihoodModel(shard="patterns") calls (flatten.evaluate_models)
            lnl = sharded_engine_log_likelihood(eng, user_in, group)
        lnl.sum().backward()
        return lnl.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t[0].item(), t[1].item() / 1e3
        return ms / steps, wall / steps

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0 and with_clocks:
        sampler.start()
    launches0 = eng.launch_count
    ms_dev, _ = timed(step_device, args.steps)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None

    # ---- phase timing for the roofline of the dominant kernel family ----
    eng.enable_timing(True)
    pre_ms, post_ms, ph = [], [], {}
    for _ in range(min(5, args.steps)):
        step_device()
        ph = eng.phase_ms()
        pre_ms.append(ph["preorder"])
        post_ms.append(ph["postorder"])
    eng.enable_timing(False)
    ph_pre, ph_post = float(np.mean(pre_ms)), float(np.mean(post_ms))

    # ---- end to end through the public API ----
    for _ in range(3):
        step_e2e()
    _, wall_e2e = timed(step_e2e, args.steps)
    lnl_value = step_e2e().reshape(-1)
    lnl_value = float(lnl_value.sum()) if D > 1 else float(lnl_value[0])

    line = None
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        kept_u = S == 61 and not os.environ.get("TTB2_GM_NO_USTORE")
        tm = traffic_model(prob.postorder, T, S, cherries=(S == 4), kept_u=kept_u)
        per = prob.pattern_count * K * local_draws * (T - 1)   # (pattern, node, cat, draw) units of this rank
        V = S * 8
        moved_pre, moved_post = per * tm["pre_bytes"], per * tm["post_bytes"]
        alg_pre, alg_post = per * 3 * V, per * 2 * V            # 3 + 2 vectors per unit
        fl_pre, fl_post = per * tm["flops_pre"], per * tm["flops_post"]
        launches_pre = max(1, ph.get("preorder_launches", 1))
        tensor_bound = S > 32
        names = {4: ("bwd4_tma_kernel<3,5> (pre-order sweep, one launch per tree level; level 1 is "
                     "bwd4_tips_tma_kernel<4>)", "fwd4c_kernel<K> (post-order sweep; cherries tabulated)"),
                 20: ("gw_bwd_kernel<20,4,2> (warp-autonomous DMMA pre-order sweep; level 1 is "
                      "gm_cherry_bwd_kernel<20,6>)",
                      "gw_fwd_kernel<20,8,3> (post-order sweep; level 1 is gm_cherry_fwd_kernel<20>)"),
                 61: ("gm_bwd3_kernel<61> (two-group DMMA pre-order sweep; level 1 is "
                      "gm_cherry_bwd_kernel<61,8>)",
                      "gm_fwd3_kernel<61> (warp-specialised DMMA post-order sweep; level 1 is "
                      "gm_cherry_fwd_kernel<61>)")}
        kpre, kpost = names.get(S, ("pre-order sweep", "post-order sweep"))
        ncu = ncu_traffic_for_build() if (cfg["index"] == 2 and world == 1 and N == 100_000
                                          and cfg.get("topology") == "random") else None
        if tensor_bound:
            ach = fl_pre / (ph_pre * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": kpre, "achieved": ach,
                    "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": ach / FP64_TENSOR_PEAK_TFLOPS,
                    "peak_kind": "fp64 DMMA peak measured on this fleet (profiles/r01_fp64_peak.jsonl, "
                                 "tools/fp64_peak.cu); MEASURED_PEAKS.json holds no fp64 figure",
                    "traffic": None,
                    "flops_per_launch": fl_pre / launches_pre,
                    "flops_note": "fp64 flops of the GEMM-shaped work the sweep executes, counted from "
                                  "the tree (unpadded, 61 not 64): q^ = P^T m per internal child, G += "
                                  "m (x) p~ per child" + (
                                      "; u = P p~ is read back from the post-order sweep (kept u), "
                                      "not recomputed and not counted" if kept_u else
                                      ", u = P p~ per internal child") +
                                  " (tip children are table look-ups)"}
        else:
            ach = moved_pre / (ph_pre * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": kpre, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_kind": peak_kind,
                    "traffic": (ncu["preorder_bytes_per_step"] / launches_pre) if ncu else None,
                    "traffic_note": ("ncu dram__bytes_read+write of the pre-order sweep of this build "
                                     "(profiles/r02_ncu_traffic.json), per launch") if ncu else
                                    "no ncu capture committed for this build",
                    "bytes_per_launch": moved_pre / launches_pre,
                    "bytes_note": "bytes the sweep has to move, counted from the tree (cherry vectors "
                                  "are tabulated, tips are 1-byte codes): %.1f B per unit instead of "
                                  "the %d B of SURVEY 8(d)" % (tm["pre_bytes"], 3 * V),
                    "algorithmic": {"bytes_per_unit": 3 * V,
                                    "achieved": alg_pre / (ph_pre * 1e-3) / 1e9,
                                    "frac": alg_pre / (ph_pre * 1e-3) / 1e9 / peak,
                                    "note": "SURVEY 8(d) figure (3 vectors per unit) / time: counts "
                                            "bytes the cherry tabulation never moves"}}
            if S >= 8:
                roof["tensor_pipe"] = {"achieved_tflops": fl_pre / (ph_pre * 1e-3) / 1e12,
                                       "frac_of_dmma_peak": fl_pre / (ph_pre * 1e-3) / 1e12
                                       / FP64_TENSOR_PEAK_TFLOPS}
        roof.update({"launches_per_step": launches_pre, "avg_launch_ms": ph_pre / launches_pre,
                     "sweep_ms": ph_pre})
        roof["postorder"] = {"kernel": kpost, "ms": ph_post,
                             "launches_per_step": ph.get("postorder_launches"),
                             "achieved_gbs": moved_post / (ph_post * 1e-3) / 1e9,
                             "frac_hbm": moved_post / (ph_post * 1e-3) / 1e9 / peak,
                             "bytes_per_unit": tm["post_bytes"],
                             "algorithmic_frac": alg_post / (ph_post * 1e-3) / 1e9 / peak}
        roof["whole_step"] = {
            "moved_gbs": (moved_pre + moved_post) / (ms_dev * 1e-3) / 1e9,
            "frac_hbm": (moved_pre + moved_post) / (ms_dev * 1e-3) / 1e9 / peak,
            "algorithmic_frac": (alg_pre + alg_post) / (ms_dev * 1e-3) / 1e9 / peak,
            "tflops": (fl_pre + fl_post) / (ms_dev * 1e-3) / 1e12}
        h2d = sum(t.numel() * 8 for t in user_in)
        line = {
            "metric": METRIC, "value": units_total / (ms_dev * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(cfg, world),
            "evals_per_s": 1e3 / ms_dev, "lnL": lnl_value,
            "e2e": {"value": units_total / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": h2d + 8 * D, "ms_per_step": wall_e2e * 1e3,
                    "api": "torchtree_b200.log_likelihood_eigen(...).backward() (torch C++ extension "
                           "autograd Function -> ttb2_loglik_q / ttb2_grad_eigen_packed), pinned host "
                           "tensors in, host gradients out" + (
                               "" if world == 1 else
                               ", inside torchtree_b200.sharded.%s (NCCL all-reduce of lnL and of the "
                               "packed gradient)" % ("draw_sharded_log_likelihood" if by_draws
                                                     else "sharded_engine_log_likelihood"))},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "phases_ms": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in ph.items()},
            "device_bytes": eng.device_bytes,
            "units_per_rank": units_rank,
        }
    eng.close()
    return line


def config3_pipeline(cfg, device):
    """BASELINE config 3 as one device pipeline (SURVEY 8(f) f1 + f2): ratios / root height ->
    node heights (ttb2_heights) -> branch lengths x clock rate -> engine lnL for all draws ->
    + constant coalescent (ttb2_coalescent_constant), mean over draws, backward to every input.
    Host tensors in, host gradients out; one evaluation = one ADVI gradient step's model part."""
    from torchtree_b200 import Engine, constant_coalescent_log_prob, log_likelihood_eigen
    from torchtree_b200.height_transform import NodeHeightPlan, node_heights
    from torchtree_b200.synthetic import make_time_tree

    T, N, D = cfg["taxa"], cfg["patterns"], cfg["draws"]
    prob = build_problem(cfg)
    tt = make_time_tree(prob.postorder, T, D, seed=cfg["seed"])
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 1, max_draws=D, device=device,
                 flags=1)
    plan = NodeHeightPlan(T, prob.postorder, tt["bounds"], device=device)
    times = torch.tensor(tt["times"])
    child, parent = torch.tensor(tt["child"]), torch.tensor(tt["parent"])
    x = torch.tensor(tt["x"], requires_grad=True)
    rate = (0.02 * (1.0 + 0.1 * torch.rand(D, 1, dtype=torch.float64))).requires_grad_(True)
    theta = torch.tensor([[4.0]], dtype=torch.float64, requires_grad=True)
    q, freqs = torch.tensor(prob.q_matrix), torch.tensor(prob.freqs)
    ones = torch.ones(1, 1, dtype=torch.float64)
    parts = {}

    def step():
        x.grad = rate.grad = theta.grad = None
        t0 = time.perf_counter()
        h = node_heights(x, plan)                                   # [D, T-1]
        allh = torch.cat((times.expand(D, -1), h), -1)               # [D, 2T-1]
        bl = (allh[:, parent] - allh[:, child]) * rate               # tree_model.py:407-424 x clock
        t1 = time.perf_counter()
        lnl = log_likelihood_eigen(eng, bl, ones, ones, q, freqs)    # [D]
        t2 = time.perf_counter()
        coal = constant_coalescent_log_prob(allh, theta, device).reshape(-1)
        t3 = time.perf_counter()
        (lnl + coal).mean().backward()
        t4 = time.perf_counter()
        parts.update(heights_branch_ms=(t1 - t0) * 1e3, likelihood_ms=(t2 - t1) * 1e3,
                     coalescent_ms=(t3 - t2) * 1e3, backward_ms=(t4 - t3) * 1e3)
        return t4 - t0

    for _ in range(3):
        step()
    best = min(step() for _ in range(8))
    eng.close()
    plan.close()
    return {"step_ms": best * 1e3, "units_per_s": N * (T - 1) * D / best,
            "split_ms_last": {k: round(v, 3) for k, v in parts.items()},
            "what": "host ratios/root height/clock rate/theta -> ttb2_heights -> branch lengths -> "
                    "engine lnL (128 draws) + ttb2_coalescent_constant -> backward to all inputs"}


# ------------------------------------------------------------------------------------------
# the reference arm
# ------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference's own CPU implementation of the path, all host threads, on the arm's config:
    the vendored reference (baseline/_ref) at the full pattern count when the host's memory holds
    its autograd tape, otherwise the largest power-of-two fraction that fits (said in `sample`);
    the oracle port only if baseline/_ref is absent."""
    if rank != 0:
        return
    cfg = args.cfg
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import reference_arm as ra

    patterns = args.cpu_patterns or cfg["patterns"]
    kind = "reference" if (ra.available() and cfg["draws"] == 1) else "port"
    note = ""
    if kind == "reference":
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = 64e9
        per_pattern = ra.estimated_tape_bytes(build_problem(cfg, patterns=64)) / 64
        while patterns > 1000 and per_pattern * patterns * 1.15 >= avail:
            patterns //= 2
        if patterns != cfg["patterns"] and not args.cpu_patterns:
            note = " (host memory %.0f GB cannot hold the tape of %d patterns)" % (
                avail / 1e9, cfg["patterns"])
        prob = build_problem(cfg, patterns=patterns)
        rp = ra.ReferenceProblem(prob)
        evaluate = rp.evaluate
        what = ("vendored reference (baseline/_ref, unmodified): GTR/GeneralSymmetric.p_t + "
                "WeibullSiteModel + calculate_treelikelihood_discrete_rescaled + .backward()")
    else:
        from oracle import treelik as orc

        patterns = args.cpu_patterns or min(cfg["patterns"], 8000)
        prob = build_problem(cfg, patterns=patterns, draws=min(cfg["draws"], 4))
        evaluate = lambda: orc.evaluate(prob, want_grad=True)  # noqa: E731
        what = "oracle port (oracle/treelik.py; baseline/_ref absent or a batch of draws)"
    # first evaluation = warm-up, timed to plan the rest inside the budget
    t0 = time.perf_counter()
    evaluate()
    t_first = time.perf_counter() - t0
    warm = max(1, args.warmup) if t_first * (args.warmup + args.steps) <= args.ref_budget_s else 1
    for _ in range(warm - 1):
        evaluate()
    left = args.ref_budget_s - t_first * warm
    steps = max(1, min(args.steps, int(left / max(t_first, 1e-9)))) if args.steps > 0 else 1
    steps = max(steps, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        evaluate()
    dt = (time.perf_counter() - t0) / steps
    value = prob.units / dt
    base = {"kind": kind, "cores": threads, "value": value, "unit": UNIT,
            "sample": "%s; %d taxa x %d patterns x K=%d x %d draw(s) per step%s; %d timed steps after "
                      "%d warm-up (%.2f s/step; steps requested %d, bounded by --ref-budget-s %.0f)"
                      % (what, cfg["taxa"], patterns, cfg["categories"], prob.draws, note, steps, warm,
                         dt, args.steps, args.ref_budget_s)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(cfg, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = args.cfg
    line = run_engine(cfg, args, rank, local_rank, world, dist)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import reference_arm as ra

            if cfg["index"] == 2 and ra.available():
                sample = args.cpu_patterns or min(cfg["patterns"], 25_000)
                line["cpu_baseline"] = parity_vs_reference(cfg, sample, local_rank)
            else:
                par = parity_vs_oracle(cfg, local_rank)
                line["cpu_baseline"] = {
                    "parity_on_sample": par, "value": par["oracle_units_per_s"], "unit": UNIT,
                    "cores": os.cpu_count() or 1, "kind": "port",
                    "sample": par["against"] + ", one evaluation"}
            if not line["cpu_baseline"]["parity_on_sample"]["ok"]:
                raise SystemExit("bench.py: engine and CPU baseline disagree on the sample: %r"
                                 % (line["cpu_baseline"]["parity_on_sample"],))
        if world == 1 and cfg["index"] == 3:
            line["pipeline"] = config3_pipeline(cfg, local_rank)
        if world == 1 and cfg["index"] == 2 and not args.no_other_configs and args.taxa is None \
                and args.patterns is None:
            # the other BASELINE shapes, measured in the same run (each takes about a second)
            others = {}
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup = min(args.steps, 5), 3
            for idx in (3, 4, 5):
                c = dict(CONFIGS[idx], index=idx, topology="random")
                o = run_engine(c, sub, 0, local_rank, 1, dist, with_clocks=False)
                par = parity_vs_oracle(c, local_rank)
                if not par["ok"]:
                    raise SystemExit("bench.py: config %d disagrees with the oracle: %r" % (idx, par))
                entry = {"workload": c["workload"], "ms_per_eval": o["ms_per_step"],
                         "units_per_s": o["value"], "e2e_units_per_s": o["e2e"]["value"],
                         "e2e_ms": o["e2e"]["ms_per_step"], "gpu_launches": o["gpu_launches"],
                         "roofline": o["roofline"], "parity_on_sample": par}
                if idx == 3:
                    entry["pipeline"] = config3_pipeline(c, local_rank)
                others[c["name"]] = entry
            line["other_configs"] = others
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
