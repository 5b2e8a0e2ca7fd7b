#!/usr/bin/env python
"""Benchmark of the tree-likelihood hot path (logL + full gradient).

    python bench.py --gpus N --steps K --warmup W          # CUDA engine
    python bench.py --impl reference --gpus N ...           # CPU arm (oracle port)

Workload (BASELINE.json configs[1]): synthetic 1,000 taxa x 100,000 site
patterns, 4 states, K=4 rate categories, unrooted, one draw; a "step" is one
logL + gradient evaluation.  With N > 1 the site patterns are sharded across
ranks (strong scaling: the problem is fixed) and the packed
{lnL, gradient} vector is all-reduced over NCCL.

Metric: patterns x internal-nodes x categories processed per second (fp64).
`value` has the inputs resident in HBM; `e2e` goes through the public API with
pinned host buffers (host->device parameters, device->host lnL + gradient).
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
BYTES_PER_UNIT = 160.0  # 5 vectors x 4 states x 8 B per (pattern, node, cat)
# ncu dram__bytes_read.sum + dram__bytes_write.sum over the pre-order sweep of one step at
# the headline size on one GPU (profiles/r01_cherry_dram_bytes.csv; 38.07e9 without the cherry
# tabulation, profiles/r01_tma_dram_bytes.csv)
NCU_PREORDER_SWEEP_BYTES = 33.83e9
BYTES_PER_UNIT_PRE = 96.0  # pre-order sweep share (3 vectors)
BYTES_PER_UNIT_POST = 64.0  # post-order sweep share (2 vectors)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=100_000)
    ap.add_argument("--categories", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20260101)
    ap.add_argument("--cpu-patterns", type=int, default=8000,
                    help="patterns of the bounded CPU-baseline sample")
    ap.add_argument("--topology", default="random", choices=["random", "caterpillar", "balanced"],
                    help="tree shape (the headline workload is the random-join tree)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engine-flags", type=int, default=0,
                    help="extra TTB2_FLAG_* bits for experiments (32 = no CUDA graphs)")
    return ap.parse_args()


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"  # B200_PROFILING.md fallback


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                mhz.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {
            "sm_mhz": float(np.median(mhz)) if mhz else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(mhz),
            "reasons": sorted(reasons),
        }


def build_problem(args, lo=0, hi=None, patterns=None):
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(args.taxa, patterns or args.patterns, 4, args.categories,
                        seed=args.seed, topology=args.topology)
    if hi is not None:
        prob.tip_states = np.ascontiguousarray(prob.tip_states[:, lo:hi])
        prob.weights = np.ascontiguousarray(prob.weights[lo:hi])
        prob.pattern_count = hi - lo
    return prob


def engine_vs_oracle(prob, ref, device):
    """Parity gate of the run: the engine on the CPU baseline's own sample
    against the oracle's result -- lnL rel <= 1e-10, gradients rel <= 1e-8."""
    from torchtree_b200 import Engine, reversible_eigensystem

    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, prob.category_count,
                 max_draws=1, device=device)
    q, f = torch.tensor(prob.q_matrix), torch.tensor(prob.freqs)
    evec, ivec, evals = reversible_eigensystem(q, f)
    lnl = eng.loglik_eigen(torch.tensor(prob.branch_lengths), torch.tensor(prob.site_rates),
                           torch.tensor(prob.site_props), evec, ivec, evals, f)
    g = eng.grad_eigen()
    eng.close()

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    out = {"lnL_rel_err": rel(lnl.cpu().numpy(), ref["lnL"]),
           "branch_grad_rel_err": rel(g["branch_lengths"].cpu().numpy(), ref["branch_lengths"]),
           "site_rate_grad_rel_err": rel(g["site_rates"].cpu().numpy(), ref["site_rates"]),
           "root_freq_grad_rel_err": rel(g["freqs"].cpu().numpy(), ref["freqs"])}
    out["ok"] = bool(out["lnL_rel_err"] <= 1e-10 and all(
        v <= 1e-8 for k, v in out.items() if k.endswith("grad_rel_err")))
    return out


def cpu_baseline(args, threads=None, check_device=None):
    """The oracle port (torch CPU ops + autograd, like the reference) on a bounded
    sample of the same workload: same tree and model, fewer patterns."""
    from oracle import treelik as orc

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    prob = build_problem(args, patterns=args.cpu_patterns)
    ref = orc.evaluate(prob, want_grad=True)  # warm-up
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        orc.evaluate(prob, want_grad=True)
        times.append(time.perf_counter() - t0)
    best = min(times)
    parity = engine_vs_oracle(prob, ref, check_device) if check_device is not None else None
    return {
        "parity_on_sample": parity,
        "value": prob.units / best,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": "%d taxa x %d patterns x K=%d, logL+autograd gradient, best of 2 (%.2f s/eval)"
        % (args.taxa, args.cpu_patterns, args.categories, best),
        "evals_per_s_at_full_size": (prob.units / best) / (
            args.patterns * (args.taxa - 1) * args.categories),
    }


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import treelik as orc

    prob = build_problem(args, patterns=args.cpu_patterns)
    for _ in range(min(args.warmup, 1)):
        orc.evaluate(prob, want_grad=True)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.evaluate(prob, want_grad=True)
    dt = (time.perf_counter() - t0) / steps
    value = prob.units / dt
    base = {
        "kind": "port", "cores": os.cpu_count() or 1, "value": value, "unit": UNIT,
        "sample": "%d taxa x %d patterns x K=%d per step (bounded sample of the %d-pattern "
                  "workload; throughput is ~linear in patterns)"
        % (args.taxa, args.cpu_patterns, args.categories, args.patterns),
    }
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1, args.cpu_patterns),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world, patterns_per_rank):
    return {
        "workload": "BASELINE.json configs[1]: synthetic %d taxa x %d site patterns, "
                    "GTR-class reversible 4-state model + 4 discrete-rate categories, unrooted, "
                    "one logL+gradient evaluation per step" % (args.taxa, args.patterns),
        "taxa": args.taxa, "patterns": args.patterns, "states": 4,
        "categories": args.categories, "draws": 1,
        "sharding": "patterns/%d" % world, "patterns_per_gpu": patterns_per_rank,
        "l2": "working set (%.1f GB of partials per GPU) >> 126 MB L2; no explicit flush"
        % (2 * patterns_per_rank * (args.taxa - 1) * args.categories * 32 / 1e9),
    }


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from torchtree_b200 import Engine, reversible_eigensystem

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- problem, sharded by patterns ----
    from torchtree_b200.sharded import shard_range

    lo, hi = shard_range(args.patterns, rank, world)
    prob = build_problem(args, lo, hi)
    units_total = args.patterns * (args.taxa - 1) * args.categories
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, args.categories,
                 max_draws=1, device=local_rank, flags=1 | args.engine_flags)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)

    q = torch.tensor(prob.q_matrix)
    f = torch.tensor(prob.freqs)
    evec, ivec, evals = reversible_eigensystem(q, f)
    host = [torch.tensor(prob.branch_lengths), torch.tensor(prob.site_rates),
            torch.tensor(prob.site_props), evec, ivec, evals, f]
    host = [t.contiguous().pin_memory() for t in host]
    devin = [t.to(dev) for t in host]
    B, K = prob.branch_count, args.categories
    packed_n = 1 + B + K + K + 16 + 4

    def make_out(device, pin=False):
        def mk(*shape):
            t = torch.empty(shape, dtype=torch.float64, device=device)
            return t.pin_memory() if pin else t
        return mk(1), dict(branch_lengths=mk(1, B), site_rates=mk(1, K), props=mk(1, K),
                           q=mk(1, 4, 4), freqs=mk(1, 4))

    lnl_d, out_d = make_out(dev)
    lnl_h, out_h = make_out("cpu", pin=True)
    packed = torch.empty(packed_n, dtype=torch.float64, device=dev)
    packed_h = torch.empty(packed_n, dtype=torch.float64).pin_memory()

    def pack(lnl, g):
        torch.cat([lnl.reshape(-1), g["branch_lengths"].reshape(-1), g["site_rates"].reshape(-1),
                   g["props"].reshape(-1), g["q"].reshape(-1), g["freqs"].reshape(-1)], out=packed)

    def step_device():
        eng.loglik_eigen(*devin, out=lnl_d)
        eng.grad_eigen(out=out_d)
        if world > 1:
            pack(lnl_d, out_d)
            dist.all_reduce(packed)

    # the call a user makes: the differentiable op of the torch extension (csrc/torch_ext.cpp),
    # host tensors in (pinned), generator decomposed on the device, .backward() hands the
    # gradients back as host tensors, lnL read on the host
    from torchtree_b200 import log_likelihood_eigen

    user_in = [t.clone().pin_memory().requires_grad_(True) for t in
               (host[0], host[1], host[2], q.contiguous(), host[6])]

    def step_e2e():
        if world == 1:
            for t in user_in:
                t.grad = None
            lnl = log_likelihood_eigen(eng, *user_in)
            lnl.sum().backward()
            return lnl.detach()
        d_in = [t.to(dev, non_blocking=True) for t in host]
        eng.loglik_eigen(*d_in, out=lnl_d)
        eng.grad_eigen(out=out_d)
        pack(lnl_d, out_d)
        dist.all_reduce(packed)
        packed_h.copy_(packed, non_blocking=True)
        stream.synchronize()
        return packed_h

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

    # ---- device-resident measurement (value) ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ms_dev, _ = timed(step_device, args.steps)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- phase timing for the roofline of the dominant kernel family ----
    eng.enable_timing(True)
    pre_ms, post_ms = [], []
    for _ in range(min(5, args.steps)):
        step_device()
        ph = eng.phase_ms()
        pre_ms.append(ph["preorder"])
        post_ms.append(ph["postorder"])
    eng.enable_timing(False)
    ph_pre = float(np.mean(pre_ms))
    ph_post = float(np.mean(post_ms))

    # ---- end-to-end through the public API with pinned host buffers ----
    for _ in range(3):
        step_e2e()
    _, wall_e2e = timed(step_e2e, args.steps)
    lnl_value = float(step_e2e()[0])

    peak, peak_kind = measured_peak_gbs()
    units_rank = prob.units
    if rank == 0:
        value = units_total / (ms_dev * 1e-3)
        e2e_value = units_total / wall_e2e
        ach_pre = units_rank * BYTES_PER_UNIT_PRE / (ph_pre * 1e-3) / 1e9
        ach_post = units_rank * BYTES_PER_UNIT_POST / (ph_post * 1e-3) / 1e9
        ach_step = units_rank * BYTES_PER_UNIT / (ms_dev * 1e-3) / 1e9
        h2d = sum(t.numel() * 8 for t in (user_in if world == 1 else host))
        d2h = packed_n * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world, hi - lo),
            "evals_per_s": 1e3 / ms_dev, "lnL": lnl_value,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": wall_e2e * 1e3,
                    "api": ("torchtree_b200.log_likelihood_eigen(...).backward(): torch C++ extension "
                            "autograd Function -> ttb2_loglik_q / ttb2_grad_eigen, pinned host tensors "
                            "in, host gradients out") if world == 1 else
                           "Engine.loglik_eigen + Engine.grad_eigen on device buffers + NCCL all-reduce "
                           "+ copy of the packed result to pinned host memory"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm",
                "kernel": "bwd4_tma_kernel<3,5> (pre-order sweep, one launch per tree level; "
                          "level 1 is bwd4_tips_tma_kernel<4>)",
                "achieved": ach_pre, "peak": peak, "unit": "GB/s", "frac": ach_pre / peak,
                "peak_kind": peak_kind,
                "traffic": (NCU_PREORDER_SWEEP_BYTES / max(1, ph["preorder_launches"])
                            if world == 1 and args.taxa == 1000 and args.patterns == 100_000
                            and args.categories == 4 and args.topology == "random" else None),
                "traffic_unit": "bytes per launch (ncu dram__bytes_read+write summed over the "
                                "pre-order sweep / launches, profiles/r01_cherry_dram_bytes.csv); below "
                                "the algorithmic bytes because cherry vectors are tabulated, not read",
                "algorithmic_bytes_per_launch": units_rank * BYTES_PER_UNIT_PRE
                / max(1, ph["preorder_launches"]),
                "algorithmic_bytes_per_unit": BYTES_PER_UNIT_PRE,
                "launches_per_step": ph["preorder_launches"],
                "avg_launch_ms": ph_pre / max(1, ph["preorder_launches"]),
                "postorder": {"kernel": "fwd4c_kernel<4> (post-order sweep; level 1 is tabulated by "
                                        "cherry_table_kernel instead of being stored)",
                              "achieved": ach_post,
                              "frac": ach_post / peak,
                              "algorithmic_bytes_per_unit": BYTES_PER_UNIT_POST,
                              "launches_per_step": ph["postorder_launches"], "ms": ph_post},
                "whole_step": {"achieved": ach_step, "frac": ach_step / peak,
                               "algorithmic_bytes_per_unit": BYTES_PER_UNIT,
                               "note": "algorithmic bytes (SURVEY 8d) / time; cherry tabulation moves "
                                       "~20 % fewer bytes than that, so this can exceed 1"},
            },
            "phases_ms": {"note": "CUDA events between kernel groups, ordinary launches (the timed "
                                  "steps above replay CUDA graphs)",
                          **{k: (round(v, 4) if isinstance(v, float) else v)
                             for k, v in ph.items()}},
            "device_bytes": eng.device_bytes,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, check_device=local_rank)
            if not line["cpu_baseline"]["parity_on_sample"]["ok"]:
                raise SystemExit("bench.py: engine and oracle disagree on the baseline sample: %r"
                                 % (line["cpu_baseline"]["parity_on_sample"],))
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
