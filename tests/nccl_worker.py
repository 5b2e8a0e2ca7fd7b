"""Worker of tests/test_sharded_nccl_gpu.py (one process per GPU under torchrun): the sharded
product API on the real CUDA engine over NCCL against the single-GPU evaluation.

  * pattern sharding: sharded_log_likelihood(lambda *a: log_likelihood_eigen(engine_shard, *a))
    == log_likelihood_eigen(engine_all) -- lnL and every gradient to 1e-12;
  * draw sharding (BASELINE config 3 shape): draw_sharded_log_likelihood over a batch of draws;
  * the packed C-ABI gradient (ttb2_grad_eigen_packed) all-reduced in place on the device.
Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from torchtree_b200 import Engine, log_likelihood_eigen  # noqa: E402
from torchtree_b200.sharded import (draw_sharded_log_likelihood, shard_range,  # noqa: E402
                                    sharded_engine_log_likelihood, sharded_log_likelihood)
from torchtree_b200.synthetic import make_problem  # noqa: E402


def leaves(prob):
    return [torch.tensor(x, requires_grad=True) for x in
            (prob.branch_lengths, prob.site_rates, prob.site_props, prob.q_matrix, prob.freqs)]


def rel(a, b):
    a, b = a.detach().numpy(), b.detach().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world}

    # ---- pattern sharding ----
    prob = make_problem(60, 1003, 4, 4, seed=5, gap_fraction=0.02)   # ragged: 1003 patterns
    lo, hi = shard_range(prob.pattern_count, rank, world)
    shard = Engine(prob.tip_states[:, lo:hi], prob.weights[lo:hi], prob.postorder, 4, 4,
                   device=local)
    full = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 4, device=local)
    a, b = leaves(prob), leaves(prob)
    v_sh = sharded_log_likelihood(lambda *t: log_likelihood_eigen(shard, *t), a)
    v_sh.sum().backward()
    v_one = log_likelihood_eigen(full, *b)
    v_one.sum().backward()
    out["patterns_lnL"] = rel(v_sh, v_one)
    out["patterns_grads"] = max(rel(x.grad, y.grad) for x, y in zip(a, b))

    # ---- the engine-aware fast path (collectives on the device): what the model's
    # "shard": "patterns" key uses over NCCL ----
    c = leaves(prob)
    v_fast = sharded_engine_log_likelihood(shard, c)
    v_fast.sum().backward()
    out["fast_lnL"] = rel(v_fast, v_one)
    out["fast_grads"] = max(rel(x.grad, y.grad) for x, y in zip(c, b))
    # (one draw: the gradient travels with the forward pass and backward() scales it; without a
    # gradient request only lnL is reduced)
    c2 = leaves(prob)
    (2.5 * sharded_engine_log_likelihood(shard, c2)).sum().backward()
    out["fast_scaled_grads"] = max(rel(x.grad, 2.5 * y.grad) for x, y in zip(c2, b))
    with torch.no_grad():
        out["fast_nograd_lnL"] = rel(sharded_engine_log_likelihood(shard, [t.detach() for t in c2]),
                                     v_one)

    # ---- the packed gradient, reduced in place on the device ----
    dev = torch.device("cuda", local)
    din = [torch.tensor(x).to(dev) for x in (prob.branch_lengths, prob.site_rates,
                                             prob.site_props, prob.q_matrix, prob.freqs)]
    shard.loglik_q(*din)
    packed = shard.grad_eigen_packed()
    dist.all_reduce(packed)
    views = shard.unpack(packed)
    out["packed_lnL"] = rel(views["lnL"].cpu(), v_one)
    out["packed_grads"] = max(rel(views[k].cpu(), t.grad) for k, t in
                              zip(("branch_lengths", "site_rates", "props", "q", "freqs"), b))
    shard.close()
    full.close()

    # ---- draw sharding: a batch of 5 draws (ragged over the ranks), per-draw branch lengths and
    # site rates, shared generator ----
    D = 5
    prob = make_problem(40, 257, 4, 2, draws=D, seed=9)
    prob.site_rates = np.repeat(prob.site_rates, D, 0) * (1 + 0.05 * np.arange(D))[:, None]
    dlo, dhi = shard_range(D, rank, world)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 2, max_draws=D, device=local)
    a, b = leaves(prob), leaves(prob)
    w = torch.linspace(0.5, 1.5, D, dtype=torch.float64)
    v_sh = draw_sharded_log_likelihood(lambda *t: log_likelihood_eigen(eng, *t), a, D)
    (v_sh * w).sum().backward()
    v_one = log_likelihood_eigen(eng, *b)
    (v_one * w).sum().backward()
    out["draws_lnL"] = rel(v_sh, v_one)
    out["draws_grads"] = max(rel(x.grad, y.grad) for x, y in zip(a, b))
    eng.close()

    dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
