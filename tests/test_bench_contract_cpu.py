"""bench.py contract pieces that can run without a GPU: the reference arm (`--impl reference`
times the oracle port on the host cores and prints one JSON line with the agreed keys) and the
refusal of the product arm to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, cwd=REPO,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--taxa", "30",
              "--patterns", "1000", "--cpu-patterns", "200"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], cwd=REPO, env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_arm_refuses_to_run_without_cuda():
    r = _run(["--steps", "1", "--warmup", "0", "--taxa", "10", "--patterns", "100"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_traffic_model_counts_what_the_sweeps_move_and_compute():
    """bench.traffic_model: bytes / flops per (pattern, category) from the topology alone.  A
    cherry ((4,0,1)) under a root ((5,4,2)), 3 tips; with cherry tabulation the cherry's vector is
    never stored; with kept u (61-state path) the pre-order sweep reads u back instead of
    recomputing it and the post-order sweep writes it."""
    import bench

    post = [(3, 0, 1), (4, 3, 2)]
    S = 4
    V = S * 8
    plain = bench.traffic_model(post, 3, S, cherries=False)
    # post-order: node 3 writes 1 vector; node 4 writes 1 and reads its stored child (node 3)
    assert plain["post_bytes"] * 2 == 3 * V
    # pre-order: both read q^ (2 V); node 4 reads node 3's vector and writes q^_3 (2 V)
    assert plain["pre_bytes"] * 2 == 4 * V
    tab = bench.traffic_model(post, 3, S, cherries=True)
    assert tab["cherry_nodes"] == 1
    assert tab["post_bytes"] * 2 == 1 * V          # only the root's vector is written
    assert tab["pre_bytes"] * 2 == 3 * V           # the tabulated cherry still receives q^
    S = 61
    V = S * 8
    rec = bench.traffic_model(post, 3, S, cherries=False)
    kept = bench.traffic_model(post, 3, S, cherries=False, kept_u=True)
    # one internal child in the tree: u = P p~ once per sweep (recompute) or once in all (kept)
    assert rec["flops_pre"] - kept["flops_pre"] == 2 * S * S / 2
    assert kept["flops_post"] == rec["flops_post"]
    assert kept["post_bytes"] - rec["post_bytes"] == V / 2
    assert kept["pre_bytes"] - rec["pre_bytes"] == V / 2
