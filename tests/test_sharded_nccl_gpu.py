"""Multi-GPU through the product API on real hardware: launches
tests/nccl_worker.py under torchrun with one rank per visible GPU (NCCL over NVLink) and
checks that the pattern-sharded and the draw-sharded evaluations equal the single-GPU one to
1e-12 -- lnL and every gradient.  Skipped on a single-GPU box (the world-size-2 gloo tests in
tests/test_sharded_gloo.py cover the host logic there)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_equals_single_gpu_over_nccl():
    world = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == world
    for key in ("patterns_lnL", "patterns_grads", "fast_lnL", "fast_grads", "fast_scaled_grads",
                "fast_nograd_lnL", "packed_lnL",
                "packed_grads", "draws_lnL", "draws_grads"):
        assert out[key] <= 1e-12, (key, out)
