"""20-state (amino-acid) path: the warp-autonomous DMMA kernels (csrc/kernels_gwarp.cu) and the
shared-memory tile kernels they replaced (TTB2_GM_LEGACY=1, still the path of every other
alphabet between 8 and 64 states) against the CPU oracle, at sizes that exercise what the
small cases in test_engine_oracle_gpu.py do not: several 8-pattern groups per warp (ring
wrap-around), several pattern chunks per node, tip code tables on both sides of the
tabulation limit, batched draws, zero-weight patterns."""
import dataclasses

import numpy as np
import pytest
import torch

from helpers import assert_grad_close, assert_lnl_close
from test_engine_oracle_gpu import _check, _run

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["warp", "tile"])
def kernels(request, monkeypatch):
    if request.param == "tile":
        monkeypatch.setenv("TTB2_GM_LEGACY", "1")
    else:
        monkeypatch.delenv("TTB2_GM_LEGACY", raising=False)
    return request.param


@pytest.mark.parametrize("topology,T,N", [("random", 40, 3000), ("caterpillar", 24, 1500),
                                          ("balanced", 32, 2100)])
def test_many_groups_and_chunks(kernels, topology, T, N):
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(T, N, 20, 4, seed=T + N, topology=topology, gap_fraction=0.03), q_rtol=1e-6)


@pytest.mark.parametrize("extra", [2, 9])   # 23 codes: tabulated tips; 30 codes: beyond the table limit
def test_ambiguity_codes(kernels, extra):
    """B / Z / J / X-like ambiguity masks as additional tip codes."""
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(20, 900, 20, 2, seed=extra, gap_fraction=0.02)
    rng = np.random.default_rng(extra)
    table = np.concatenate([np.eye(20), np.ones((1, 20))], 0)
    rows = []
    for _ in range(extra):
        m = np.zeros(20)
        m[rng.choice(20, size=rng.integers(2, 5), replace=False)] = 1.0
        rows.append(m)
    table = np.concatenate([table, np.array(rows)], 0)
    tips = prob.tip_states.copy()
    hit = rng.random(tips.shape) < 0.1
    tips[hit] = rng.integers(21, 21 + extra, size=int(hit.sum()))
    prob = dataclasses.replace(prob, tip_states=tips.astype(np.uint8), code_partials=table)
    _check(prob, q_rtol=1e-6)


def test_draws_and_masked_patterns(kernels):
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(16, 700, 20, 3, draws=2, seed=4, per_draw_model=True)
    w = prob.weights.copy()
    w[::5] = 0.0
    _check(dataclasses.replace(prob, weights=w), q_rtol=1e-6)


def test_bitwise_reproducible_and_kernel_families_agree(monkeypatch):
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(30, 2500, 20, 4, seed=8, gap_fraction=0.05)
    monkeypatch.delenv("TTB2_GM_LEGACY", raising=False)
    e1, a = _run(prob)
    e2, b = _run(prob)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    monkeypatch.setenv("TTB2_GM_LEGACY", "1")
    e3, c = _run(prob)
    assert_lnl_close(a["lnL"], c["lnL"], rtol=1e-13)
    for k in ("branch_lengths", "site_rates", "props", "freqs", "q"):
        assert_grad_close(a[k], c[k], rtol=1e-10, what=k)
    for e in (e1, e2, e3):
        e.close()


@pytest.mark.parametrize("S,extra", [(8, 12), (20, 16)])
def test_tile_kernels_with_untabulated_tips(monkeypatch, S, extra):
    """More tip codes than the table that replaces the tip side's staged P can hold
    (gm_utab_fits): the tile kernels multiply tip children like any other child."""
    from torchtree_b200.synthetic import make_problem

    monkeypatch.setenv("TTB2_GM_LEGACY", "1")
    prob = make_problem(14, 500, S, 2, seed=S + extra, gap_fraction=0.02)
    rng = np.random.default_rng(extra)
    rows = []
    for _ in range(extra):
        m = np.zeros(S)
        m[rng.choice(S, size=rng.integers(2, 4), replace=False)] = 1.0
        rows.append(m)
    table = np.concatenate([np.eye(S), np.ones((1, S)), np.array(rows)], 0)
    tips = prob.tip_states.copy()
    hit = rng.random(tips.shape) < 0.15
    tips[hit] = rng.integers(S + 1, S + 1 + extra, size=int(hit.sum()))
    prob = dataclasses.replace(prob, tip_states=tips.astype(np.uint8), code_partials=table)
    _check(prob, q_rtol=1e-6)
