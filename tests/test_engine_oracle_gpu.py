"""CUDA engine vs the pinned CPU oracle on seeded synthetic problems, plus
size-independent properties at the BASELINE.json configuration sizes."""
import numpy as np
import pytest
import torch

from helpers import assert_grad_close, assert_lnl_close

pytestmark = pytest.mark.gpu


def _run(prob, flags=0, want_grad=True):
    from torchtree_b200 import Engine, reversible_eigensystem

    eng = Engine(prob.tip_states, prob.weights, prob.postorder, prob.state_count,
                 prob.category_count, code_partials=prob.code_partials,
                 max_draws=prob.draws, flags=flags)
    q = torch.tensor(prob.q_matrix)
    f = torch.tensor(prob.freqs)
    evec, ivec, evals = reversible_eigensystem(q, f)
    lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props,
                           evec, ivec, evals, prob.freqs)
    out = {"lnL": lnl.numpy().copy()}
    if want_grad:
        g = eng.grad_eigen()
        out.update({k: v.numpy().copy() for k, v in g.items()})
    return eng, out


def _check(prob, flags=0, q_rtol=1e-7):
    from oracle import treelik as orc

    want = orc.evaluate(prob, want_grad=True, through_q=True)
    want["q_matrix"] = orc.evaluate(prob, want_grad=True, route="expm")["q_matrix"]
    eng, got = _run(prob, flags)
    assert_lnl_close(got["lnL"], want["lnL"])
    assert_grad_close(got["branch_lengths"], want["branch_lengths"], what="d_bl")
    assert_grad_close(got["site_rates"], want["site_rates"], what="d_rates")
    assert_grad_close(got["props"], want["site_props"], what="d_props")
    assert_grad_close(got["freqs"], want["freqs"], what="d_freqs")
    assert_grad_close(got["q"], want["q_matrix"], rtol=q_rtol, what="d_q")
    eng.close()


@pytest.mark.parametrize("flags", [0, 64], ids=["levels", "nocherry"])
@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8])
def test_category_counts(K, flags):
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(33, 257, 4, K, seed=100 + K, gap_fraction=0.05), flags=flags)


@pytest.mark.parametrize("flags", [0, 64], ids=["levels", "nocherry"])
@pytest.mark.parametrize("topology", ["random", "caterpillar", "balanced"])
def test_topologies_with_rescaling(topology, flags):
    from torchtree_b200.synthetic import make_problem

    # iid tips on >= 300 taxa underflow without rescaling
    _check(make_problem(320, 96, 4, 4, seed=7, topology=topology), flags=flags)


def test_ragged_pattern_counts_and_single_pattern():
    from torchtree_b200.synthetic import make_problem

    for n in (1, 31, 32, 33, 255, 1000):
        _check(make_problem(12, n, 4, 4, seed=n))


def test_two_tips_tree():
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(2, 40, 4, 4, seed=3))


def test_batched_draws_with_per_draw_models():
    from torchtree_b200.synthetic import make_problem

    for flags in (0, 4, 8, 16):
        _check(make_problem(25, 130, 4, 4, draws=5, seed=17, per_draw_model=True), flags=flags)
        _check(make_problem(25, 130, 4, 4, draws=5, seed=18, per_draw_model=False), flags=flags)


def test_generic_kernels_equal_specialised_for_nucleotides():
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(40, 500, 4, 4, seed=5, gap_fraction=0.1)
    e1, a = _run(prob, flags=0)
    e2, b = _run(prob, flags=2)
    assert_lnl_close(a["lnL"], b["lnL"], rtol=1e-13)
    for k in ("branch_lengths", "site_rates", "props", "freqs", "q"):
        assert_grad_close(a[k], b[k], rtol=1e-10, what=k)
    _check(prob, flags=2)
    e1.close(); e2.close()


@pytest.mark.parametrize("flags", [0, 8], ids=["mma", "simt"])
@pytest.mark.parametrize("S,K,T,N", [(20, 4, 14, 70), (61, 4, 9, 40), (20, 1, 30, 33),
                                     (7, 3, 10, 64), (2, 2, 10, 64), (8, 2, 12, 100),
                                     (64, 1, 6, 33), (21, 3, 40, 700)])
def test_other_state_counts(S, K, T, N, flags):
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(T, N, S, K, seed=S + K, gap_fraction=0.05), flags=flags, q_rtol=1e-6)


def test_weights_linearity_and_shard_additivity():
    """lnL is linear in the pattern weights and additive over pattern shards
    (the property pattern-sharding across GPUs relies on)."""
    from torchtree_b200.synthetic import Problem, make_problem
    import dataclasses

    prob = make_problem(60, 4000, 4, 4, seed=11)
    eng, full = _run(prob)
    site = eng.site_loglik().numpy()
    assert_lnl_close((site * prob.weights).sum(-1), full["lnL"], rtol=1e-12)
    half = prob.pattern_count // 2
    parts = []
    for sl in (slice(0, half), slice(half, None)):
        sub = dataclasses.replace(prob, pattern_count=len(prob.weights[sl]),
                                  tip_states=prob.tip_states[:, sl].copy(),
                                  weights=prob.weights[sl].copy())
        e, o = _run(sub)
        parts.append(o)
        e.close()
    assert_lnl_close(parts[0]["lnL"] + parts[1]["lnL"], full["lnL"], rtol=1e-12)
    for k in ("branch_lengths", "site_rates", "props", "freqs", "q"):
        assert_grad_close(parts[0][k] + parts[1][k], full[k], rtol=1e-9, what=k)
    eng.close()


def test_bitwise_reproducible():
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(100, 3000, 4, 4, seed=2)
    e1, a = _run(prob)
    e2, b = _run(prob)
    e3, c = _run(prob, flags=32)  # without CUDA-graph replay: same kernels, same bits
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a[k], c[k]), k
    e1.close(); e2.close(); e3.close()


def test_graph_replay_many_evaluations():
    """The captured graphs are replayed with new parameter values every call."""
    from oracle import treelik as orc
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(40, 300, 4, 4, seed=12)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 4)
    evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix),
                                               torch.tensor(prob.freqs))
    rng = np.random.default_rng(0)
    for it in range(6):
        prob.branch_lengths = prob.branch_lengths * rng.uniform(0.8, 1.25, prob.branch_lengths.shape)
        prob.branch_lengths[:, -1] = 0.0
        lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props,
                               evec, ivec, evals, prob.freqs)
        if it % 2 == 0:  # forward-only calls interleaved with forward+backward
            continue
        g = eng.grad_eigen()
        want = orc.evaluate(prob, want_grad=True, through_q=True)
        assert_lnl_close(lnl.numpy(), want["lnL"])
        assert_grad_close(g["branch_lengths"].numpy(), want["branch_lengths"], what="d_bl")
    eng.close()


def test_headline_size_properties():
    """Config 2 shape (1000 taxa, K=4) at 100k patterns: finite, reproducible,
    shard-additive, and the branch gradient agrees with a central difference of
    the engine's own lnL on a few branches."""
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(1000, 100_000, 4, 4, seed=20260101)
    eng, out = _run(prob)
    assert np.isfinite(out["lnL"]).all()
    for k in ("branch_lengths", "site_rates", "props", "freqs", "q"):
        assert np.isfinite(out[k]).all(), k
    site = eng.site_loglik().numpy()
    assert_lnl_close((site * prob.weights).sum(-1), out["lnL"], rtol=1e-12)
    q = torch.tensor(prob.q_matrix)
    f = torch.tensor(prob.freqs)
    evec, ivec, evals = reversible_eigensystem(q, f)
    rng = np.random.default_rng(0)
    for b in rng.choice(prob.branch_count - 1, size=3, replace=False):
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            bl = prob.branch_lengths.copy()
            bl[0, b] += sgn * h
            vals.append(eng.loglik_eigen(bl, prob.site_rates, prob.site_props, evec, ivec,
                                         evals, prob.freqs).item())
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - out["branch_lengths"][0, b]) <= 1e-5 * max(1.0, abs(fd)), (b, fd)
    eng.close()


def test_invalid_inputs_raise():
    from torchtree_b200 import Engine, EngineError
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(8, 16, 4, 2, seed=1)
    bad = prob.postorder.copy()
    bad[0, 1] = bad[0, 2]
    with pytest.raises(EngineError):
        Engine(prob.tip_states, prob.weights, bad, 4, 2)
    eng = Engine(prob.tip_states, prob.weights, prob.postorder, 4, 2)
    with pytest.raises(EngineError):
        eng.grad_eigen()
    with pytest.raises(EngineError):
        eng.loglik_mats(np.zeros((2, prob.branch_count, 2, 4, 4)), prob.freqs, prob.site_props)
    eng.close()


def test_nan_in_nan_out():
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(10, 64, 4, 4, seed=4)
    prob.branch_lengths[0, 3] = np.nan
    eng, out = _run(prob, want_grad=False)
    assert np.isnan(out["lnL"]).all()
    eng.close()


def test_invariant_category_zero_rate():
    """Rate-0 category (InvariantSiteModel, site_model.py:77-112): P = I exactly,
    so variable patterns have an all-zero vector in that category."""
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(30, 200, 4, 4, seed=77)
    prob.site_rates = prob.site_rates.copy()
    prob.site_rates[0, 0] = 0.0
    prob.site_rates /= (prob.site_rates * prob.site_props).sum()
    for flags in (0, 4, 8, 16):
        _check(prob, flags=flags)


@pytest.mark.parametrize("S", [4, 20])
def test_set_postorder_rebuilds_schedule(S):
    """Topology change on a live engine (TreeModel.update_traversals,
    tree_model.py:187-195): same tips, new post-order."""
    import dataclasses

    from oracle import treelik as orc
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    a = make_problem(30, 200, S, 3, seed=41, topology="random")
    b = make_problem(30, 200, S, 3, seed=42, topology="caterpillar")
    b = dataclasses.replace(b, tip_states=a.tip_states, weights=a.weights)
    eng = Engine(a.tip_states, a.weights, a.postorder, S, 3)
    for prob in (a, b, a):
        eng.set_postorder(prob.postorder)
        evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix),
                                                   torch.tensor(prob.freqs))
        lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props,
                               evec, ivec, evals, prob.freqs)
        g = eng.grad_eigen()
        want = orc.evaluate(prob, want_grad=True, through_q=True)
        assert_lnl_close(lnl.numpy(), want["lnL"])
        assert_grad_close(g["branch_lengths"].numpy(), want["branch_lengths"], what="d_bl")
    eng.close()


def test_varying_draw_counts_on_one_engine():
    """An engine sized for 6 draws evaluates 1, 6, 2 draws in turn (ELBO with many
    draws, then a logger / convergence check with one)."""
    import dataclasses

    from oracle import treelik as orc
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    full = make_problem(21, 150, 4, 4, draws=6, seed=51)
    eng = Engine(full.tip_states, full.weights, full.postorder, 4, 4, max_draws=6)
    evec, ivec, evals = reversible_eigensystem(torch.tensor(full.q_matrix),
                                               torch.tensor(full.freqs))
    for d in (1, 6, 2):
        prob = dataclasses.replace(full, branch_lengths=full.branch_lengths[:d].copy())
        lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props,
                               evec, ivec, evals, prob.freqs)
        g = eng.grad_eigen()
        want = orc.evaluate(prob, want_grad=True, through_q=True)
        assert_lnl_close(lnl.numpy(), want["lnL"])
        assert_grad_close(g["branch_lengths"].numpy(), want["branch_lengths"], what="d_bl")
        assert_grad_close(g["site_rates"].numpy(), want["site_rates"], what="d_rates")
    eng.close()


@pytest.mark.parametrize("S,K", [(20, 16), (20, 17), (4, 16), (4, 11)])
def test_many_categories(S, K):
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(9, 70, S, K, seed=61, weibull_shape=1.5), q_rtol=1e-6)


def _with_code_table(prob, C, rng):
    """Give the tips an ambiguity alphabet of C codes: unit vectors, all-ones, then random
    0/1 masks (as `use_ambiguities` produces, datatype.py:78-97)."""
    S = prob.state_count
    table = [np.eye(S)[s] for s in range(S)] + [np.ones(S)]
    while len(table) < C:
        m = (rng.random(S) < 0.5).astype(np.float64)
        if m.sum() >= 2 and not any(np.array_equal(m, t) for t in table):
            table.append(m)
    prob.code_partials = np.array(table)
    extra = rng.integers(S, C, size=prob.tip_states.shape)
    use = rng.random(prob.tip_states.shape) < 0.15
    prob.tip_states = np.where(use, extra, prob.tip_states).astype(np.uint8)
    return prob


@pytest.mark.parametrize("flags", [0, 64], ids=["levels", "nocherry"])
@pytest.mark.parametrize("C", [6, 8, 9, 15])   # 15 = every 0/1 mask with >= 2 states + the unit vectors
def test_ambiguity_code_tables(C, flags):
    """Pair alphabets on both sides of the cherry-tabulation limit (C*C <= 64)."""
    from torchtree_b200.synthetic import make_problem

    rng = np.random.default_rng(C)
    prob = _with_code_table(make_problem(41, 200, 4, 4, seed=40 + C), C, rng)
    _check(prob, flags=flags)


def test_zero_weight_patterns_do_not_contribute():
    """Patterns masked with weight 0 (SRD06-style partitions, padding) drop out of lnL and
    of every gradient."""
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(24, 160, 4, 4, seed=77)
    masked = np.arange(0, 160, 7)
    prob.weights[masked] = 0.0
    eng, got = _run(prob)
    eng.close()
    keep = np.setdiff1d(np.arange(160), masked)
    sub = make_problem(24, 160, 4, 4, seed=77)
    sub.tip_states = np.ascontiguousarray(sub.tip_states[:, keep])
    sub.weights = np.ascontiguousarray(sub.weights[keep])
    sub.pattern_count = len(keep)
    eng2, want = _run(sub)
    eng2.close()
    assert_lnl_close(got["lnL"], want["lnL"])
    for k in ("branch_lengths", "site_rates", "props", "freqs", "q"):
        assert_grad_close(got[k], want[k], what=k)
    # a masked pattern the tree can hardly explain (one tip differs from all others on very
    # short branches) leaves everything finite
    prob.branch_lengths[:] = 1e-6
    prob.tip_states[0, masked[0]] = 0
    prob.tip_states[1:, masked[0]] = 1
    eng3, out = _run(prob)
    eng3.close()
    assert np.all(np.isfinite(out["lnL"]))
    assert all(np.all(np.isfinite(out[k])) for k in ("branch_lengths", "site_rates", "props", "freqs"))


def test_two_engines_interleaved():
    """Engines own their buffers: interleaving forward / backward calls of two engines (the
    two SRD06 partitions of one model) gives what each gives alone."""
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    probs = [make_problem(30, 150, 4, 4, seed=5), make_problem(30, 90, 4, 2, seed=6)]
    alone = []
    for p in probs:
        e, out = _run(p)
        e.close()
        alone.append(out)
    engines, args = [], []
    for p in probs:
        engines.append(Engine(p.tip_states, p.weights, p.postorder, 4, p.category_count,
                              code_partials=p.code_partials, max_draws=1))
        evec, ivec, evals = reversible_eigensystem(torch.tensor(p.q_matrix), torch.tensor(p.freqs))
        args.append((p.branch_lengths, p.site_rates, p.site_props, evec, ivec, evals, p.freqs))
    l0 = engines[0].loglik_eigen(*args[0])
    l1 = engines[1].loglik_eigen(*args[1])
    g1 = engines[1].grad_eigen()
    g0 = engines[0].grad_eigen()
    for lnl, g, want in ((l0, g0, alone[0]), (l1, g1, alone[1])):
        assert np.array_equal(lnl.numpy(), want["lnL"])
        assert np.array_equal(g["branch_lengths"].numpy(), want["branch_lengths"])
    for e in engines:
        e.close()


@pytest.fixture
def chain_mode_for_small_problems(monkeypatch):
    """Chain launches (runs of sparse levels walked by one launch per sweep) are reserved for
    wide pattern axes; switch them on for test-sized problems."""
    monkeypatch.setenv("TTB2_CHAIN_MIN_PATTERNS", "0")
    yield
    monkeypatch.delenv("TTB2_CHAIN_MIN_PATTERNS")


@pytest.mark.parametrize("flags", [0, 64], ids=["levels", "nocherry"])
@pytest.mark.parametrize("topology,T", [("caterpillar", 120), ("random", 200), ("balanced", 64)])
@pytest.mark.parametrize("K", [1, 4])
def test_chain_launches(chain_mode_for_small_problems, topology, T, K, flags):
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(T, 300, 4, K, seed=T + K, topology=topology, gap_fraction=0.03), flags=flags)


def test_chain_launches_draws_and_set_postorder(chain_mode_for_small_problems):
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    _check(make_problem(60, 130, 4, 4, draws=3, seed=9, topology="caterpillar", per_draw_model=True))
    # a topology change re-plans the runs
    a = make_problem(50, 200, 4, 4, seed=21, topology="caterpillar")
    b = make_problem(50, 200, 4, 4, seed=21, topology="random")
    b.tip_states, b.weights = a.tip_states, a.weights
    eng = Engine(a.tip_states, a.weights, a.postorder, 4, 4, max_draws=1)
    out = []
    for prob in (a, b, a):
        eng.set_postorder(prob.postorder)
        evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix), torch.tensor(prob.freqs))
        lnl = eng.loglik_eigen(prob.branch_lengths, prob.site_rates, prob.site_props, evec, ivec,
                               evals, prob.freqs)
        out.append((lnl.numpy().copy(), eng.grad_eigen()["branch_lengths"].numpy().copy()))
    eng.close()
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
    from oracle import treelik as orc
    want = orc.evaluate(b, want_grad=True)
    assert_lnl_close(out[1][0], want["lnL"])
    assert_grad_close(out[1][1], want["branch_lengths"], what="d_bl after set_postorder")


@pytest.mark.parametrize("gaps,ambiguous", [(0.0, False), (0.1, False), (0.1, True)],
                         ids=["unit", "gaps", "ambiguity-codes"])
def test_codon_sweeps_with_kept_u_vectors(gaps, ambiguous, monkeypatch):
    """61 states: the first evaluation of an engine repeats u_c = P_c p_c in the pre-order sweep;
    once the gradient buffers exist the post-order sweep keeps them and the pre-order sweep reads
    them back (kernels_gmma.cu gm_bwd3_kernel<61, true>).  Every evaluation of a sequence with
    changing parameters must equal the oracle, and equal an engine that never keeps u bit for bit
    (the same products in the same order, only computed once)."""
    import dataclasses

    from oracle import treelik as orc
    from torchtree_b200 import Engine, reversible_eigensystem
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(23, 150, 61, 2, seed=9, gap_fraction=gaps)
    if ambiguous:   # a third kind of code: the tip tables no longer are unit vectors + gap only
        table = np.concatenate([np.eye(61), np.ones((1, 61)), np.zeros((1, 61))])
        table[-1, [3, 17, 44]] = 1.0
        tips = prob.tip_states.copy()
        tips[::3, ::7] = table.shape[0] - 1
        prob = dataclasses.replace(prob, code_partials=table, tip_states=tips)

    def engine():
        return Engine(prob.tip_states, prob.weights, prob.postorder, 61, 2,
                      code_partials=prob.code_partials, max_draws=1)

    evec, ivec, evals = reversible_eigensystem(torch.tensor(prob.q_matrix), torch.tensor(prob.freqs))
    rng = np.random.default_rng(4)
    kept = engine()
    monkeypatch.setenv("TTB2_GM_NO_USTORE", "1")
    plain = engine()
    base_bytes = None
    for trip in range(3):
        bl = prob.branch_lengths * rng.uniform(0.5, 1.5, prob.branch_lengths.shape)
        bl[..., -1] = 0.0
        outs = []
        for eng in (kept, plain):
            if eng is kept:
                monkeypatch.delenv("TTB2_GM_NO_USTORE", raising=False)
            else:
                monkeypatch.setenv("TTB2_GM_NO_USTORE", "1")
            lnl = eng.loglik_eigen(bl, prob.site_rates, prob.site_props, evec, ivec, evals,
                                   prob.freqs).numpy().copy()
            g = {k: v.numpy().copy() for k, v in eng.grad_eigen().items()}
            outs.append((lnl, g))
        want = orc.evaluate(dataclasses.replace(prob, branch_lengths=bl), want_grad=True,
                            through_q=True)
        assert_lnl_close(outs[0][0], want["lnL"])
        assert_grad_close(outs[0][1]["branch_lengths"], want["branch_lengths"], what="d_bl")
        assert_grad_close(outs[0][1]["site_rates"], want["site_rates"], what="d_rates")
        assert np.array_equal(outs[0][0], outs[1][0])
        for key in outs[0][1]:
            assert np.array_equal(outs[0][1][key], outs[1][1][key]), (trip, key)
        if trip == 0:
            base_bytes = (kept.device_bytes, plain.device_bytes)
    # the kept-u engine holds one more buffer of the size of the partials
    assert base_bytes[0] > base_bytes[1]
    kept.close(); plain.close()
