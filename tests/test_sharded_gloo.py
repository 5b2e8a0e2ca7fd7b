"""Pattern sharding + collective logic on CPU: world_size 2, gloo backend.
The per-shard evaluator is the pinned oracle (no GPU here); what is tested is
shard_range, the all-reduce pair and that every rank ends with the full value
and the full gradient."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dataclasses

    from oracle import treelik as orc
    from torchtree_b200.sharded import shard_range, sharded_log_likelihood
    from torchtree_b200.synthetic import make_problem

    prob = make_problem(20, 101, 4, 4, seed=9, gap_fraction=0.03)
    lo, hi = shard_range(prob.pattern_count, rank, world)
    sub = dataclasses.replace(prob, pattern_count=hi - lo,
                              tip_states=prob.tip_states[:, lo:hi].copy(),
                              weights=prob.weights[lo:hi].copy())
    tips = orc.tip_partials_from_states(sub.tip_states, 4)
    w = torch.tensor(sub.weights)

    def local(bl, rates, props, q, freqs):
        t = bl.unsqueeze(-1) * rates.unsqueeze(-2)
        mats = orc.p_t_expm(q, t)
        return orc.log_likelihood(tips, w, sub.postorder, mats, freqs.unsqueeze(-2),
                                  props[..., None, None]).squeeze(-1)

    names = ("branch_lengths", "site_rates", "site_props", "q_matrix", "freqs")
    tensors = [torch.tensor(getattr(prob, n), requires_grad=True) for n in names]
    lnl = sharded_log_likelihood(local, tensors)
    lnl.sum().backward()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lnL=lnl.detach().numpy(),
             **{n: t.grad.numpy() for n, t in zip(names, tensors)})
    dist.destroy_process_group()


def test_pattern_sharding_two_ranks(tmp_path):
    from oracle import treelik as orc
    from torchtree_b200.synthetic import make_problem

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    prob = make_problem(20, 101, 4, 4, seed=9, gap_fraction=0.03)
    want = orc.evaluate(prob, want_grad=True, route="expm")
    wanted = {"branch_lengths": want["branch_lengths"], "site_rates": want["site_rates"],
              "site_props": want["site_props"], "q_matrix": want["q_matrix"],
              "freqs": want["freqs"]}
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert abs(got["lnL"][0] - want["lnL"][0]) <= 1e-11 * abs(want["lnL"][0])
        for k, v in wanted.items():
            np.testing.assert_allclose(got[k], v, rtol=1e-9, atol=1e-9 * np.abs(v).max(), err_msg=k)


def test_shard_range_covers_everything():
    from torchtree_b200.sharded import shard_range

    for n in (1, 7, 100, 100_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(hi >= lo for lo, hi in spans)


def _draw_worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import treelik as orc
    from torchtree_b200.sharded import draw_sharded_log_likelihood
    from torchtree_b200.synthetic import make_problem

    # 5 draws on 2 ranks (3 + 2): per-draw branch lengths / generator / frequencies, shared site model
    prob = make_problem(12, 40, 4, 3, draws=5, seed=21, per_draw_model=True)
    tips = orc.tip_partials_from_states(prob.tip_states, 4)
    w = torch.tensor(prob.weights)

    def local(bl, rates, props, q, freqs):
        d = bl.shape[0]
        t = bl.unsqueeze(-1) * rates.reshape(-1, 1, rates.shape[-1])
        mats = orc.p_t_expm(q, t)
        return orc.log_likelihood(tips, w, prob.postorder, mats, freqs.expand(d, -1).unsqueeze(-2),
                                  props.expand(d, -1)[..., None, None]).squeeze(-1)

    names = ("branch_lengths", "site_rates", "site_props", "q_matrix", "freqs")
    vals = [getattr(prob, n) for n in names]
    vals[1], vals[2] = vals[1][:1], vals[2][:1]          # the site model is shared by all draws
    tensors = [torch.tensor(np.ascontiguousarray(v), requires_grad=True) for v in vals]
    lnl = draw_sharded_log_likelihood(local, tensors, draws=5)
    wts = torch.tensor([1.0, -0.5, 2.0, 0.25, 1.5])
    (lnl * wts).sum().backward()
    np.savez(os.path.join(out_dir, "draw_rank%d.npz" % rank), lnL=lnl.detach().numpy(),
             **{n: t.grad.numpy() for n, t in zip(names, tensors)})
    dist.destroy_process_group()


def test_draw_sharding_two_ranks(tmp_path):
    """Config 3's sharding: draws split 3 + 2 over two ranks; every rank ends with lnL of all draws
    and the full gradients (per-draw tensors assembled, shared tensors summed)."""
    from oracle import treelik as orc
    from torchtree_b200.synthetic import make_problem

    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_draw_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    prob = make_problem(12, 40, 4, 3, draws=5, seed=21, per_draw_model=True)
    tips = orc.tip_partials_from_states(prob.tip_states, 4)
    names = ("branch_lengths", "site_rates", "site_props", "q_matrix", "freqs")
    vals = [getattr(prob, n) for n in names]
    vals[1], vals[2] = vals[1][:1], vals[2][:1]
    tensors = [torch.tensor(np.ascontiguousarray(v), requires_grad=True) for v in vals]
    bl, rates, props, q, freqs = tensors
    t = bl.unsqueeze(-1) * rates.reshape(-1, 1, rates.shape[-1])
    lnl = orc.log_likelihood(tips, torch.tensor(prob.weights), prob.postorder, orc.p_t_expm(q, t),
                             freqs.unsqueeze(-2), props.expand(5, -1)[..., None, None]).squeeze(-1)
    (lnl * torch.tensor([1.0, -0.5, 2.0, 0.25, 1.5])).sum().backward()
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "draw_rank%d.npz" % rank))
        np.testing.assert_allclose(got["lnL"], lnl.detach().numpy(), rtol=1e-12)
        for n, tt in zip(names, tensors):
            want = tt.grad.numpy()
            assert got[n].shape == want.shape, n
            np.testing.assert_allclose(got[n], want, rtol=1e-10, atol=1e-10 * np.abs(want).max(),
                                       err_msg=n)


def _idle_rank_worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torchtree_b200.sharded import draw_sharded_log_likelihood

    x = torch.tensor([[1.5, 2.0]], dtype=torch.float64, requires_grad=True)   # one draw only
    shared = torch.tensor([[3.0]], dtype=torch.float64, requires_grad=True)
    calls = []

    def local(xl, sl):
        calls.append(xl.shape[0])
        return (xl ** 2).sum(-1) * sl.reshape(-1)

    out = draw_sharded_log_likelihood(local, [x, shared], draws=1)
    out.sum().backward()
    np.savez(os.path.join(out_dir, "idle_rank%d.npz" % rank), out=out.detach().numpy(),
             gx=x.grad.numpy(), gs=shared.grad.numpy(), calls=np.array(calls))
    dist.destroy_process_group()


def test_draw_sharding_with_more_ranks_than_draws(tmp_path):
    """One draw on two ranks: the idle rank evaluates nothing but joins both collectives and ends
    with the same value and gradients."""
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_idle_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "idle_rank%d.npz" % rank))
        np.testing.assert_allclose(got["out"], [(1.5 ** 2 + 2.0 ** 2) * 3.0])
        np.testing.assert_allclose(got["gx"], [[2 * 1.5 * 3.0, 2 * 2.0 * 3.0]])
        np.testing.assert_allclose(got["gs"], [[1.5 ** 2 + 2.0 ** 2]])
        assert got["calls"].tolist() == ([1] if rank == 0 else [])


# ---- the drop-in class with its "shard" JSON key -------------------------
def _model_worker(rank, world, port, out_dir, shard):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank), TTB200_DIST_BACKEND="gloo")
    import refenv

    refenv.activate()
    torch.set_default_dtype(torch.float64)
    import test_plugin_reference as tp
    import torchtree_b200.flatten as flatten
    import torchtree_b200.tree_likelihood as tlmod

    flatten.log_likelihood_eigen = tp.fake_eigen      # no GPU here: the pinned oracle evaluates
    tlmod.Engine = tp.FakeEngine                      # this rank's shard
    objs, like = tp._flu_json(batch=3 if shard == "draws" else None)
    like = dict(like, shard=shard)
    dic = tp._build(objs, like, "torchtree_b200.TreeLikelihoodModel")
    model = dic["like"]
    assert model._world == world and model.shard == shard
    if shard == "patterns":
        lo, hi = model.pattern_range
        assert (hi - lo) in (119, 119) and model._get_engine(1).tip_codes.shape[1] == hi - lo
    val, grads = tp._grads(dic, ["blens", "shape", "rates", "freqs"])
    np.savez(os.path.join(out_dir, "model_rank%d.npz" % rank), lnL=val.numpy(),
             **{k: v.numpy() for k, v in grads.items()})
    dist.destroy_process_group()


@pytest.mark.reference
@pytest.mark.parametrize("shard", ["patterns", "draws"])
def test_model_shard_key_two_ranks(tmp_path, shard):
    """`"shard": "patterns" | "draws"` on the drop-in class: two gloo ranks, each with its slice of
    the 238 fluA patterns (or of a batch of 3 draws), end up with the reference's value and the full
    gradient of every torchtree Parameter."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import refenv

    refenv.activate()
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        import test_plugin_reference as tp
        import torchtree_b200.tree_likelihood  # noqa: F401  (imports and registers torchtree's classes)

        port = 31500 + (os.getpid() % 2000) + (7 if shard == "draws" else 0)
        mp.spawn(_model_worker, args=(2, port, str(tmp_path), shard), nprocs=2, join=True)
        objs, like = tp._flu_json(batch=3 if shard == "draws" else None)
        ref = tp._build(objs, like, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
        v_ref, g_ref = tp._grads(ref, ["blens", "shape", "rates", "freqs"])
    finally:
        torch.set_default_dtype(old)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "model_rank%d.npz" % rank))
        np.testing.assert_allclose(got["lnL"], v_ref.numpy(), rtol=1e-10)
        for k, v in g_ref.items():
            tol = 1e-7 if k in ("rates", "freqs") else 1e-8
            np.testing.assert_allclose(got[k], v.numpy(), rtol=tol, atol=tol * float(v.abs().max()),
                                       err_msg=k)
