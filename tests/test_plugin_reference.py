"""The torchtree plug-in class against the REAL reference, in two set-ups:

* `backend = "cuda"` (`-m gpu`, the B200 box): `torchtree_b200.TreeLikelihoodModel` + the real
  CUDA engine (libttb200.so through the torch extension) + the real torchtree, vendored into
  `baseline/_ref` by tools/vendor_reference.py -- the product as a user runs it.  Every test
  builds the same JSON twice (reference class / drop-in class) and compares lnL at 1e-10 and every
  `Parameter.grad` at 1e-8 (substitution parameters whose reference gradient goes through the
  `eigh` backward: 1e-7; tests/test_truth_mpmath_gpu.py attributes that slack to the
  reference), and runs `torchtree-cli advi|hmc|map|mcmc ... --b200` + the stock runner.
* `backend = "oracle"` (`-m "not gpu"`, the authoring container, no GPU): the two engine entry
  points the glue calls are replaced by the pinned CPU oracle (test-only injection); what is
  verified is everything *around* the engine: JSON parsing, class registration / type
  resolution, tip-code extraction, the flattening of torchtree's sub-models, batch shapes, the
  output contract and the gradient hand-back to torchtree Parameters."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.reference

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402

DATA = refenv.data_dir() or "/root/reference/data"
BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module")
def torchtree_env():
    added = refenv.activate()
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)
    for p in added:
        if p != REPO:
            sys.path.remove(p)


class FakeEngine:
    """Carries what Engine() was given; the oracle does the arithmetic."""

    def __init__(self, tip_codes, weights, postorder, state_count, category_count,
                 code_partials=None, max_draws=1, device=0, flags=0):
        from torchtree_b200.engine import default_code_partials

        self.tip_codes, self.weights, self.postorder = tip_codes, weights, postorder
        self.S, self.K, self.max_draws = state_count, category_count, max_draws
        self.code_partials = default_code_partials(state_count) if code_partials is None \
            else code_partials

    def set_postorder(self, postorder):
        self.postorder = postorder

    def close(self):
        pass

    def release(self):
        pass


def _tips(engine):
    from oracle import treelik as orc

    return orc.tip_partials_from_states(engine.tip_codes, engine.S, engine.code_partials)


def fake_eigen(engine, bls, rates, props, q, freqs):
    from oracle import treelik as orc

    D = bls.shape[0]
    t = bls.unsqueeze(-1) * rates.reshape(-1, 1, engine.K)
    # same contract as the engine: P depends on Q only, all S*S entries of Q
    # independent (matrix_exp autograd), freqs enter at the root only
    mats = orc.p_t_expm(q.expand(max(q.shape[0], 1), -1, -1), t)
    return orc.log_likelihood(
        _tips(engine), torch.tensor(engine.weights), engine.postorder, mats,
        freqs.expand(D, -1).unsqueeze(-2), props.expand(D, -1)[..., None, None]).squeeze(-1)


def fake_mats(engine, mats, freqs, props):
    from oracle import treelik as orc

    D = mats.shape[0]
    return orc.log_likelihood(
        _tips(engine), torch.tensor(engine.weights), engine.postorder, mats,
        freqs.expand(D, -1).unsqueeze(-2), props.expand(D, -1)[..., None, None]).squeeze(-1)


@pytest.fixture(params=BACKENDS)
def patched(request, torchtree_env, monkeypatch):
    """The drop-in module, backed by the CUDA engine ("cuda") or by the oracle ("oracle")."""
    import torchtree_b200.flatten as flatten
    import torchtree_b200.tree_likelihood as tlmod

    monkeypatch.setattr(tlmod, "BACKEND", request.param, raising=False)
    if request.param == "oracle":
        monkeypatch.setattr(flatten, "log_likelihood_eigen", fake_eigen)
        monkeypatch.setattr(flatten, "log_likelihood_expm", fake_eigen)
        monkeypatch.setattr(flatten, "log_likelihood_mats", fake_mats)
        monkeypatch.setattr(tlmod, "Engine", FakeEngine)
    else:
        assert torch.cuda.is_available(), "the cuda backend needs a GPU"
        from torchtree_b200 import _lib

        _lib.load()  # the product path: fail loudly when the CUDA library is missing
    return tlmod


def _grad_tol(name):
    """north_star: gradients at 1e-8; parameters whose *reference* gradient flows through the
    eigh backward (1/eigen-gap error amplification) at 1e-7."""
    return 1e-7 if name in ("rates", "freqs", "kappa") else 1e-8


def _assert_grads(g_new, g_ref, names):
    for n in names:
        assert g_new[n].shape == g_ref[n].shape, n
        tol = _grad_tol(n)
        assert torch.allclose(g_new[n], g_ref[n], rtol=tol, atol=tol * g_ref[n].abs().max()), \
            (n, (g_new[n] - g_ref[n]).abs().max().item(), g_ref[n].abs().max().item())


def _flu_json(tree_type="unrooted", model="GTR", batch=None):
    taxa = []
    with open(DATA + "/fluA.fa") as fp:
        for line in fp:
            if line.startswith(">"):
                name = line[1:].strip()
                taxa.append({"id": name, "type": "Taxon",
                             "attributes": {"date": float(name.split("_")[-1])}})
    with open(DATA + "/fluA.tree") as fp:
        newick = fp.read().strip()
    rng = np.random.default_rng(5)
    T = len(taxa)
    shape = (2 * T - 3,) if batch is None else (batch, 2 * T - 3)
    objs = [
        {"id": "taxa", "type": "Taxa", "taxa": taxa},
        {"id": "alignment", "type": "Alignment", "datatype": "nucleotide",
         "file": DATA + "/fluA.fa", "taxa": "taxa"},
    ]
    like = {
        "id": "like", "type": "TreeLikelihoodModel",
        "tree_model": {"id": "tree", "type": "UnRootedTreeModel", "newick": newick,
                       "taxa": "taxa",
                       "branch_lengths": {"id": "blens", "type": "Parameter",
                                          "tensor": rng.uniform(0.01, 0.1, shape).tolist()}},
        "site_model": {"id": "sm", "type": "WeibullSiteModel", "categories": 4,
                       "shape": {"id": "shape", "type": "Parameter",
                                 "tensor": [0.7] if batch is None else [[0.7]] * batch}},
        "site_pattern": {"id": "sp", "type": "SitePattern", "alignment": "alignment"},
    }
    if model == "GTR":
        like["substitution_model"] = {
            "id": "gtr", "type": "GTR",
            "rates": {"id": "rates", "type": "Parameter",
                      "tensor": [0.9, 3.1, 0.6, 1.3, 4.2, 1.0]},
            "frequencies": {"id": "freqs", "type": "Parameter",
                            "tensor": [0.33, 0.19, 0.22, 0.26]}}
    elif model == "JC69":
        like["substitution_model"] = {"id": "jc", "type": "JC69"}
    return objs, like


def _build(objs, like, like_type):
    from torchtree.core.utils import process_objects

    dic = {}
    for o in objs:
        process_objects(json.loads(json.dumps(o)), dic)
    data = json.loads(json.dumps(like))
    data["type"] = like_type
    process_objects(data, dic)
    return dic


def _grads(dic, names):
    for n in names:
        dic[n].requires_grad = True
    val = dic["like"]()
    val.sum().backward()
    return val.detach().clone(), {n: dic[n].grad.clone() for n in names}


@pytest.mark.parametrize("batch", [None, 3])
def test_dropin_equals_reference_on_fluA_gtr(patched, batch):
    objs, like = _flu_json(batch=batch)
    ref = _build(objs, like, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
    new = _build(objs, like, "torchtree_b200.TreeLikelihoodModel")
    assert type(new["like"]).__module__ == "torchtree_b200.tree_likelihood"
    names = ["blens", "shape", "rates", "freqs"]
    v_ref, g_ref = _grads(ref, names)
    v_new, g_new = _grads(new, names)
    assert v_new.shape == v_ref.shape  # sample_shape + (1,)
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0)
    _assert_grads(g_new, g_ref, names)


def test_tip_states_and_ambiguities_flags(patched):
    objs, like = _flu_json()
    for extra in ({"use_tip_states": True}, {"use_ambiguities": True}):
        cfg = dict(like, **extra)
        ref = _build(objs, cfg, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
        new = _build(objs, cfg, "torchtree_b200.TreeLikelihoodModel")
        assert torch.allclose(new["like"](), ref["like"](), rtol=1e-10, atol=0), extra


def test_install_overrides_bare_and_dotted_type_names(patched):
    from torchtree.core.utils import REGISTERED_CLASSES, get_class

    import torchtree.evolution.tree_likelihood as refmod

    saved_reg = dict(REGISTERED_CLASSES)
    saved_cls = refmod.TreeLikelihoodModel
    try:
        patched.install(override_reference=True)
        assert get_class("TreeLikelihoodModel") is patched.TreeLikelihoodModel
        assert get_class("torchtree.evolution.tree_likelihood.TreeLikelihoodModel") \
            is patched.TreeLikelihoodModel
        assert get_class("torchtree_b200.TreeLikelihoodModel") is patched.TreeLikelihoodModel
        assert get_class("LG").__name__ == "LG"
        objs, like = _flu_json(model="JC69")
        dic = _build(objs, like, "TreeLikelihoodModel")
        assert isinstance(dic["like"], patched.TreeLikelihoodModel)
        assert dic["like"]().shape == (1,)
    finally:
        REGISTERED_CLASSES.update(saved_reg)
        refmod.TreeLikelihoodModel = saved_cls


def test_cli_plugin_rewrites_type(torchtree_env):
    import argparse

    from torchtree.cli.plugin_manager import PluginManager

    import torchtree_b200

    sys.path.insert(0, REPO)
    pm = PluginManager()
    pm.load_plugins()
    assert "torchtree_b200" in pm._plugins
    parser = argparse.ArgumentParser()
    sub = parser.add_subparsers()
    for name in ("advi", "hmc", "map", "mcmc"):
        sub.add_parser(name)
    pm.load_arguments(sub)
    args = parser.parse_args(["advi", "--b200"])
    data = {"id": "like", "type": "TreeLikelihoodModel"}
    for plugin in pm.plugins():
        plugin.process_tree_likelihood(args, data)
    assert data["type"] == "torchtree_b200.TreeLikelihoodModel"
    args = parser.parse_args(["map"])
    data = {"id": "like", "type": "TreeLikelihoodModel"}
    for plugin in pm.plugins():
        plugin.process_tree_likelihood(args, data)
    assert data["type"] == "TreeLikelihoodModel"


def test_time_tree_with_strict_clock(patched):
    """test/test_tree_likelihood.py:268-342 through the drop-in class."""
    from torchtree import Parameter
    from torchtree.evolution.branch_model import StrictClockModel
    from torchtree.evolution.site_model import WeibullSiteModel
    from torchtree.evolution.site_pattern import SitePattern
    from torchtree.evolution.substitution_model import JC69
    from torchtree.evolution.taxa import Taxa, Taxon
    from torchtree.evolution.tree_model import ReparameterizedTimeTreeModel

    taxa_list = []
    with open(DATA + "/fluA.fa") as fp:
        for line in fp:
            if line.startswith(">"):
                t = line[1:].strip()
                taxa_list.append(Taxon(t, {"date": float(t.split("_")[-1])}))
    dic = {"taxa": Taxa("taxa", taxa_list)}
    with open(DATA + "/fluA.tree") as fp:
        newick = fp.read().strip()
    tree_model = ReparameterizedTimeTreeModel.from_json(
        ReparameterizedTimeTreeModel.json_factory(
            "tree_model", newick, "taxa", ratios=[0.5] * 67, root_height=[20.0],
            **{"keep_branch_lengths": True}), dic)
    sp = SitePattern.from_json({
        "id": "sp", "type": "SitePattern",
        "alignment": {"id": "a", "type": "Alignment", "datatype": "nucleotide",
                      "file": DATA + "/fluA.fa", "taxa": "taxa"}}, dic)
    site_model = WeibullSiteModel("sm", Parameter(None, torch.tensor([[0.1]])), 4)
    clock = StrictClockModel(None, Parameter(None, torch.tensor([[0.001]])), tree_model)
    like = patched.TreeLikelihoodModel("like", sp, tree_model, JC69("jc"), site_model, clock)
    assert torch.allclose(torch.tensor([-4618.2062529058]), like())
    clock._rates.tensor = clock._rates.tensor.repeat(3, 1)
    site_model._parameter.tensor = site_model._parameter.tensor.repeat(3, 1)
    tree_model._internal_heights.tensor = tree_model._internal_heights.tensor.repeat(3, 68)
    like.lp_needs_update = True
    assert torch.allclose(torch.tensor([[-4618.2062529058] * 3]), like())


def test_install_height_transform_rebinds_and_fails_loudly_without_gpu(patched):
    """install(height_transform=True): ReparameterizedTimeTreeModel picks up the device
    transform; without a GPU it must raise, not fall back to the Python loop."""
    from torchtree.core.utils import REGISTERED_CLASSES
    from torchtree.evolution.taxa import Taxa, Taxon

    import torchtree.evolution.tree_height_transform as ref_transform
    import torchtree.evolution.tree_likelihood as refmod
    import torchtree.evolution.tree_model as ref_tree_model
    from torchtree_b200._lib import EngineError
    from torchtree_b200.height_transform import GeneralNodeHeightTransform

    saved_reg = dict(REGISTERED_CLASSES)
    saved = (refmod.TreeLikelihoodModel, ref_transform.GeneralNodeHeightTransform,
             ref_tree_model.GeneralNodeHeightTransform)
    try:
        patched.install(override_reference=True, height_transform=True)
        assert ref_tree_model.GeneralNodeHeightTransform is GeneralNodeHeightTransform
        assert ref_transform.GeneralNodeHeightTransform is GeneralNodeHeightTransform
        assert ref_tree_model.ReferenceGeneralNodeHeightTransform is saved[2]
        taxa = Taxa("taxa", [Taxon(n, {"date": d}) for n, d in
                             (("A", 0.0), ("B", 1.0), ("C", 2.0))])
        build = lambda: ref_tree_model.ReparameterizedTimeTreeModel.from_json(  # noqa: E731
            ref_tree_model.ReparameterizedTimeTreeModel.json_factory(
                "tree", "((A:1,B:1):1,C:1);", "taxa", ratios=[0.5], root_height=[5.0]),
            {"taxa": taxa})
        if torch.cuda.is_available():
            assert build().node_heights.shape[-1] == 5
        else:
            with pytest.raises(EngineError, match="no CUDA device"):
                build().branch_lengths()
    finally:
        REGISTERED_CLASSES.update(saved_reg)
        refmod.TreeLikelihoodModel = saved[0]
        ref_transform.GeneralNodeHeightTransform = saved[1]
        ref_tree_model.GeneralNodeHeightTransform = saved[2]


# ---- the variants existing configs contain (SURVEY 8(b) "Variants the drop-in will meet") -------
_P = lambda id_, v: {"id": id_, "type": "Parameter", "tensor": v}  # noqa: E731
_SUBST = {
    "HKY": {"id": "hky", "type": "HKY", "kappa": _P("kappa", [3.2]),
            "frequencies": _P("freqs", [0.3, 0.2, 0.24, 0.26])},
    "GTR": {"id": "gtr", "type": "GTR", "rates": _P("rates", [0.9, 3.1, 0.6, 1.3, 4.2, 1.0]),
            "frequencies": _P("freqs", [0.33, 0.19, 0.22, 0.26])},
    # matrix_exp route: the model's own p_t supplies the matrices (abstract.py:89-94)
    "NONSYM": {"id": "ns", "type": "GeneralNonSymmetricSubstitutionModel",
               "data_type": {"id": "dt", "type": "NucleotideDataType"},
               "mapping": list(range(12)),
               "rates": _P("rates", [0.5, 1.7, 0.8, 1.1, 2.3, 0.4, 0.9, 1.4, 0.7, 1.9, 0.6, 1.2]),
               "frequencies": _P("freqs", [0.28, 0.22, 0.24, 0.26])},
}
_SITE = {
    "constant": ({"id": "sm", "type": "ConstantSiteModel"}, []),
    "constant_mu": ({"id": "sm", "type": "ConstantSiteModel", "mu": _P("mu", [1.7])}, ["mu"]),
    "invariant": ({"id": "sm", "type": "InvariantSiteModel", "invariant": _P("pinv", [0.23])},
                  ["pinv"]),
    "weibull_inv_mu": ({"id": "sm", "type": "WeibullSiteModel", "categories": 3,
                        "shape": _P("shape", [0.8]), "invariant": _P("pinv", [0.15]),
                        "mu": _P("mu", [0.6])}, ["shape", "pinv", "mu"]),
}


@pytest.mark.parametrize("subst,site", [("HKY", "constant"), ("GTR", "invariant"),
                                        ("GTR", "weibull_inv_mu"), ("HKY", "constant_mu"),
                                        ("NONSYM", "weibull_inv_mu")])
def test_model_variants_match_reference(patched, subst, site):
    objs, like = _flu_json()
    like = dict(like, substitution_model=_SUBST[subst], site_model=_SITE[site][0])
    ref = _build(objs, like, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
    new = _build(objs, like, "torchtree_b200.TreeLikelihoodModel")
    names = ["blens", "freqs"] + _SITE[site][1] + (["kappa"] if subst == "HKY" else ["rates"])
    v_ref, g_ref = _grads(ref, names)
    v_new, g_new = _grads(new, names)
    assert v_new.shape == v_ref.shape
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0), (v_new, v_ref)
    _assert_grads(g_new, g_ref, names)


def test_srd06_two_likelihoods_share_a_tree(patched):
    """SRD06 (cli/evolution.py:630-671): codon positions 1+2 and 3 as two TreeLikelihoodModels
    over `SitePattern.indices`, one tree -> two engines."""
    objs, like = _flu_json()
    totals = {}
    for kind in ("torchtree.evolution.tree_likelihood.TreeLikelihoodModel",
                 "torchtree_b200.TreeLikelihoodModel"):
        dic = _build(objs, dict(like, id="like12",
                                site_pattern=dict(like["site_pattern"], id="sp12",
                                                  indices="::3,1::3")), kind)
        from torchtree.core.utils import process_objects

        second = json.loads(json.dumps(dict(
            like, id="like3", tree_model="tree", site_model="sm", substitution_model="gtr",
            site_pattern=dict(like["site_pattern"], id="sp3", indices="2::3"))))
        second["type"] = kind
        process_objects(second, dic)
        dic["blens"].requires_grad = True
        total = dic["like12"]() + dic["like3"]()
        total.sum().backward()
        totals[kind] = (total.detach().clone(), dic["blens"].grad.clone(),
                        dic["like12"].weights.sum().item() + dic["like3"].weights.sum().item())
    (v_ref, g_ref, n_ref), (v_new, g_new, n_new) = totals.values()
    assert n_ref == n_new == 987  # every site of fluA lands in exactly one partition
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0)
    assert torch.allclose(g_new, g_ref, rtol=1e-8, atol=1e-8 * g_ref.abs().max())


def test_amino_acid_lg_weibull(patched):
    """Empirical 20-state model (LG, amino_acid.py:12-57) on an inline protein alignment with
    gaps and ambiguity letters: the eigen route through `normalised_generator`."""
    rng = np.random.default_rng(12)
    names = ["t%d" % i for i in range(6)]
    letters = np.array(list("ACDEFGHIKLMNPQRSTVWY"))
    base = rng.integers(0, 20, 60)
    seqs = []
    for _ in names:
        s = base.copy()
        flip = rng.random(60) < 0.25
        s[flip] = rng.integers(0, 20, int(flip.sum()))
        chars = letters[s]
        chars[rng.random(60) < 0.05] = "-"
        chars[rng.random(60) < 0.03] = "X"
        seqs.append("".join(chars))
    objs = [
        {"id": "taxa", "type": "Taxa", "taxa": [{"id": n, "type": "Taxon"} for n in names]},
        {"id": "alignment", "type": "Alignment",
         "datatype": {"id": "aa", "type": "AminoAcidDataType"}, "taxa": "taxa",
         "sequences": [{"taxon": n, "sequence": s} for n, s in zip(names, seqs)]},
    ]
    like = {
        "id": "like", "type": "TreeLikelihoodModel",
        "tree_model": {"id": "tree", "type": "UnRootedTreeModel",
                       "newick": "((t0:0.1,t1:0.2):0.05,(t2:0.1,t3:0.3):0.1,(t4:0.2,t5:0.1):0.1);",
                       "taxa": "taxa",
                       "branch_lengths": _P("blens", rng.uniform(0.05, 0.4, 9).tolist())},
        "site_model": {"id": "sm", "type": "WeibullSiteModel", "categories": 4,
                       "shape": _P("shape", [0.6])},
        "substitution_model": {"id": "lg", "type": "torchtree.evolution.substitution_model.amino_acid.LG"},
        "site_pattern": {"id": "sp", "type": "SitePattern", "alignment": "alignment"},
    }
    ref = _build(objs, like, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
    new = _build(objs, like, "torchtree_b200.TreeLikelihoodModel")
    assert new["like"]._state_count == 20
    v_ref, g_ref = _grads(ref, ["blens", "shape"])
    v_new, g_new = _grads(new, ["blens", "shape"])
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0), (v_new, v_ref)
    _assert_grads(g_new, g_ref, ["blens", "shape"])


def _run_cli(argv, capsys):
    from torchtree.cli.cli import main as cli_main

    old = sys.argv
    sys.argv = ["torchtree-cli"] + argv
    try:
        capsys.readouterr()
        cli_main()
        return capsys.readouterr().out
    finally:
        sys.argv = old


def _run_torchtree(path, capsys, seed=1):
    from torchtree.torchtree import main as torchtree_main

    old = sys.argv
    sys.argv = ["torchtree", path, "-s", str(seed)]
    try:
        capsys.readouterr()
        torchtree_main()
        return capsys.readouterr().out
    finally:
        sys.argv = old


def test_cli_generated_advi_config_runs_unchanged(patched, tmp_path, capsys):
    """BASELINE config 1 end to end on the host side: `torchtree-cli advi -m GTR -C 4` writes the
    JSON, `--b200` only swaps the likelihood's type, and the stock runner optimises both with the
    same seed -- the ELBO traces must agree (the engine is the pinned oracle here; the CUDA engine
    meets the same oracle in the -m gpu tests)."""
    import re

    sys.path.insert(0, REPO)
    base = ["advi", "-i", DATA + "/fluA.fa", "-t", DATA + "/fluA.tree", "-m", "GTR", "-C", "4",
            "--iter", "6", "--elbo_samples", "3", "--grad_samples", "1", "--convergence_every", "2",
            "--stem", str(tmp_path / "run")]
    traces = {}
    for tag, extra in (("reference", []), ("b200", ["--b200"])):
        cfg = _run_cli(base + extra, capsys)
        data = json.loads(cfg)
        types = re.findall(r'"type": "([A-Za-z_.0-9]*TreeLikelihoodModel)"', cfg)
        assert types == (["torchtree_b200.TreeLikelihoodModel"] if extra else ["TreeLikelihoodModel"])
        path = tmp_path / (tag + ".json")
        path.write_text(json.dumps(data))
        out = _run_torchtree(str(path), capsys)
        # the runner prints a table: iter, ELBO, delta_ELBO_mean, delta_ELBO_med
        elbos = [float(m.group(1)) for m in
                 re.finditer(r"^\s*\d+\s+(-?\d+\.\d+)\s+\d+\.\d+\s+\d+\.\d+", out, flags=re.M)]
        assert len(elbos) >= 3, out[-2000:]
        traces[tag] = elbos
    ref, new = np.array(traces["reference"]), np.array(traces["b200"])
    assert ref.shape == new.shape
    # same seed, same draws: the traces agree to the accumulated round-off of six Adam steps
    np.testing.assert_allclose(new, ref, rtol=1e-7)


@pytest.mark.parametrize("command,extra", [
    ("map", ["--max_iter", "3"]),
    ("hmc", ["--iter", "4", "--steps", "3"]),
    ("mcmc", ["--iter", "20"]),
])
def test_other_cli_drivers_run_unchanged(patched, tmp_path, capsys, command, extra):
    """`torchtree-cli map / hmc / mcmc ... [--b200]` + the stock runner: the drivers see the same
    log-posterior values, accept the same proposals and print the same numbers."""
    import re

    sys.path.insert(0, REPO)
    outs = {}
    for tag, flag in (("reference", []), ("b200", ["--b200"])):
        argv = [command, "-i", DATA + "/fluA.fa", "-t", DATA + "/fluA.tree", "-m", "JC69",
                "--stem", str(tmp_path / (tag + "_run"))] + extra + flag
        cfg = _run_cli(argv, capsys)
        assert ("torchtree_b200.TreeLikelihoodModel" in cfg) == bool(flag)
        path = tmp_path / (tag + ".json")
        path.write_text(cfg)
        outs[tag] = _run_torchtree(str(path), capsys)
    num = re.compile(r"-?\d+\.\d+(?:e[-+]?\d+)?|nan")
    ref = [float(x) for x in num.findall(outs["reference"])]
    new = [float(x) for x in num.findall(outs["b200"])]
    assert len(ref) == len(new) and len(ref) >= 2, (outs["reference"][-800:], outs["b200"][-800:])
    np.testing.assert_allclose(new, ref, rtol=1e-6, equal_nan=True)


def test_device_coalescent_class_is_a_dropin(patched, monkeypatch):
    """`ConstantCoalescentModel` with the device log-density (coalescent.py): same JSON, same value
    and gradients as the reference class (the kernel is replaced by the pinned oracle here), bare /
    dotted type names rebound by install(coalescent=True), loud failure without a GPU."""
    from torchtree.core.utils import REGISTERED_CLASSES, get_class, process_objects

    import torchtree.evolution.coalescent as refmod

    import torchtree_b200.coalescent as cmod
    from oracle.coalescent import constant_log_prob

    rng = np.random.default_rng(3)
    T = 7
    tips = rng.uniform(0, 3, T)
    inner = tips.max() + np.cumsum(rng.exponential(0.5, T - 1))
    heights = np.concatenate([tips, inner])

    def build(type_name, batch):
        dic = {}
        data = {"id": "coal", "type": type_name,
                "theta": _P("theta", [4.0] if batch is None else [[4.0], [2.5], [7.0]]),
                "times": heights.tolist(), "events": [1] * T + [0] * (T - 1)}
        process_objects(json.loads(json.dumps(data)), dic)
        return dic

    saved_reg, saved_cls = dict(REGISTERED_CLASSES), refmod.ConstantCoalescentModel
    backend, install = patched.BACKEND, patched.install
    try:
        for batch in (None, 3):
            ref = build("torchtree.evolution.coalescent.ConstantCoalescentModel", batch)
            new = build("torchtree_b200.coalescent.ConstantCoalescentModel", batch)
            assert type(new["coal"]).__module__ == "torchtree_b200.coalescent"
            assert isinstance(new["coal"], refmod.ConstantCoalescentModel)
            if backend == "oracle":
                if not torch.cuda.is_available():
                    with pytest.raises(RuntimeError, match="no CUDA device"):
                        new["coal"]()
                monkeypatch.setattr(cmod, "constant_coalescent_log_prob",
                                    lambda h, th, device=0: constant_log_prob(h, th))
                new["coal"].lp_needs_update = True
            for dic in (ref, new):
                dic["theta"].requires_grad = True
                dic["coal"]().sum().backward()
            assert new["coal"]().shape == ref["coal"]().shape
            assert torch.allclose(new["coal"](), ref["coal"](), rtol=1e-12, atol=0)
            assert torch.allclose(new["theta"].grad, ref["theta"].grad, rtol=1e-10, atol=0)
            if backend == "oracle":
                monkeypatch.undo()
        install(override_reference=False, coalescent=True)
        assert get_class("ConstantCoalescentModel") is cmod.ConstantCoalescentModel
        assert get_class("torchtree.evolution.coalescent.ConstantCoalescentModel") \
            is cmod.ConstantCoalescentModel
    finally:
        REGISTERED_CLASSES.update(saved_reg)
        refmod.ConstantCoalescentModel = saved_cls
        for other in ("PiecewiseConstantCoalescentModel", "PiecewiseConstantCoalescentGridModel"):
            if hasattr(refmod, "Reference" + other):
                setattr(refmod, other, getattr(refmod, "Reference" + other))


def test_time_tree_with_per_branch_clock_rates(patched):
    """Relaxed clock (`SimpleClockModel`: one rate per branch, branch_model.py:58-80) on the
    reparameterised fluA time tree: value and gradients w.r.t. ratios, root height and rates equal the
    reference class's (the scaled branch lengths reach the engine through
    flatten.scaled_branch_lengths, tree_likelihood.py:323-344)."""
    from torchtree import Parameter
    from torchtree.evolution.branch_model import SimpleClockModel
    from torchtree.evolution.site_model import ConstantSiteModel
    from torchtree.evolution.site_pattern import SitePattern
    from torchtree.evolution.substitution_model import JC69
    from torchtree.evolution.taxa import Taxa, Taxon
    from torchtree.evolution.tree_likelihood import TreeLikelihoodModel as Reference
    from torchtree.evolution.tree_model import ReparameterizedTimeTreeModel

    taxa_list = []
    with open(DATA + "/fluA.fa") as fp:
        for line in fp:
            if line.startswith(">"):
                t = line[1:].strip()
                taxa_list.append(Taxon(t, {"date": float(t.split("_")[-1])}))
    with open(DATA + "/fluA.tree") as fp:
        newick = fp.read().strip()
    rng = np.random.default_rng(8)
    rates0 = rng.uniform(5e-4, 3e-3, 136)
    results = []
    for cls in (Reference, patched.TreeLikelihoodModel):
        dic = {"taxa": Taxa("taxa", taxa_list)}
        tree_model = ReparameterizedTimeTreeModel.from_json(
            ReparameterizedTimeTreeModel.json_factory(
                "tree_model", newick, "taxa", ratios=[0.5] * 67, root_height=[20.0],
                **{"keep_branch_lengths": True}), dic)
        sp = SitePattern.from_json({
            "id": "sp", "type": "SitePattern",
            "alignment": {"id": "a", "type": "Alignment", "datatype": "nucleotide",
                          "file": DATA + "/fluA.fa", "taxa": "taxa"}}, dic)
        rates = Parameter("rates", torch.tensor(rates0))
        clock = SimpleClockModel("clock", rates, tree_model)
        like = cls("like", sp, tree_model, JC69("jc"), ConstantSiteModel("sm"), clock)
        # ratios and root height are the two members of the CatParameter behind the tree
        leaves = [rates] + list(tree_model._internal_heights._parameter_container.parameters())
        assert len(leaves) == 3
        for p in leaves:
            p.requires_grad = True
        value = like()
        value.sum().backward()
        results.append((value.detach().clone(), [p.grad.clone() for p in leaves]))
    (v_ref, g_ref), (v_new, g_new) = results
    assert v_new.shape == v_ref.shape
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0), (v_new, v_ref)
    for a, b in zip(g_new, g_ref):
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-8, atol=1e-8 * b.abs().max())


def test_codon_mg94_on_fluA(patched):
    """61-state codon model (MG94, substitution_model/codon.py) on fluA read as codons (the
    shape of BASELINE config 5 on real data): value and gradients w.r.t. branch lengths, Weibull
    shape, kappa / alpha / beta and the 61 codon frequencies equal the reference class's."""
    from torchtree import Parameter
    from torchtree.evolution.alignment import Alignment, Sequence
    from torchtree.evolution.datatype import CodonDataType
    from torchtree.evolution.site_model import WeibullSiteModel
    from torchtree.evolution.site_pattern import SitePattern
    from torchtree.evolution.substitution_model import MG94
    from torchtree.evolution.taxa import Taxa, Taxon
    from torchtree.evolution.tree_likelihood import TreeLikelihoodModel as Reference
    from torchtree.evolution.tree_model import UnRootedTreeModel

    # MG94.handle_parameter_changed calls a method that does not exist; patched in
    # the harness (as tests/golden/make_golden.py does), never in the reference
    MG94.handle_parameter_changed = lambda self, v, i, e: self.fire_model_changed()
    names, seqs = [], []
    with open(DATA + "/fluA.fa") as fp:
        for line in fp:
            line = line.strip()
            if line.startswith(">"):
                names.append(line[1:])
                seqs.append("")
            elif line:
                seqs[-1] += line
    with open(DATA + "/fluA.tree") as fp:
        newick = fp.read().strip()
    rng = np.random.default_rng(17)
    bl0 = rng.uniform(0.005, 0.05, 2 * len(names) - 3)
    f0 = rng.dirichlet(np.full(61, 20.0))
    results = []
    for cls in (Reference, patched.TreeLikelihoodModel):
        datatype = CodonDataType("codon", "Universal")
        taxa = Taxa("taxa", [Taxon(n, None) for n in names])
        aln = Alignment("aln", [Sequence(n, s) for n, s in zip(names, seqs)], taxa, datatype)
        sp = SitePattern("sp", aln)
        dic = {"taxa": taxa, "blens": Parameter("blens", torch.tensor(bl0))}
        tree = UnRootedTreeModel.from_json(
            {"id": "tree", "type": "UnRootedTreeModel", "newick": newick,
             "branch_lengths": "blens", "taxa": "taxa"}, dic)
        leaves = {"blens": dic["blens"], "shape": Parameter("shape", torch.tensor([0.6])),
                  "kappa": Parameter("kappa", torch.tensor([2.7])),
                  "alpha": Parameter("alpha", torch.tensor([1.3])),
                  "beta": Parameter("beta", torch.tensor([0.4])),
                  "freqs": Parameter("freqs", torch.tensor(f0))}
        subst = MG94("mg94", datatype, leaves["alpha"], leaves["beta"], leaves["kappa"],
                     leaves["freqs"])
        like = cls("like", sp, tree, subst, WeibullSiteModel("sm", leaves["shape"], 4))
        for p in leaves.values():
            p.requires_grad = True
        value = like()
        value.sum().backward()
        results.append((value.detach().clone(), {k: p.grad.clone() for k, p in leaves.items()},
                        int(like.weights.sum())))
    (v_ref, g_ref, n_ref), (v_new, g_new, n_new) = results
    assert n_ref == n_new == 329      # 987 nucleotides = 329 codon sites
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0), (v_new, v_ref)
    for n in g_ref:
        tol = 1e-8 if n in ("blens", "shape") else 1e-7   # eigh backward of the reference (F12)
        assert torch.allclose(g_new[n], g_ref[n], rtol=tol, atol=tol * g_ref[n].abs().max()), \
            (n, (g_new[n] - g_ref[n]).abs().max().item(), g_ref[n].abs().max().item())


@pytest.mark.parametrize("coalescent", ["constant", "skyride", "skygrid"])
def test_cli_time_tree_advi_with_device_heights_and_coalescent(patched, tmp_path, capsys, monkeypatch,
                                                               coalescent):
    """BASELINE config 3's model on real data, end to end: `torchtree-cli advi -m JC69 --clock strict
    --coalescent constant --heights ratio` writes the JSON; the stock runner optimises it once as
    generated and once with the likelihood (`--b200`), the constant coalescent (`--b200_coalescent`)
    and the ratio -> node-height transform (`--b200-heights`, console-script flag) on the device:
    same seed, same draws -> the same ELBO trace."""
    import re

    from torchtree.core.utils import REGISTERED_CLASSES

    import torchtree.evolution.tree_height_transform as ref_transform
    import torchtree.evolution.tree_likelihood as refmod
    import torchtree.evolution.tree_model as ref_tree_model

    sys.path.insert(0, REPO)
    base = ["advi", "-i", DATA + "/fluA.fa", "-t", DATA + "/fluA.tree", "-m", "JC69",
            "--clock", "strict", "--coalescent", coalescent, "--heights", "ratio",
            "--iter", "4", "--elbo_samples", "3", "--grad_samples", "2", "--convergence_every", "2",
            "--stem", str(tmp_path / "run")]
    if coalescent == "skygrid":
        base += ["--grid", "6", "--cutoff", "30"]
    model_name = {"constant": "ConstantCoalescentModel",
                  "skyride": "PiecewiseConstantCoalescentModel",
                  "skygrid": "PiecewiseConstantCoalescentGridModel"}[coalescent]
    traces = {}
    saved_reg = dict(REGISTERED_CLASSES)
    saved = (refmod.TreeLikelihoodModel, ref_transform.GeneralNodeHeightTransform,
             ref_tree_model.GeneralNodeHeightTransform)
    try:
        for tag, extra in (("reference", []), ("b200", ["--b200", "--b200_coalescent"])):
            cfg = _run_cli(base + extra, capsys)
            assert ("torchtree_b200.TreeLikelihoodModel" in cfg) == bool(extra)
            assert (("torchtree_b200.coalescent." + model_name) in cfg) == bool(extra)
            path = tmp_path / (tag + ".json")
            path.write_text(cfg)
            if extra and patched.BACKEND == "cuda":
                # the device height transform as well (what `torchtree-b200 --b200-heights` installs)
                patched.install(override_reference=False, height_transform=True)
            if extra and patched.BACKEND == "oracle":
                import torchtree_b200.coalescent as cmod
                from oracle.coalescent import (constant_log_prob, piecewise_grid_log_prob,
                                               piecewise_log_prob)

                monkeypatch.setattr(cmod, "constant_coalescent_log_prob",
                                    lambda h, th, device=0: constant_log_prob(h, th))
                monkeypatch.setattr(
                    cmod, "piecewise_coalescent_log_prob",
                    lambda h, th, grid=None, device=0: piecewise_log_prob(h, th) if grid is None
                    else piecewise_grid_log_prob(h, th, grid))
            out = _run_torchtree(str(path), capsys)
            elbos = [float(m.group(1)) for m in
                     re.finditer(r"^\s*\d+\s+(-?\d+\.\d+)\s+\d+\.\d+\s+\d+\.\d+", out, flags=re.M)]
            assert len(elbos) >= 2, out[-2000:]
            traces[tag] = elbos
    finally:
        REGISTERED_CLASSES.update(saved_reg)
        refmod.TreeLikelihoodModel = saved[0]
        ref_transform.GeneralNodeHeightTransform = saved[1]
        ref_tree_model.GeneralNodeHeightTransform = saved[2]
    # (the runner prints three decimals: allow one unit in the last printed place)
    np.testing.assert_allclose(np.array(traces["b200"]), np.array(traces["reference"]), rtol=1e-7,
                               atol=1.5e-3)


def test_discrete_trait_likelihood_general_nonsymmetric(patched):
    """The phylogeography likelihood the CLI builds for a discrete trait (cli/evolution.py:540-611):
    `AttributePattern` (one pattern, weight 1.0) + `GeneralNonSymmetricSubstitutionModel` over a
    `GeneralDataType` + `ConstantSiteModel` on the fluA tree.  The reference computes P with
    torch.matrix_exp (abstract.py:89-94); the drop-in runs the matrix exponential and its adjoint on
    the device (csrc/expm.cu, `expm` route of flatten.substitution_route)."""
    from torchtree_b200.flatten import substitution_route

    objs, like = _flu_json()
    places = ["HK", "NY", "TX", "SF", "UK"]
    rng = np.random.default_rng(23)
    for taxon in objs[0]["taxa"]:
        taxon["attributes"]["location"] = places[int(rng.integers(0, 5))]
    objs[0]["taxa"][3]["attributes"]["location"] = "?"        # unknown location: all-ones partial
    n = len(places)
    trait = {
        "id": "like", "type": "TreeLikelihoodModel",
        "tree_model": like["tree_model"],
        "site_model": {"id": "sm", "type": "ConstantSiteModel"},
        "substitution_model": {
            "id": "subst", "type": "GeneralNonSymmetricSubstitutionModel",
            "data_type": {"id": "loc", "type": "GeneralDataType", "codes": places},
            "mapping": list(range(n * (n - 1))),
            "rates": _P("rates", rng.gamma(2.0, 0.5, n * (n - 1)).tolist()),
            "frequencies": _P("freqs", rng.dirichlet(np.full(n, 8.0)).tolist())},
        "site_pattern": {"id": "sp", "type": "torchtree.evolution.attribute_pattern.AttributePattern",
                         "taxa": "taxa",
                         "data_type": "loc", "attribute": "location"},
    }
    ref = _build(objs[:1], trait, "torchtree.evolution.tree_likelihood.TreeLikelihoodModel")
    new = _build(objs[:1], trait, "torchtree_b200.TreeLikelihoodModel")
    assert substitution_route(new["like"].subst_model) == "expm"
    assert new["like"]._state_count == n and int(new["like"].weights.numel()) == 1
    names = ["blens", "rates", "freqs"]
    v_ref, g_ref = _grads(ref, names)
    v_new, g_new = _grads(new, names)
    assert v_new.shape == v_ref.shape
    assert torch.allclose(v_new, v_ref, rtol=1e-10, atol=0), (v_new, v_ref)
    for name in names:   # matrix_exp on both sides: no eigen-gap slack, 1e-8 throughout
        assert torch.allclose(g_new[name], g_ref[name], rtol=1e-8,
                              atol=1e-8 * g_ref[name].abs().max()), name


@pytest.mark.parametrize("kind", ["skyride", "skygrid"])
def test_device_piecewise_coalescent_classes_are_dropins(patched, monkeypatch, kind):
    """`PiecewiseConstantCoalescentModel` (skyride) and `PiecewiseConstantCoalescentGridModel`
    (skygrid) with the device log-density: same JSON, same value and gradients as the reference
    classes; `install(coalescent=True)` rebinds their bare and dotted names."""
    from torchtree.core.utils import REGISTERED_CLASSES, get_class, process_objects

    import torchtree.evolution.coalescent as refmod

    import torchtree_b200.coalescent as cmod
    from oracle.coalescent import piecewise_grid_log_prob, piecewise_log_prob

    rng = np.random.default_rng(31)
    T = 9
    tips = rng.uniform(0, 3, T)
    inner = tips.max() + np.cumsum(rng.exponential(0.5, T - 1))
    heights = np.concatenate([tips, inner])
    name = "PiecewiseConstantCoalescentModel" if kind == "skyride" \
        else "PiecewiseConstantCoalescentGridModel"
    M = T - 1 if kind == "skyride" else 6

    def build(type_name, batch):
        dic = {}
        th = rng.uniform(1.0, 6.0, M if batch is None else (batch, M))
        data = {"id": "coal", "type": type_name, "theta": _P("theta", th.tolist()),
                "times": heights.tolist(), "events": [1] * T + [0] * (T - 1)}
        if kind == "skygrid":
            data["cutoff"] = float(0.8 * heights.max())
        process_objects(json.loads(json.dumps(data)), dic)
        return dic

    saved_reg, saved_cls = dict(REGISTERED_CLASSES), getattr(refmod, name)
    backend, install = patched.BACKEND, patched.install
    if backend == "oracle":
        monkeypatch.setattr(
            cmod, "piecewise_coalescent_log_prob",
            lambda h, th, grid=None, device=0: piecewise_log_prob(h, th) if grid is None
            else piecewise_grid_log_prob(h, th, grid))
    try:
        for batch in (None, 3):
            state = rng.bit_generator.state
            ref = build("torchtree.evolution.coalescent." + name, batch)
            rng.bit_generator.state = state          # the same thetas for both classes
            new = build("torchtree_b200.coalescent." + name, batch)
            assert type(new["coal"]).__module__ == "torchtree_b200.coalescent"
            assert isinstance(new["coal"], saved_cls)
            for dic in (ref, new):
                dic["theta"].requires_grad = True
                dic["coal"]().sum().backward()
            assert new["coal"]().shape == ref["coal"]().shape
            assert torch.allclose(new["coal"](), ref["coal"](), rtol=1e-12, atol=0)
            assert torch.allclose(new["theta"].grad, ref["theta"].grad, rtol=1e-9,
                                  atol=1e-9 * ref["theta"].grad.abs().max())
        install(override_reference=False, coalescent=True)
        assert get_class(name) is getattr(cmod, name)
        assert get_class("torchtree.evolution.coalescent." + name) is getattr(cmod, name)
    finally:
        REGISTERED_CLASSES.update(saved_reg)
        setattr(refmod, name, saved_cls)
        for other in ("ConstantCoalescentModel", "PiecewiseConstantCoalescentModel",
                      "PiecewiseConstantCoalescentGridModel"):
            if hasattr(refmod, "Reference" + other):
                setattr(refmod, other, getattr(refmod, "Reference" + other))
