"""Drop-in `TreeLikelihoodModel` for torchtree, computed by the ttb200 CUDA engine.

Same constructor arguments, JSON keys and output contract as the reference class
(torchtree/evolution/tree_likelihood.py:281-450):

    {"id": ..., "type": "torchtree_b200.TreeLikelihoodModel",
     "tree_model": ..., "site_model": ..., "substitution_model": ...,
     "site_pattern": ..., ["branch_model": ...],
     ["use_ambiguities": false], ["use_tip_states": false],
     ["device": 0], ["devices": [0, 1, ...]], ["shard": "patterns" | "draws"]}

Multi-GPU: one process per GPU (`torchrun --nproc-per-node G torchtree-b200 cfg.json`,
or any launcher that sets RANK / LOCAL_RANK / WORLD_SIZE).  With `"shard": "patterns"` every
rank builds its engine on its slice of the site patterns; with `"shard": "draws"` every rank
holds the whole alignment and evaluates its slice of the batch of draws.  `"devices"` maps the
local rank to a CUDA ordinal (default: the local rank itself).  The exchange per evaluation is
one all-reduce of lnL and one of the packed gradient (torchtree_b200/sharded.py).

dtype policy: the engine computes in fp64 whatever `--dtype` says; float32 parameters are
converted on the way in and lnL / gradients come back as float32 (csrc/torch_ext.cpp).

This module needs torchtree importable; the engine underneath does not.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
from torchtree.core.model import CallableModel
from torchtree.core.utils import process_object, register_class
from torchtree.evolution.branch_model import BranchModel
from torchtree.evolution.site_model import SiteModel
from torchtree.evolution.site_pattern import SitePattern
from torchtree.evolution.substitution_model.abstract import SubstitutionModel
from torchtree.evolution.tree_model import TreeModel

from .engine import Engine, codes_from_tip_partials
from .flatten import evaluate_models


MAX_STATES = 64   # ttb2_create's limit (include/ttb200.h)


def _join_process_group():
    """(rank, world size) of this process; initialises torch.distributed from the launcher's
    environment (RANK / WORLD_SIZE / MASTER_*) when a launcher is present and nobody did yet."""
    if not dist.is_available():
        return 0, 1
    if not dist.is_initialized():
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return 0, 1
        backend = os.environ.get("TTB200_DIST_BACKEND", "nccl")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


class TreeLikelihoodModel(CallableModel):
    """B200-native replacement of torchtree's TreeLikelihoodModel.

    Always rescales (power-of-two factors per node and pattern), so there is
    no `rescale` latch (tree_likelihood.py:378-388): the value returned equals
    what the reference returns once it has switched to its rescaled path.
    """

    def __init__(self, id_, site_pattern, tree_model, subst_model, site_model,
                 clock_model=None, use_ambiguities=False, use_tip_states=False,
                 device: int = 0, devices=None, shard=None):
        super().__init__(id_)
        state_count = subst_model.frequencies.shape[-1]
        if state_count > MAX_STATES:
            # no CPU fallback and no silent substitute (SURVEY 8(b) policy): say so at construction
            raise NotImplementedError(
                "torchtree_b200.TreeLikelihoodModel supports at most %d states (got %d); keep the "
                "reference TreeLikelihoodModel for this likelihood" % (MAX_STATES, state_count))
        if shard not in (None, "patterns", "draws"):
            raise ValueError('"shard" must be "patterns" or "draws", got %r' % (shard,))
        self.shard = shard
        self._group = None
        self._rank, self._world = 0, 1
        if shard is not None:
            self._rank, self._world = _join_process_group()
        local_rank = int(os.environ.get("LOCAL_RANK", self._rank))
        if devices:
            device = list(devices)[local_rank % len(devices)]
        elif shard is not None and self._world > 1:
            device = local_rank
        self.site_pattern = site_pattern
        self.tree_model = tree_model
        self.subst_model = subst_model
        self.site_model = site_model
        self.clock_model = clock_model
        self.use_tip_states = use_tip_states
        self.use_ambiguities = use_ambiguities
        self.device_index = int(device)
        if type(site_pattern) is SitePattern:
            # stock SitePattern: compress natively (patterns.py, ttb2_compress_patterns)
            # instead of the per-character Python loops of site_pattern.py:69-151
            from .patterns import tip_codes_from_alignment
            codes, table, weights = tip_codes_from_alignment(
                site_pattern.alignment, use_ambiguities and not use_tip_states,
                site_pattern.indices)
            self._tip_codes = codes
            self._code_partials = None if use_tip_states else table
            self.weights = torch.from_numpy(weights).to(torch.int64)
        elif use_tip_states:
            # integer codes, gaps = S (site_pattern.py:127-151)
            states, self.weights = site_pattern.compute_tips_states()
            self._tip_codes = torch.stack(list(states)).clamp(max=state_count).to(torch.uint8).numpy()
            self._code_partials = None
        else:
            partials, self.weights = site_pattern.compute_tips_partials(use_ambiguities)
            self._tip_codes, self._code_partials = codes_from_tip_partials(
                [p.numpy() for p in partials], state_count)
        self._state_count = int(state_count)
        self.pattern_range = (0, int(self.weights.shape[0]))
        if shard == "patterns" and self._world > 1:
            from .sharded import shard_range

            lo, hi = shard_range(int(self.weights.shape[0]), self._rank, self._world)
            self.pattern_range = (lo, hi)
        self._engine = None
        self._engine_postorder = None

    # -- engine management ---------------------------------------------------
    def _get_engine(self, draws: int):
        postorder = self.tree_model.postorder
        K = self.site_model.rates().shape[-1]
        lo, hi = self.pattern_range
        if hi <= lo:
            return None   # more ranks than patterns: this rank only joins the collectives
        if self.shard == "draws" and self._world > 1:
            draws = (draws + self._world - 1) // self._world
        if (self._engine is None or self._engine.max_draws < draws or self._engine.K != K):
            if self._engine is not None:
                # not close(): a pending backward of an earlier evaluation may still need it
                self._engine.release()
            self._engine = Engine(
                self._tip_codes[:, lo:hi], self.weights.to(torch.float64).numpy()[lo:hi],
                postorder, self._state_count, K, code_partials=self._code_partials,
                max_draws=draws, device=self.device_index)
            self._engine_postorder = list(postorder)
        elif self._engine_postorder != list(postorder):
            self._engine.set_postorder(postorder)
            self._engine_postorder = list(postorder)
        return self._engine

    # -- CallableModel ---------------------------------------------------------
    def _call(self, *args, **kwargs) -> torch.Tensor:
        sample_shape = self.sample_shape
        draws = 1
        for n in sample_shape:
            draws *= int(n)
        engine = self._get_engine(draws)
        shard = (self.shard, self._group) if self.shard is not None and self._world > 1 else None
        return evaluate_models(engine, self.tree_model, self.site_model, self.subst_model,
                               self.clock_model, sample_shape, shard=shard)

    def handle_parameter_changed(self, variable, index, event):
        pass

    def _sample_shape(self) -> torch.Size:
        return max([model.sample_shape for model in self._models.values()], key=len)

    @classmethod
    def from_json(cls, data, dic):
        id_ = data["id"]
        tree_model = process_object(data[TreeModel.tag], dic)
        site_model = process_object(data[SiteModel.tag], dic)
        subst_model = process_object(data[SubstitutionModel.tag], dic)
        site_pattern = process_object(data[SitePattern.tag], dic)
        use_ambiguities = data.get("use_ambiguities", False)
        use_tip_states = data.get("use_tip_states", False)
        clock_model = None
        if BranchModel.tag in data:
            clock_model = process_object(data[BranchModel.tag], dic)
        return cls(id_, site_pattern, tree_model, subst_model, site_model, clock_model,
                   use_ambiguities, use_tip_states, device=data.get("device", 0),
                   devices=data.get("devices"), shard=data.get("shard"))


def install(override_reference: bool = True, height_transform: bool = False,
            coalescent: bool = False) -> None:
    """Make existing configs resolve to this class (SURVEY 8(b) "Resolution"):

    * bare `"type": "TreeLikelihoodModel"` -> registry entry replaced;
    * dotted `torchtree.evolution.tree_likelihood.TreeLikelihoodModel` -> the
      module attribute is rebound;
    * bare `"LG"` / `"WAG"` are registered (the reference forgets to);
    * `height_transform=True` additionally rebinds `GeneralNodeHeightTransform`
      where `ReparameterizedTimeTreeModel` looks it up (tree_model.py:539, :587), so
      time trees map ratios to node heights on the GPU (height_transform.py);
    * `coalescent=True` makes `ConstantCoalescentModel`, `PiecewiseConstantCoalescentModel` (skyride)
      and `PiecewiseConstantCoalescentGridModel` (skygrid), bare and dotted, resolve to the device
      versions (coalescent.py).
    """
    import torchtree.evolution.tree_likelihood as ref_module
    from torchtree.evolution.substitution_model.amino_acid import LG, WAG

    register_class(TreeLikelihoodModel, "torchtree_b200.TreeLikelihoodModel")
    if override_reference:
        register_class(TreeLikelihoodModel, "TreeLikelihoodModel")
        if not hasattr(ref_module, "ReferenceTreeLikelihoodModel"):
            ref_module.ReferenceTreeLikelihoodModel = ref_module.TreeLikelihoodModel
        ref_module.TreeLikelihoodModel = TreeLikelihoodModel
    register_class(LG, "LG")
    register_class(WAG, "WAG")
    if coalescent:
        import torchtree.evolution.coalescent as ref_coalescent

        from . import coalescent as b200_coalescent
        for name in ("ConstantCoalescentModel", "PiecewiseConstantCoalescentModel",
                     "PiecewiseConstantCoalescentGridModel"):
            cls = getattr(b200_coalescent, name)
            register_class(cls, "torchtree_b200." + name)
            register_class(cls, name)
            if not hasattr(ref_coalescent, "Reference" + name):
                setattr(ref_coalescent, "Reference" + name, getattr(ref_coalescent, name))
            setattr(ref_coalescent, name, cls)
    if height_transform:
        import torchtree.evolution.tree_height_transform as ref_transform
        import torchtree.evolution.tree_model as ref_tree_model

        from .height_transform import GeneralNodeHeightTransform
        for module in (ref_transform, ref_tree_model):
            if not hasattr(module, "ReferenceGeneralNodeHeightTransform"):
                module.ReferenceGeneralNodeHeightTransform = module.GeneralNodeHeightTransform
            module.GeneralNodeHeightTransform = GeneralNodeHeightTransform
