"""Drop-in `TreeLikelihoodModel` for torchtree, computed by the ttb200 CUDA engine.

Same constructor arguments, JSON keys and output contract as the reference class
(torchtree/evolution/tree_likelihood.py:281-450):

    {"id": ..., "type": "torchtree_b200.TreeLikelihoodModel",
     "tree_model": ..., "site_model": ..., "substitution_model": ...,
     "site_pattern": ..., ["branch_model": ...],
     ["use_ambiguities": false], ["use_tip_states": false],
     ["device": 0]}

This module needs torchtree importable; the engine underneath does not.
"""
from __future__ import annotations

import torch
from torchtree.core.model import CallableModel
from torchtree.core.utils import process_object, register_class
from torchtree.evolution.branch_model import BranchModel
from torchtree.evolution.site_model import SiteModel
from torchtree.evolution.site_pattern import SitePattern
from torchtree.evolution.substitution_model.abstract import SubstitutionModel
from torchtree.evolution.tree_model import TreeModel

from .engine import Engine, codes_from_tip_partials
from .flatten import evaluate_models


class TreeLikelihoodModel(CallableModel):
    """B200-native replacement of torchtree's TreeLikelihoodModel.

    Always rescales (power-of-two factors per node and pattern), so there is
    no `rescale` latch (tree_likelihood.py:378-388): the value returned equals
    what the reference returns once it has switched to its rescaled path.
    """

    def __init__(self, id_, site_pattern, tree_model, subst_model, site_model,
                 clock_model=None, use_ambiguities=False, use_tip_states=False,
                 device: int = 0):
        super().__init__(id_)
        self.site_pattern = site_pattern
        self.tree_model = tree_model
        self.subst_model = subst_model
        self.site_model = site_model
        self.clock_model = clock_model
        self.use_tip_states = use_tip_states
        self.use_ambiguities = use_ambiguities
        self.device_index = int(device)
        state_count = subst_model.frequencies.shape[-1]
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
        self._engine = None
        self._engine_postorder = None

    # -- engine management ---------------------------------------------------
    def _get_engine(self, draws: int) -> Engine:
        postorder = self.tree_model.postorder
        K = self.site_model.rates().shape[-1]
        if (self._engine is None or self._engine.max_draws < draws or self._engine.K != K):
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(
                self._tip_codes, self.weights.to(torch.float64).numpy(), postorder,
                self._state_count, K, code_partials=self._code_partials,
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
        return evaluate_models(engine, self.tree_model, self.site_model, self.subst_model,
                               self.clock_model, sample_shape)

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
                   use_ambiguities, use_tip_states, device=data.get("device", 0))


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
    * `coalescent=True` makes `ConstantCoalescentModel` (bare and dotted) resolve to the device
      version (coalescent.py).
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
        cls = b200_coalescent.ConstantCoalescentModel
        register_class(cls, "torchtree_b200.ConstantCoalescentModel")
        register_class(cls, "ConstantCoalescentModel")
        if not hasattr(ref_coalescent, "ReferenceConstantCoalescentModel"):
            ref_coalescent.ReferenceConstantCoalescentModel = ref_coalescent.ConstantCoalescentModel
        ref_coalescent.ConstantCoalescentModel = cls
    if height_transform:
        import torchtree.evolution.tree_height_transform as ref_transform
        import torchtree.evolution.tree_model as ref_tree_model

        from .height_transform import GeneralNodeHeightTransform
        for module in (ref_transform, ref_tree_model):
            if not hasattr(module, "ReferenceGeneralNodeHeightTransform"):
                module.ReferenceGeneralNodeHeightTransform = module.GeneralNodeHeightTransform
            module.GeneralNodeHeightTransform = GeneralNodeHeightTransform
