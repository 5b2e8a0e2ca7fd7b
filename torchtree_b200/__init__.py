"""torchtree_b200 -- B200-native tree-likelihood engine, shipped as a torchtree plug-in.

`__plugin__` is what torchtree's plug-in manager looks for
(torchtree/cli/plugin_manager.py:10-23).  The engine itself (`Engine`,
`log_likelihood_eigen`, `log_likelihood_mats`) does not need torchtree; the
drop-in `TreeLikelihoodModel` (torchtree_b200.tree_likelihood) does.
"""
__version__ = "0.1.0"
__plugin__ = "cli.B200Plugin"

from .engine import Engine, codes_from_tip_partials, default_code_partials  # noqa: F401
from .function import (  # noqa: F401
    log_likelihood_eigen,
    log_likelihood_expm,
    log_likelihood_mats,
    reversible_eigensystem,
)
from ._lib import EngineError  # noqa: F401
from .coalescent import constant_coalescent_log_prob, piecewise_coalescent_log_prob  # noqa: F401


def __getattr__(name):
    # resolved lazily so that the engine can be used where torchtree is absent;
    # `"type": "torchtree_b200.TreeLikelihoodModel"` reaches this through
    # torchtree.core.utils.get_class (core/utils.py:116-125)
    if name == "TreeLikelihoodModel":
        from .tree_likelihood import TreeLikelihoodModel

        return TreeLikelihoodModel
    if name == "ConstantCoalescentModel":
        from .coalescent import ConstantCoalescentModel

        return ConstantCoalescentModel
    if name == "install":
        from .tree_likelihood import install

        return install
    raise AttributeError(name)
