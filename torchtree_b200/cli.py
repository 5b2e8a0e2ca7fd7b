"""torchtree-cli plug-in hooks (torchtree/cli/plugin.py:4-16,
plugin_manager.py:10-23): adds `--b200` to the advi/hmc/map/mcmc sub-commands and
rewrites the generated tree-likelihood block to the B200 class.

Also provides `main()`: `python -m torchtree_b200.cli cfg.json` runs the stock
torchtree runner with the drop-in installed, so that *existing* JSON files
(bare or dotted reference type names) run unchanged on the engine.
"""
from __future__ import annotations

import sys

from torchtree.cli.plugin import Plugin

MAX_STATES = 64  # ttb2_create's limit


def engine_supports(data: dict):
    """(ok, why) for one generated tree-likelihood block (SURVEY 8(b) policy: rewrite the type only
    for what the engine implements; anything else stays on the reference class -- there is no CPU
    fallback to hide behind).  The engine serves every reversible model through its eigen route and
    every other SubstitutionModel through the model's own p_t (matrices route), any site model and
    any clock; what it cannot hold is a state space beyond 64 states."""
    subst = data.get("substitution_model")
    states = None
    if isinstance(subst, dict):
        states = subst.get("state_count")
        data_type = subst.get("data_type")
        if states is None and isinstance(data_type, dict) and "codes" in data_type:
            states = len(data_type["codes"])
    if states is not None and int(states) > MAX_STATES:
        return False, "%d states > %d" % (int(states), MAX_STATES)
    return True, ""


class B200Plugin(Plugin):
    def load_arguments(self, subparsers):
        for name, parser in subparsers._name_parser_map.items():
            if name in ("advi", "hmc", "map", "mcmc"):
                parser.add_argument(
                    "--b200", action="store_true",
                    help="compute the tree likelihood with the torchtree_b200 CUDA engine")
                parser.add_argument(
                    "--b200_coalescent", action="store_true",
                    help="compute the constant-population coalescent on the GPU as well")
                parser.add_argument(
                    "--b200_device", type=int, default=0,
                    help="CUDA device ordinal used by the torchtree_b200 engine")
                parser.add_argument(
                    "--b200_shard", choices=["patterns", "draws"], default=None,
                    help="multi-GPU runs (one process per GPU, e.g. torchrun): shard the site "
                         "patterns or the batch of draws across the ranks")

    def process_coalescent(self, arg, data):
        # constant, skyride and skygrid coalescents on the device (coalescent.py); other
        # demographic models keep the reference class
        if getattr(arg, "b200_coalescent", False) and data.get("type") in (
                "ConstantCoalescentModel", "PiecewiseConstantCoalescentModel",
                "PiecewiseConstantCoalescentGridModel") and "temperature" not in data:
            data["type"] = "torchtree_b200.coalescent." + data["type"]

    def process_tree_likelihood(self, arg, data):
        if not getattr(arg, "b200", False):
            return
        ok, why = engine_supports(data)
        if not ok:
            sys.stderr.write("torchtree_b200: %s keeps the reference TreeLikelihoodModel (%s)\n"
                             % (data.get("id"), why))
            return
        data["type"] = "torchtree_b200.TreeLikelihoodModel"
        if getattr(arg, "b200_device", 0):
            data["device"] = arg.b200_device
        if getattr(arg, "b200_shard", None):
            data["shard"] = arg.b200_shard


def main(argv=None):
    """torchtree runner with the drop-in installed (console entry point)."""
    from torchtree.torchtree import main as torchtree_main

    from .tree_likelihood import install

    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    # --b200-heights: also run the ratio -> node-height transform of time trees on the GPU
    heights = "--b200-heights" in sys.argv
    if heights:
        sys.argv.remove("--b200-heights")
    # --b200-coalescent: ConstantCoalescentModel on the GPU as well
    coalescent = "--b200-coalescent" in sys.argv
    if coalescent:
        sys.argv.remove("--b200-coalescent")
    install(override_reference=True, height_transform=heights, coalescent=coalescent)
    torchtree_main()


if __name__ == "__main__":
    main()
