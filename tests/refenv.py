"""Where the tests find the real reference (4ment/torchtree 1.0.2).

Preference order: the vendored install `baseline/_ref/` (written by tools/vendor_reference.py,
git-ignored, travels to the GPU box), then the read-only checkout `/root/reference` of the
authoring container.  `dendropy` (parsing only) is the stand-in under `oracle/dendropy_shim`.
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_VENDORED = os.path.join(REPO, "baseline", "_ref")
_CHECKOUT = "/root/reference"
SHIM = os.path.join(REPO, "oracle", "dendropy_shim")


def reference_root():
    """(python path entry, data directory) or (None, None)."""
    if os.path.isfile(os.path.join(_VENDORED, "torchtree", "__init__.py")) and \
            os.path.isfile(os.path.join(_VENDORED, "data", "fluA.fa")):
        return _VENDORED, os.path.join(_VENDORED, "data")
    if os.path.isdir(os.path.join(_CHECKOUT, "torchtree")):
        return _CHECKOUT, os.path.join(_CHECKOUT, "data")
    return None, None


def available() -> bool:
    return reference_root()[0] is not None


def data_dir() -> str:
    return reference_root()[1]


def activate():
    """Put the reference and the dendropy stand-in on sys.path; returns the entries added."""
    root, _ = reference_root()
    if root is None:
        raise RuntimeError("the reference is not available (run tools/vendor_reference.py)")
    added = []
    for p in (SHIM, root, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
            added.append(p)
    return added
