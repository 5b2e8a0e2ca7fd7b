#!/usr/bin/env python
"""Install the UNMODIFIED reference (4ment/torchtree) into git-ignored `baseline/_ref/`.

    python tools/vendor_reference.py [--force]

`baseline/_ref/` is ignored by git (no reference source enters the history) but NOT by
gpurun, so it travels to the GPU box with the snapshot: the `-m gpu` product tests build
the same JSON once with the reference `TreeLikelihoodModel` and once with
`torchtree_b200.TreeLikelihoodModel` on the real CUDA engine, and `bench.py --impl reference`
times the reference's own CPU implementation at the stated configuration.

What it does (the recipe the build contract prescribes):
  pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>
from a scratch copy under /tmp (the checkout is read-only and setuptools writes an egg-info
directory), then copies the two fluA data files the reference's own tests and BASELINE config 1
use into `baseline/_ref/data/`.  `dendropy` (a parsing-only dependency, not installable here)
is supplied at run time by the stand-in in `oracle/dendropy_shim`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(REPO, "baseline", "_ref")
SOURCE = "/root/reference"
DATA_FILES = ("fluA.fa", "fluA.tree", "tiny.fa", "tiny.nwk")


def installed() -> bool:
    return (os.path.isfile(os.path.join(TARGET, "torchtree", "evolution", "tree_likelihood.py"))
            and all(os.path.isfile(os.path.join(TARGET, "data", f)) for f in DATA_FILES))


def vendor(force: bool = False) -> str:
    """Returns "present", "installed" or "unavailable: <why>"."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(os.path.join(SOURCE, "torchtree")):
        return "unavailable: %s does not exist on this machine" % SOURCE
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="ttref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
               "--no-deps", "--upgrade", "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return "unavailable: pip install failed: %s" % (r.stderr.strip().splitlines() or ["?"])[-1]
    os.makedirs(os.path.join(TARGET, "data"), exist_ok=True)
    for name in DATA_FILES:
        shutil.copyfile(os.path.join(SOURCE, "data", name), os.path.join(TARGET, "data", name))
    return "installed" if installed() else "unavailable: install finished but files are missing"


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv))
