"""Time native site-pattern compression against the reference's `compress`
(site_pattern.py:69-97).  Runs here (needs /root/reference); host-only code.

    python tools/bench_patterns.py [taxa] [sites]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "dendropy_shim"))
sys.path.insert(0, "/root/reference")

from torchtree.evolution.alignment import Alignment, Sequence   # noqa: E402
from torchtree.evolution.datatype import NucleotideDataType     # noqa: E402
from torchtree.evolution.site_pattern import compress_alignment  # noqa: E402
from torchtree.evolution.taxa import Taxa, Taxon                 # noqa: E402

from torchtree_b200.patterns import tip_codes_from_alignment     # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
    rng = np.random.default_rng(0)
    # an evolving-ish alignment: a root sequence with per-taxon mutations
    root = rng.integers(0, 4, L)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for _ in range(T):
        s = root.copy()
        mut = rng.random(L) < 0.1
        s[mut] = rng.integers(0, 4, int(mut.sum()))
        seqs.append(letters[s].tobytes().decode())
    names = [f"t{i}" for i in range(T)]
    taxa = Taxa("taxa", [Taxon(n, {}) for n in names])
    aln = Alignment("a", [Sequence(n, s) for n, s in zip(names, seqs)], taxa, NucleotideDataType("nuc"))
    t0 = time.perf_counter()
    codes, table, w = tip_codes_from_alignment(aln, False)
    t_native = time.perf_counter() - t0
    out = {"taxa": T, "sites": L, "patterns": int(codes.shape[1]), "native_s": round(t_native, 4)}
    if os.environ.get("SKIP_REFERENCE") != "1":
        t0 = time.perf_counter()
        partials, w_ref = compress_alignment(aln, None, False)
        out["reference_s"] = round(time.perf_counter() - t0, 3)
        out["speedup"] = round(out["reference_s"] / t_native, 1)
        assert w_ref.tolist() == w.tolist()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
