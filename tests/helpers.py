"""Shared test helpers: golden fixture loading and comparison utilities."""
import glob
import os

import numpy as np

from torchtree_b200.synthetic import Problem

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LNL_RTOL = 1e-10  # north_star tolerance on log-likelihood (relative)
GRAD_RTOL = 1e-8  # north_star tolerance on gradients (relative)


def golden_names():
    return sorted(
        os.path.splitext(os.path.basename(p))[0]
        for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
        if not p.endswith("kats.npz")
    )


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    prob = Problem(
        tip_count=int(rec["T"]),
        pattern_count=int(rec["N"]),
        state_count=int(rec["S"]),
        category_count=int(rec["K"]),
        postorder=rec["postorder"].astype(np.int32),
        tip_states=rec["tip_states"].astype(np.uint8),
        weights=rec["weights"].astype(np.float64),
        branch_lengths=rec["branch_lengths"],
        site_rates=rec["site_rates"],
        site_props=rec["site_props"],
        freqs=rec["freqs"],
        q_matrix=rec["q_matrix"],
        code_partials=rec["code_partials"],
    )
    return prob, rec


def assert_grad_close(got, want, rtol=GRAD_RTOL, what=""):
    """Relative 1e-8 per component with an absolute floor scaled to the
    largest component (entries that are ~0 by cancellation)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    floor = rtol * max(1.0, float(np.max(np.abs(want)))) if want.size else 0.0
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + floor
    bad = err > tol
    assert not bad.any(), "%s: max rel err %.3e at %s (got %r want %r)" % (
        what,
        float((err / np.maximum(np.abs(want), 1e-300)).max()),
        np.argwhere(bad)[:3].tolist(),
        got[bad][:3],
        want[bad][:3],
    )


def assert_lnl_close(got, want, rtol=LNL_RTOL, what=""):
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    rel = np.abs(got - want) / np.abs(want)
    assert (rel <= rtol).all(), "%s lnL got %r want %r rel %r" % (what, got, want, rel)
