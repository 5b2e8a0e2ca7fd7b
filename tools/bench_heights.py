"""Time the ratio -> node-height transform (forward + backward).

    python tools/bench_heights.py native [T] [D]      # on a GPU box: ttb2_heights_* through host tensors
    python tools/bench_heights.py reference [T] [D]   # here: the reference's Python loop + autograd (CPU)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torchtree_b200.synthetic import random_postorder  # noqa: E402


def problem(T, D):
    rng = np.random.default_rng(7)
    post = random_postorder(T, rng, "random")
    times = torch.tensor(rng.uniform(0, 10, T) * (rng.random(T) < 0.7))
    x = torch.rand(D, T - 1, dtype=torch.float64) * 0.98 + 0.01
    x[:, int(post[-1][0]) - T] = 15.0
    g = torch.randn(D, T - 1, dtype=torch.float64)
    return post, times, x, g


def best(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "native"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    post, times, x, g = problem(T, D)

    class Tree:
        taxa_count = T
        postorder = [tuple(int(v) for v in t) for t in post]
        sampling_times = times

    if mode == "native":
        from torchtree_b200.height_transform import GeneralNodeHeightTransform
        tr = GeneralNodeHeightTransform(Tree())
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "dendropy_shim"))
        sys.path.insert(0, "/root/reference")
        from torchtree.evolution.tree_height_transform import GeneralNodeHeightTransform
        parent = {}
        for n, l, r in Tree.postorder:
            parent[l] = n
            parent[r] = n
        kids = {n: (l, r) for n, l, r in Tree.postorder}
        root = Tree.postorder[-1][0]
        pairs, stack = [], [root]
        while stack:
            n = stack.pop()
            if n != root:
                pairs.append((parent[n], n))
            if n in kids:
                stack.extend(reversed(kids[n]))
        Tree.preorder = torch.tensor(pairs)
        tr = GeneralNodeHeightTransform(Tree())

    def step():
        xx = x.clone().requires_grad_(True)
        h = tr(xx)
        (h * g).sum().backward()
        return xx.grad

    t = best(step, 5 if mode == "reference" else 50)
    print(json.dumps({"mode": mode, "taxa": T, "draws": D, "fwd_bwd_ms": round(t * 1e3, 4),
                      "threads": torch.get_num_threads()}))


if __name__ == "__main__":
    main()
