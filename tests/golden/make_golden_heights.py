"""Golden vectors for the node-height transform, from the REAL reference class
(torchtree/evolution/tree_height_transform.py) driven by a duck-typed tree object.

    PYTHONPATH=oracle/dendropy_shim:/root/reference python tests/golden/make_golden_heights.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "dendropy_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from torchtree.evolution.tree_height_transform import GeneralNodeHeightTransform  # noqa: E402

from torchtree_b200.synthetic import random_postorder  # noqa: E402


class Tree:
    """What the reference transform reads from a TimeTreeModel."""

    def __init__(self, T, postorder, sampling_times):
        self.taxa_count = T
        self.postorder = [tuple(int(v) for v in t) for t in postorder]
        self.sampling_times = sampling_times
        parent = {}
        children = {}
        for n, l, r in self.postorder:
            parent[l] = n
            parent[r] = n
            children[n] = (l, r)
        root = self.postorder[-1][0]
        pairs, stack = [], [root]
        while stack:                      # pre-order (parent, node) pairs
            n = stack.pop()
            if n != root:
                pairs.append((parent[n], n))
            if n in children:
                stack.extend(reversed(children[n]))
        self.preorder = torch.tensor(pairs)


def make(name, T, D, seed, topology):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    postorder = random_postorder(T, rng, topology)
    times = torch.tensor(rng.uniform(0.0, 10.0, T) * (rng.random(T) < 0.7), dtype=torch.float64)
    tree = Tree(T, postorder, times)
    transform = GeneralNodeHeightTransform(tree)
    x = torch.rand(D, T - 1, dtype=torch.float64) * 0.98 + 0.01
    root = tree.postorder[-1][0] - T
    x[:, root] = 12.0 + 30.0 * torch.rand(D, dtype=torch.float64)
    x.requires_grad_(True)
    heights = transform(x)
    logdet = transform.log_abs_det_jacobian(x, heights)
    g = torch.randn(D, T - 1, dtype=torch.float64)
    (grad_x,) = torch.autograd.grad((heights * g).sum(), x, retain_graph=True)
    (grad_x_logdet,) = torch.autograd.grad(logdet.sum(), x)
    # the reference's _inverse concatenates on dim 0: it only handles unbatched input
    inv = torch.stack([transform.inv(heights.detach()[d]) for d in range(D)])
    np.savez_compressed(
        os.path.join(HERE, "heights", name + ".npz"), T=T, D=D, postorder=np.asarray(postorder, dtype=np.int32),
        sampling_times=times.numpy(), bounds=transform._bounds[T:].numpy(), x=x.detach().numpy(),
        heights=heights.detach().numpy(), logdet=logdet.detach().numpy(), g=g.numpy(),
        grad_x=grad_x.numpy(), grad_x_logdet=grad_x_logdet.numpy(), inverse=inv.numpy())
    print(name, "T", T, "D", D, "max|inv - x|", float((inv - x.detach()).abs().max()))


if __name__ == "__main__":
    make("heights_random60", 60, 3, 1, "random")
    make("heights_caterpillar40", 40, 2, 2, "caterpillar")
    make("heights_random500", 500, 4, 3, "random")
