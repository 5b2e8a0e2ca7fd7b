"""CPU restatement of the reference's ratio -> node-height transform.

TEST INFRASTRUCTURE ONLY (see oracle/treelik.py): imported by tests/ and nothing else.

Follows GeneralNodeHeightTransform (torchtree/evolution/tree_height_transform.py):
`update_bounds` :36-56, `_call` :58-66, `_inverse` :68-93, `log_abs_det_jacobian` :95-98.
Pinned by tests/golden/heights_*.npz, generated from the real reference class by
tests/golden/make_golden_heights.py.
"""
import torch


def parents_from_postorder(tip_count, postorder):
    """parent[node] for every node (root: -1)."""
    parent = [-1] * (2 * tip_count - 1)
    for node, left, right in postorder:
        parent[int(left)] = int(node)
        parent[int(right)] = int(node)
    return parent


def internal_bounds(tip_count, postorder, sampling_times):
    """Lower bound of each internal node = latest sampling time among its tips (:36-56)."""
    b = [float(t) for t in sampling_times] + [0.0] * (tip_count - 1)
    for node, left, right in postorder:
        b[int(node)] = max(b[int(left)], b[int(right)])
    return torch.tensor(b[tip_count:], dtype=torch.float64)


def heights_from_ratios(tip_count, postorder, bounds, x):
    """x [..., T-1] (ratios, root height at the root's index) -> heights [..., T-1] (:58-66).
    Written with autograd-friendly (out-of-place) ops so that `.backward()` gives the
    reference gradient."""
    T = tip_count
    parent = parents_from_postorder(T, postorder)
    cols = [None] * (T - 1)
    root = int(postorder[-1][0])
    cols[root - T] = x[..., root - T]
    for node, left, right in reversed([tuple(int(v) for v in t) for t in postorder]):
        for c in (left, right):
            if c >= T:
                cols[c - T] = bounds[c - T] + x[..., c - T] * (cols[parent[c] - T] - bounds[c - T])
    return torch.stack(cols, -1)


def log_abs_det_jacobian(tip_count, postorder, bounds, heights):
    """sum over non-root internal nodes of log(h[parent] - b[node]) (:95-98)."""
    T = tip_count
    parent = parents_from_postorder(T, postorder)
    root = int(postorder[-1][0])
    nodes = [n for n in range(T, 2 * T - 1) if n != root]
    pidx = torch.tensor([parent[n] - T for n in nodes])
    nidx = torch.tensor([n - T for n in nodes])
    return torch.log(heights[..., pidx] - bounds[nidx]).sum(-1)


def ratios_from_heights(tip_count, postorder, bounds, heights):
    """Inverse transform (:68-93)."""
    T = tip_count
    parent = parents_from_postorder(T, postorder)
    root = int(postorder[-1][0])
    cols = []
    for n in range(T, 2 * T - 1):
        if n == root:
            cols.append(heights[..., n - T])
        else:
            cols.append((heights[..., n - T] - bounds[n - T]) /
                        (heights[..., parent[n] - T] - bounds[n - T]))
    return torch.stack(cols, -1)
