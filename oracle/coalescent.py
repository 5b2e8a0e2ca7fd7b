"""CPU restatement of the constant-population coalescent log-density -- TEST INFRASTRUCTURE ONLY
(imported by tests/ alone; the product path never touches it).

Follows ConstantCoalescent.log_prob, torchtree/evolution/coalescent.py:112-134, on the same
substrate (torch CPU ops + autograd).  Pinned by tests/golden/coalescent/*.npz, generated from the
real reference class (tests/golden/make_golden_coalescent.py)."""
import torch


def constant_log_prob(node_heights: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """node_heights [..., 2T-1] (tips first), theta [..., 1] -> [..., 1]."""
    n = node_heights.shape[-1]
    taxa = (n + 1) // 2
    # +1 when a lineage appears (tip), -1 when two coalesce (coalescent.py:113-122)
    mask = torch.cat([torch.ones(taxa, dtype=torch.long), -torch.ones(taxa - 1, dtype=torch.long)])
    mask = mask.expand(node_heights.shape)
    order = torch.argsort(node_heights, descending=False, stable=True)       # :124
    heights = torch.gather(node_heights, -1, order)                          # :125
    lineages = torch.gather(mask, -1, order).cumsum(-1)[..., :-1]            # :126-127
    durations = heights[..., 1:] - heights[..., :-1]                         # :129
    # :130 -- the reference divides the integer product by 2.0, which yields the DEFAULT float
    # dtype (float64 under the runner's fp64 default, torchtree.py:43-66); stated explicitly here
    # so that the oracle does not depend on the process-wide default (float32 would round
    # C(k, 2) beyond 2^24, i.e. from ~5800 lineages on)
    lchoose2 = (lineages * (lineages - 1)).to(node_heights.dtype) / 2.0
    return torch.sum(-lchoose2 * durations / theta, -1, keepdim=True) - (taxa - 1) * torch.log(theta)
