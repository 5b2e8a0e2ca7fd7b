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
    lchoose2 = lineages * (lineages - 1) / 2.0                               # :130
    return torch.sum(-lchoose2 * durations / theta, -1, keepdim=True) - (taxa - 1) * torch.log(theta)
