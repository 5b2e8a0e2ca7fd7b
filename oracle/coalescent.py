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


def _sorted_terms(node_heights, grid=None):
    """Shared front end of the piecewise models (coalescent.py:312-341, :470-502): events sorted by
    height, their masks (+1 tip, -1 coalescence, 0 grid point), C(k, 2) and the interval lengths."""
    n = node_heights.shape[-1]
    taxa = (n + 1) // 2
    masks = [torch.ones(taxa, dtype=torch.long), -torch.ones(taxa - 1, dtype=torch.long)]
    heights = node_heights
    if grid is not None:
        heights = torch.cat([node_heights, grid.expand(node_heights.shape[:-1] + (-1,))], -1)
        masks.append(torch.zeros(grid.shape[-1], dtype=torch.long))
    mask = torch.cat(masks).expand(heights.shape)
    order = torch.argsort(heights, descending=False, stable=True)
    sorted_heights = torch.gather(heights, -1, order)
    sorted_mask = torch.gather(mask, -1, order)
    lineages = sorted_mask.cumsum(-1)[..., :-1]
    lchoose2 = (lineages * (lineages - 1)).to(node_heights.dtype) / 2.0
    return sorted_mask, lchoose2, sorted_heights[..., 1:] - sorted_heights[..., :-1]


def piecewise_log_prob(node_heights: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """Skyride: PiecewiseConstantCoalescent.log_prob (coalescent.py:379-396).
    node_heights [..., 2T-1], theta [..., T-1] -> [..., 1]."""
    batch = torch.broadcast_shapes(node_heights.shape[:-1], theta.shape[:-1])
    h = node_heights.expand(batch + node_heights.shape[-1:])
    mask, lchoose2, durations = _sorted_terms(h)
    index = (mask == -1).long().cumsum(-1)[..., :-1]                               # :382-388
    thetas = theta.expand(batch + theta.shape[-1:]).gather(-1, index)              # :390
    return -torch.sum(lchoose2 * durations / thetas, -1, keepdim=True) - \
        theta.log().sum(-1, keepdim=True).expand(batch + (1,))                     # :391-395


def piecewise_grid_log_prob(node_heights: torch.Tensor, theta: torch.Tensor,
                            grid: torch.Tensor) -> torch.Tensor:
    """Skygrid: PiecewiseConstantCoalescentGrid.log_prob (coalescent.py:523-549).
    node_heights [..., 2T-1], theta [..., G+1], grid [G] -> [..., 1]."""
    batch = torch.broadcast_shapes(node_heights.shape[:-1], theta.shape[:-1])
    h = node_heights.expand(batch + node_heights.shape[-1:])
    mask, lchoose2, durations = _sorted_terms(h, grid)
    index = (mask == 0).long().cumsum(-1)                                          # :527-531
    thetas = theta.expand(batch + theta.shape[-1:]).gather(-1, index)              # :533-538
    log_thetas = torch.where(mask == -1, torch.log(thetas), torch.zeros_like(thetas))  # :540-544
    return torch.sum(-lchoose2 * durations / thetas[..., :-1] - log_thetas[..., 1:], -1,
                     keepdim=True)                                                 # :545-549
