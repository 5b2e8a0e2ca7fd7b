"""Duck-typed stand-ins for torchtree's model classes (test infrastructure).

torchtree is not installed on the GPU box, so the host-side glue
(torchtree_b200.flatten) is exercised there with these minimal objects, which
expose exactly the attributes the glue reads -- with the same names, shapes and
semantics as the reference classes they are named after.
"""
import torch

from oracle import treelik as orc


class UnRootedTreeModel:  # tree_model.py:243-345
    def __init__(self, blens, postorder):
        self._blens = blens
        self.postorder = [tuple(int(x) for x in row) for row in postorder]

    def branch_lengths(self):
        return self._blens

    @property
    def sample_shape(self):
        return self._blens.shape[:-1]


class TimeTreeModel(UnRootedTreeModel):  # branch_lengths() already 2T-2 long
    pass


class StrictClockModel:  # branch_model.py:36-53
    def __init__(self, rate, branch_count):
        self._rate = rate
        self.branch_count = branch_count

    @property
    def rates(self):
        return self._rate.expand([-1] * (self._rate.dim() - 1) + [self.branch_count])

    @property
    def sample_shape(self):
        return self._rate.shape[:-1]


class WeibullSiteModel:  # site_model.py:140-247
    def __init__(self, shape, categories, invariant=None):
        self.shape, self.categories, self.invariant = shape, categories, invariant

    def rates(self):
        return orc.weibull_site_model(self.shape, self.categories, self.invariant)[0]

    def probabilities(self):
        return orc.weibull_site_model(self.shape, self.categories, self.invariant)[1]

    @property
    def sample_shape(self):
        return self.shape.shape[:-1]


class ConstantSiteModel:  # site_model.py:36-50
    def rates(self):
        return torch.ones(1, dtype=torch.float64)

    def probabilities(self):
        return torch.ones(1, dtype=torch.float64)

    sample_shape = torch.Size([])


class UserDefinedSubstitutionModel:
    """A model the engine knows nothing about (no reference base class in its MRO): its own p_t
    supplies the matrices and autograd carries d lnL / d P back (the "mats" route)."""

    def __init__(self, rates, freqs):
        self._rates, self._freqs = rates, freqs

    @property
    def frequencies(self):
        return self._freqs

    def p_t(self, t):
        q = orc.gtr_q_unnorm(self._rates, self._freqs)
        norm = -(torch.diagonal(q, dim1=-2, dim2=-1) * self._freqs).sum(-1)
        return orc.p_t_expm(q / norm[..., None, None], t)

    sample_shape = torch.Size([])


class SymmetricSubstitutionModel:  # substitution_model/abstract.py:53-85
    def norm(self, Q):
        return -torch.sum(torch.diagonal(Q, dim1=-2, dim2=-1) * self.frequencies, -1)

    def p_t(self, t):
        q = self.q()
        return orc.p_t_reversible(q / self.norm(q)[..., None, None], self.frequencies, t)


class GTR(SymmetricSubstitutionModel):  # nucleotide.py:274-380
    def __init__(self, rates, freqs):
        self._rates, self._freqs = rates, freqs

    @property
    def frequencies(self):
        return self._freqs

    def q(self):
        return orc.gtr_q_unnorm(self._rates, self._freqs)

    @property
    def sample_shape(self):
        return max(self._rates.shape[:-1], self._freqs.shape[:-1], key=len)


class HKY(SymmetricSubstitutionModel):  # nucleotide.py:170-271
    def __init__(self, kappa, freqs):
        self._kappa, self._freqs = kappa, freqs

    @property
    def frequencies(self):
        return self._freqs

    def q(self):
        return orc.hky_q_unnorm(self._kappa, self._freqs)

    sample_shape = torch.Size([])


class JC69:  # nucleotide.py:60-140
    frequencies = torch.full((4,), 0.25, dtype=torch.float64)
    sample_shape = torch.Size([])

    def q(self):
        q = torch.full((4, 4), 1.0 / 3, dtype=torch.float64)
        return q - torch.diag_embed(q.sum(-1)) * 1.0

    def p_t(self, t):
        return orc.p_t_jc69(t)


class NonSymmetricSubstitutionModel(SymmetricSubstitutionModel):  # abstract.py:88-97
    """Same generator as GTR but routed like the reference's matrix_exp models."""

    def __init__(self, rates, freqs):
        self._rates, self._freqs = rates, freqs

    @property
    def frequencies(self):
        return self._freqs

    def q(self):
        return orc.gtr_q_unnorm(self._rates, self._freqs)

    def p_t(self, t):
        q = self.q()
        return orc.p_t_expm(q / self.norm(q)[..., None, None], t)

    sample_shape = torch.Size([])
