"""Host-side shape logic of torchtree_b200.flatten (no GPU, no torchtree): the broadcasting of
partially batched sub-model tensors and the route selection."""
import torch

from torchtree_b200.flatten import _flat, substitution_route


def test_flat_broadcasts_partial_batch_shapes():
    n, m, K = 3, 2, 4
    sample_shape = torch.Size([n, m])
    full = torch.arange(n * m * K, dtype=torch.float64).reshape(n, m, K)
    assert torch.equal(_flat(full, sample_shape, (K,)), full.reshape(n * m, K))
    shared = torch.arange(K, dtype=torch.float64)
    assert _flat(shared, sample_shape, (K,)).shape == (1, K)            # stays shared by all draws
    assert _flat(shared.reshape(1, 1, K), sample_shape, (K,)).shape == (1, K)
    inner = torch.arange(m * K, dtype=torch.float64).reshape(m, K)        # [m, K] under [n, m]
    got = _flat(inner, sample_shape, (K,))
    assert got.shape == (n * m, K)
    assert torch.equal(got, inner.expand(n, m, K).reshape(n * m, K))
    outer = torch.arange(n * K, dtype=torch.float64).reshape(n, 1, K)     # [n, 1, K] under [n, m]
    got = _flat(outer, sample_shape, (K,))
    assert torch.equal(got, outer.expand(n, m, K).reshape(n * m, K))
    mats = torch.zeros(n, m, 5, 5)
    assert _flat(mats, sample_shape, (5, 5)).shape == (n * m, 5, 5)


def test_substitution_routes():
    class SymmetricSubstitutionModel:
        pass

    class GTR(SymmetricSubstitutionModel):
        pass

    class NonSymmetricSubstitutionModel(SymmetricSubstitutionModel):
        pass

    class GeneralNonSymmetricSubstitutionModel(NonSymmetricSubstitutionModel):
        pass

    class Custom:
        pass

    assert substitution_route(GTR()) == "eigen"
    assert substitution_route(GeneralNonSymmetricSubstitutionModel()) == "expm"
    assert substitution_route(Custom()) == "mats"
