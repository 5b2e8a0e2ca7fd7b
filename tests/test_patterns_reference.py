"""Native site-pattern compression == the reference's `compress` family
(torchtree/evolution/site_pattern.py:69-151).  Host code: runs without a GPU.
Needs the reference importable: `baseline/_ref` (tools/vendor_reference.py; travels to the GPU box,
where the "box" variants run under `-m gpu`) or /root/reference (authoring container)."""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refenv  # noqa: E402

REF = os.path.dirname(refenv.data_dir() or "/root/reference/data")


@pytest.fixture(scope="module", params=["host", pytest.param("box", marks=pytest.mark.gpu)])
def ref(request):
    refenv.activate()
    import torchtree  # noqa: F401
    from torchtree.evolution import alignment, datatype, site_pattern, taxa
    return dict(alignment=alignment, datatype=datatype, site_pattern=site_pattern, taxa=taxa)


def _alignment(ref, names, seqs, data_type, taxa_order=None):
    Taxa, Taxon = ref["taxa"].Taxa, ref["taxa"].Taxon
    Alignment, Sequence = ref["alignment"].Alignment, ref["alignment"].Sequence
    order = taxa_order if taxa_order is not None else names
    taxa = Taxa("taxa", [Taxon(n, {}) for n in order])
    return Alignment("aln", [Sequence(n, s) for n, s in zip(names, seqs)], taxa, data_type)


def _read_fasta(path):
    names, seqs = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            names.append(line[1:].split()[0])
            seqs.append("")
        elif line:
            seqs[-1] += line
    return names, seqs


def _check(ref, aln, use_ambiguities, indices=None):
    from torchtree_b200.engine import codes_from_tip_partials
    from torchtree_b200.patterns import compress_alignment_patterns, tip_codes_from_alignment
    sp = ref["site_pattern"]
    ref_patterns, ref_weights = sp.compress(aln, indices)
    patterns, weights = compress_alignment_patterns(aln, indices)
    assert weights.tolist() == ref_weights.tolist()
    for row, taxon in zip(patterns, aln.taxa):
        got = [bytes(sym).decode() for sym in row]
        want = ["".join(sym) for sym in ref_patterns[taxon.id]]
        assert got == want
    # all the way to what the engine consumes
    torch.set_default_dtype(torch.float64)
    try:
        partials, w2 = sp.compress_alignment(aln, indices, use_ambiguities)
    finally:
        torch.set_default_dtype(torch.float32)
    want_codes, want_table = codes_from_tip_partials(
        [p.numpy() for p in partials], aln.data_type.state_count)
    codes, table, w = tip_codes_from_alignment(aln, use_ambiguities, indices)
    assert w.tolist() == w2.tolist()
    # same partial vector at every (tip, pattern), whatever the code numbering
    assert np.array_equal(table[codes], want_table[want_codes])
    return patterns.shape[1]


@pytest.mark.reference
@pytest.mark.parametrize("use_ambiguities", [False, True])
def test_fluA(ref, use_ambiguities):
    names, seqs = _read_fasta(os.path.join(REF, "data", "fluA.fa"))
    aln = _alignment(ref, names, seqs, ref["datatype"].NucleotideDataType("nuc"))
    assert _check(ref, aln, use_ambiguities) == 238     # 987 sites -> 238 patterns


@pytest.mark.reference
def test_taxa_order_differs_from_alignment_order(ref):
    rng = random.Random(3)
    names = [f"t{i}" for i in range(9)]
    seqs = ["".join(rng.choice("ACGTRYN-acgt?") for _ in range(400)) for _ in names]
    order = names[:]
    rng.shuffle(order)
    aln = _alignment(ref, names, seqs, ref["datatype"].NucleotideDataType("nuc"), order)
    _check(ref, aln, True)
    _check(ref, aln, False)


@pytest.mark.reference
def test_indices(ref):
    rng = random.Random(4)
    names = [f"t{i}" for i in range(6)]
    seqs = ["".join(rng.choice("ACGT") for _ in range(300)) for _ in names]
    aln = _alignment(ref, names, seqs, ref["datatype"].NucleotideDataType("nuc"))
    _check(ref, aln, False, [slice(0, None, 3), slice(1, None, 3)])
    _check(ref, aln, False, [slice(10, 50), 7, slice(100, 300, 2)])


@pytest.mark.reference
def test_amino_acids(ref):
    rng = random.Random(5)
    names = [f"t{i}" for i in range(7)]
    seqs = ["".join(rng.choice("ARNDCQEGHILKMFPSTWYVBZX-") for _ in range(500)) for _ in names]
    aln = _alignment(ref, names, seqs, ref["datatype"].AminoAcidDataType("aa"))
    _check(ref, aln, True)
    _check(ref, aln, False)


@pytest.mark.reference
def test_codons(ref):
    rng = random.Random(6)
    names = [f"t{i}" for i in range(5)]
    codons = ["ATG", "AAA", "CCC", "GGT", "TTC", "N--", "---", "ACN", "GAT"]
    seqs = ["".join(rng.choice(codons) for _ in range(300)) for _ in names]
    aln = _alignment(ref, names, seqs, ref["datatype"].CodonDataType("codon", "Universal"))
    _check(ref, aln, False)


def test_argument_validation():
    from torchtree_b200._lib import EngineError
    from torchtree_b200.patterns import compress_sequences
    with pytest.raises(EngineError):
        compress_sequences(["ACG", "AC"])
    with pytest.raises(EngineError):
        compress_sequences(["ACGT", "ACGT"], group=3)
    p, w = compress_sequences(["AAAA", "CCCC"])
    assert p.shape == (2, 1, 1) and w.tolist() == [4.0]


@pytest.mark.reference
def test_fuzz_small_alignments(ref):
    """Edge shapes against the reference's `compress`: one site, one / two taxa, identical
    columns, all-distinct columns, low-entropy alphabets (many repeated columns)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    dt = ref["datatype"].NucleotideDataType("nuc")

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(st.integers(1, 5), st.integers(1, 40), st.sampled_from(["AC", "ACGT", "ACGT-RN", "A"]),
           st.integers(0, 2 ** 31 - 1))
    def run(taxa, length, alphabet, seed):
        rng = random.Random(seed)
        names = [f"t{i}" for i in range(taxa)]
        seqs = ["".join(rng.choice(alphabet) for _ in range(length)) for _ in names]
        aln = _alignment(ref, names, seqs, dt)
        n = _check(ref, aln, True)
        assert 1 <= n <= length

    run()
