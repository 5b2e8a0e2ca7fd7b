"""Site-pattern compression through the native library.

`compress_sequences` is the native counterpart of the reference's `compress`
(torchtree/evolution/site_pattern.py:69-97); `tip_codes_from_alignment` goes all
the way to what the engine consumes (uint8 tip codes, code table, weights) and
equals `codes_from_tip_partials(compress_alignment(...))` on the reference side
(site_pattern.py:100-151).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import EngineError


def compress_sequences(sequences, group: int = 1):
    """Unique alignment columns in the reference's order and their counts.

    sequences: list of equal-length str / bytes (taxa order).  Returns
    (patterns uint8 [T, N, group], weights float64 [N])."""
    lib = _lib.load()
    rows = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in sequences]
    T = len(rows)
    L = len(rows[0])
    if any(len(r) != L for r in rows):
        raise EngineError("sequences must have equal lengths")
    seq = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(T, L)
    patterns = np.empty(T * L, dtype=np.uint8)
    weights = np.empty(L // group, dtype=np.float64)
    n = ctypes.c_int64(0)
    _lib.check(lib.ttb2_compress_patterns(
        seq.ctypes.data_as(ctypes.c_void_p), T, L, group,
        patterns.ctypes.data_as(ctypes.c_void_p), weights.ctypes.data_as(ctypes.c_void_p),
        ctypes.byref(n)), "ttb2_compress_patterns")
    N = int(n.value)
    return patterns[: T * N * group].reshape(T, N, group), weights[:N].copy()


def _select_sites(sequences, indices):
    """The reference's `indices` argument (site_pattern.py:85-89): ints and slices
    over the characters of each sequence, concatenated in the order given."""
    cols = []
    L = len(sequences[0])
    for index in indices:
        cols.append(np.arange(L)[index].reshape(-1))
    cols = np.concatenate(cols)
    out = []
    for s in sequences:
        row = np.frombuffer(s.encode("ascii") if isinstance(s, str) else bytes(s), dtype=np.uint8)
        out.append(row[cols].tobytes())
    return out


def compress_alignment_patterns(alignment, indices=None):
    """`compress` of the reference (site_pattern.py:69-97) on a torchtree
    Alignment: (patterns uint8 [T, N, group] with rows in `alignment.taxa` order,
    weights [N]).  The pattern ORDER is the reference's: it sorts the columns
    with the sequences in the alignment's own order (`zip(*alignment)`), which
    may differ from the Taxa order the rows are then looked up in."""
    group = alignment.data_type.size
    names = [s.taxon for s in alignment]
    sequences = [s.sequence for s in alignment]
    if indices is not None:
        sequences = _select_sites(sequences, indices)
    patterns, weights = compress_sequences(sequences, group)
    row_of = {name: i for i, name in enumerate(names)}
    rows = [row_of[t.id] for t in alignment.taxa]
    if rows != list(range(len(rows))):
        patterns = patterns[rows]
    return patterns, weights


def tip_codes_from_alignment(alignment, use_ambiguities: bool = False, indices=None):
    """(tip_codes uint8 [T,N], code_partials [C,S], weights [N]) for a torchtree
    Alignment, without Python loops over characters: the per-character work is a
    table lookup over the *distinct* site symbols only."""
    data_type = alignment.data_type
    S = data_type.state_count
    group = data_type.size
    patterns, weights = compress_alignment_patterns(alignment, indices)
    T, N, _ = patterns.shape
    # distinct symbols (strings of `group` characters) -> partial vectors -> engine codes
    # pack each site symbol into one integer so that `unique` is a 1-D sort
    if group > 8:
        raise EngineError("site symbols longer than 8 characters are not supported")
    if group == 1:
        present = np.zeros(256, dtype=bool)
        present[patterns.reshape(-1)] = True
        values = np.flatnonzero(present).astype(np.uint64)
        inverse = None
    else:
        packed = np.zeros(T * N, dtype=np.uint64)
        for g in range(group):
            packed = (packed << np.uint64(8)) | patterns[:, :, g].reshape(-1).astype(np.uint64)
        values, inverse = np.unique(packed, return_inverse=True)
    symbols = np.array([[(int(v) >> (8 * (group - 1 - g))) & 0xFF for g in range(group)]
                        for v in values], dtype=np.uint8).reshape(len(values), group)
    table = [tuple(row) for row in np.concatenate([np.eye(S), np.ones((1, S))], 0)]
    index = {v: i for i, v in enumerate(table)}
    lut = np.empty(len(symbols), dtype=np.uint8)
    for j, sym in enumerate(symbols):
        string = bytes(sym).decode("ascii")
        key = tuple(float(x) for x in data_type.partial(string, use_ambiguities))
        if key not in index:
            if len(table) >= 255:
                raise EngineError("more than 255 distinct tip partial vectors")
            index[key] = len(table)
            table.append(key)
        lut[j] = index[key]
    if inverse is None:
        lut256 = np.zeros(256, dtype=np.uint8)
        lut256[values.astype(np.int64)] = lut
        # bytes.translate is a plain C table walk (numpy's take would widen the indices)
        codes = np.frombuffer(patterns.tobytes().translate(lut256.tobytes()),
                              dtype=np.uint8).reshape(T, N)
    else:
        codes = lut[inverse.reshape(-1)].reshape(T, N)
    return codes, np.array(table, dtype=np.float64), weights
