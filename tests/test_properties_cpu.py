"""Property tests of the host logic and of the oracle (no GPU): formats round trips, triangle indexing, shard
and row-range partitions, FASTA record extraction (C entry point vs the oracle's restatement on random text),
and identities of the CGR counts the kernels rely on (wrap, pyramid pooling)."""
import ctypes

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from kameris_b200 import formats
from oracle import oracle as O

SET = settings(max_examples=60, deadline=None)


@SET
@given(st.integers(0, 5), st.integers(1, 7), st.integers(1, 40), st.integers(0, 2 ** 32 - 1))
def test_repr_round_trip_any_element_type(tmp_path_factory, code, count, cols, seed):
    dt = formats.ELEMENT_TYPES[code]
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = (rng.random((count, cols)) * 1000).astype(dt)
    p = str(tmp_path_factory.mktemp("rt") / "m.mm-repr")
    w = formats.repr_writer(p, mats[:1], count, create_file=True)
    for m in mats:
        w.write_matrix(m)
    w.close()
    with formats.repr_reader(p) as r:
        assert (r.count, r.rows, r.cols, r.value_type, r.is_sparse) == (count, 1, cols, code, False)
        assert np.array_equal(r.read_all(), mats)
        i = seed % count
        assert np.array_equal(r.read_matrix(i), mats[i:i + 1])


@SET
@given(st.integers(2, 200))
def test_triangle_index_is_the_row_major_strict_upper_triangle(n):
    iu = np.triu_indices(n, 1)
    idx = [formats.triangle_index(n, int(i), int(j)) for i, j in zip(*iu)]
    assert idx == list(range(n * (n - 1) // 2))


@SET
@given(st.lists(st.integers(0, 10 ** 7), min_size=0, max_size=60), st.integers(1, 9))
def test_shard_by_bases_partitions_in_order(lengths, world):
    from kameris_b200.parallel import shard_by_bases
    sh = shard_by_bases(lengths, world)
    assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == len(lengths)
    assert all(a <= b for a, b in sh) and all(sh[i][1] == sh[i + 1][0] for i in range(world - 1))
    total = sum(lengths)
    if total and lengths:
        worst = max(sum(lengths[a:b]) for a, b in sh)
        assert worst <= total / world + max(lengths)          # no shard exceeds its share by more than one sequence


@SET
@given(st.integers(2, 5000), st.integers(1, 16))
def test_balanced_row_ranges_cover_the_triangle(n, parts):
    from kameris_b200.dist import balanced_row_ranges, tri_len
    from kameris_b200.parallel import tri_row_offset
    rr = balanced_row_ranges(n, parts)
    assert rr[0][0] == 0 and rr[-1][1] == n and all(rr[i][1] == rr[i + 1][0] for i in range(parts - 1))
    sizes = [tri_row_offset(n, b) - tri_row_offset(n, a) for a, b in rr]
    assert sum(sizes) == tri_len(n) + 0 * n and all(s >= 0 for s in sizes)
    assert max(sizes) <= tri_len(n) / parts + n              # within one row of the ideal share
    for a, _ in rr:
        if a + 1 < n:
            assert tri_row_offset(n, a) == formats.triangle_index(n, a, a + 1)


TEXT = st.text(alphabet="ACGTNacgt> \t\r\n-*", min_size=0, max_size=300)


@SET
@given(TEXT)
def test_fasta_record0_c_entry_point_equals_oracle(text):
    from kameris_b200 import _lib
    lib = _lib.load()
    data = text.encode()
    out = ctypes.create_string_buffer(max(len(data), 1))
    n = lib.kam_fasta_record0(data, len(data), out)
    # A.1 in plain Python: lines split at '\n', trimmed of " \t\n\v\f\r"; a blank line or a '>' line closes the
    # record being collected; the first record wins
    cur, first = [], None
    for line in data.split(b"\n"):
        line = line.strip(b" \t\n\v\f\r")
        if not line or line[:1] == b">":
            if cur and first is None:
                first = b"".join(cur)
            cur = []
        else:
            cur.append(line)
    if first is None and cur:
        first = b"".join(cur)
    if first is None:
        assert n < 0
        with pytest.raises(ValueError):
            O.fasta_record0(data)
    else:
        assert n == len(first) and out.raw[:n] == first and O.fasta_record0(data) == first


SEQ = st.text(alphabet="ACGTN", min_size=0, max_size=400).map(str.encode)


@SET
@given(SEQ, st.integers(1, 6))
def test_cgr_counts_sum_and_16_bit_wrap(rec, k):
    c32 = O.cgr_count(rec, k, 32)
    c16 = O.cgr_count(rec, k, 16)
    assert np.array_equal(c16, (c32 & 0xffff).astype(np.uint16))
    bases = sum(1 for ch in rec if ch in b"ACGT")
    assert int(c32.sum()) <= bases                         # one increment per emitting base at most
    if b"N" not in rec and bases >= k:
        assert int(c32.sum()) == bases - k + 1             # clean record: every full k-mer once


@SET
@given(st.text(alphabet="ACGT", min_size=8, max_size=400).map(str.encode), st.integers(1, 5))
def test_pyramid_identity_for_clean_records(rec, k):
    """the order-k table is the 2x2 sum-pooling of the order-(k+1) CGR image plus the one k-mer
    (the first) that has no (k+1)-mer ending at the same base."""
    hi = O.cgr_count(rec, k + 1, 32).reshape(2 ** (k + 1), 2 ** (k + 1))
    lo = O.cgr_count(rec, k, 32)
    pooled = hi.reshape(2 ** k, 2, 2 ** k, 2).sum(axis=(1, 3)).reshape(-1)
    diff = lo.astype(np.int64) - pooled.astype(np.int64)
    assert diff.sum() == 1 and (diff >= 0).all() and np.count_nonzero(diff) == 1
