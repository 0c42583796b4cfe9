"""Host half of the sparse return path of kam_kmers_host (no GPU): kam_sparse_expand turns sorted
(bin << 8 | count) lists into the dense rows generation_cgr + the numpy pass (backend.py:45-49) produce.
The lists here are made from the oracle's count vectors."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import synth_records


def _lists(vectors):
    ent, start, nnz = [], [0], []
    for v in vectors:
        b = np.nonzero(v)[0].astype(np.uint32)
        ent.append((b << np.uint32(8)) | v[b].astype(np.uint32))
        nnz.append(len(b))
        start.append(start[-1] + len(b) + 3)          # slack between lists, like the character offsets
        ent.append(np.zeros(3, dtype=np.uint32))
    return np.concatenate(ent), np.array(start[:-1], dtype=np.uint64), np.array(nnz, dtype=np.uint32)


@pytest.mark.parametrize("k", [7, 9])
@pytest.mark.parametrize("threads", [1, 5])
def test_expand_matches_oracle_and_numpy(k, threads):
    from kameris_b200 import _lib, repr as R
    lib = _lib.load()
    recs = synth_records(9, 3000, seed=k, dirty=0.002, jitter=2500) + [b"", b"NNNN", b"ACGTACGTAC", b"T" * 200]
    counts = np.stack([O.cgr_count(r, k, 16) for r in recs])
    assert counts.max() <= 255
    ent, start, nnz = _lists(counts)
    lib.kam_host_set_threads(threads)
    try:
        assert lib.kam_host_threads() == threads
        got16 = R.sparse_expand(ent, start, nnz, k, 16, "counts")
        got32 = R.sparse_expand(ent, start, nnz, k, 32, "counts")
        f64 = R.sparse_expand(ent, start, nnz, k, 16, "freq64")
        f32 = R.sparse_expand(ent, start, nnz, k, 16, "freq32")
    finally:
        lib.kam_host_set_threads(0)
    assert np.array_equal(got16, counts) and np.array_equal(got32, counts.astype(np.uint32))
    with np.errstate(invalid="ignore", divide="ignore"):
        for i, c in enumerate(counts):
            want = c.reshape(1, -1) / c.reshape(1, -1).sum()          # backend.py:49, executed by numpy
            assert np.array_equal(f64[i], want[0], equal_nan=True), i
            assert np.array_equal(f32[i], want[0].astype(np.float32), equal_nan=True), i
    assert np.isnan(f64[len(recs) - 4]).all()                          # empty record: 0/0 everywhere, like numpy


def test_expand_misaligned_rows_and_small_k():
    """rows that are not 64-byte multiples / 16-byte aligned take the plain-store path"""
    from kameris_b200 import repr as R
    recs = [b"ACGTTGCAAC", b"GATTACA" * 30]
    for k in (1, 2, 3, 5):
        counts = np.stack([O.cgr_count(r, k, 16) for r in recs])
        ent, start, nnz = _lists(counts)
        assert np.array_equal(R.sparse_expand(ent, start, nnz, k, 16, "counts"), counts)
        f = R.sparse_expand(ent, start, nnz, k, 16, "freq64")
        assert np.array_equal(f, counts / counts.sum(axis=1, keepdims=True))


def test_expand_rejects_bad_lists():
    from kameris_b200 import _lib, repr as R
    bad = np.array([(5 << 8) | 1, (5 << 8) | 2], dtype=np.uint32)              # not ascending
    with pytest.raises(_lib.KamError):
        R.sparse_expand(bad, np.array([0], dtype=np.uint64), np.array([2], dtype=np.uint32), 7)
    with pytest.raises(_lib.KamError):
        R.sparse_expand(np.array([(4 ** 7) << 8 | 1], dtype=np.uint32), np.array([0], dtype=np.uint64),
                        np.array([1], dtype=np.uint32), 7)                      # bin out of range
    with pytest.raises(_lib.KamError):
        R.sparse_expand(np.zeros(1, dtype=np.uint32), np.array([0], dtype=np.uint64),
                        np.array([0xffffffff], dtype=np.uint32), 7)             # a declined vector


def test_expand_scalar_bodies_still_agree():
    """count rows take the AVX-512 VPEXPAND path where the CPU has it; KAM_HOST_EXPAND=scalar (read once per
    process) keeps the line-composition bodies: the same tests in a child process with it set"""
    import os
    import subprocess
    import sys
    if os.environ.get("KAM_HOST_EXPAND") == "scalar":
        pytest.skip("already the child")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_sparse_cpu.py"), "-x", "-q",
                        "-k", "matches_oracle or misaligned"], cwd=root, env=dict(os.environ, KAM_HOST_EXPAND="scalar"),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-2000:]


@pytest.mark.parametrize("k,maxc", [(6, 3), (8, 15), (8, 16), (9, 31), (9, 32), (10, 255)])
def test_expand_line_paths_dense_clusters_and_lookup_limits(k, maxc):
    """the AVX-512 line paths (vpexpand of a line's occupancy mask and its consecutive entries): lines with more
    than 16 entries (a fully occupied uint16 line has 32), empty lines, the last line of a row, and largest counts
    right at and past what the two-register quotient lookups hold (15 / 16 for float64, 31 / 32 for float32)"""
    from kameris_b200 import repr as R
    rng = np.random.Generator(np.random.PCG64(1000 * k + maxc))
    bins = 4 ** k
    rows = []
    for kind in range(5):
        v = np.zeros(bins, dtype=np.uint32)
        if kind == 0:                                   # sparse, like a k-mer vector
            idx = rng.choice(bins, size=min(bins // 30 + 1, 9000), replace=False)
        elif kind == 1:                                 # dense clusters: whole lines occupied
            starts = rng.choice(bins // 64, size=5, replace=False) * 64
            idx = np.concatenate([np.arange(s, s + 64) for s in starts] + [np.array([0, 1, bins - 1])])
        elif kind == 2:                                 # every second bin of a stretch: 16 per uint16 line, 8 per uint32 line
            idx = np.arange(128, 128 + 512, 2)
        elif kind == 3:                                 # one entry only, in the last bin
            idx = np.array([bins - 1])
        else:                                           # nothing at all (0/0 rows for the frequencies)
            idx = np.array([], dtype=np.int64)
        v[idx] = rng.integers(1, maxc + 1, size=len(idx))
        if len(idx):
            v[idx[0]] = maxc
        rows.append(v)
    counts = np.stack(rows)
    ent, start, nnz = _lists(counts)
    assert np.array_equal(R.sparse_expand(ent, start, nnz, k, 16, "counts"), counts.astype(np.uint16))
    assert np.array_equal(R.sparse_expand(ent, start, nnz, k, 32, "counts"), counts)
    f64 = R.sparse_expand(ent, start, nnz, k, 16, "freq64")
    f32 = R.sparse_expand(ent, start, nnz, k, 16, "freq32")
    with np.errstate(invalid="ignore", divide="ignore"):
        for i, c in enumerate(counts):
            want = c.astype(np.float64) / c.sum()
            assert np.array_equal(f64[i], want, equal_nan=True), i
            assert np.array_equal(f32[i], want.astype(np.float32), equal_nan=True), i
