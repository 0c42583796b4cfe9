"""CPU tests: the oracle against the reference-derived golden vectors and Appendix B values."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import synth_records


def test_oracle_matches_reference_machine_code(refcode_cases):
    """C restatement and numpy restatement == outputs of the reference's own CGR fill code."""
    n = 0
    for c in refcode_cases:
        rec = c["record"].encode("latin-1")
        for ks, e in c["k"].items():
            k = int(ks)
            v = O.cgr_count(rec, k, 16)
            assert hashlib.sha1(v.astype("<u2").tobytes()).hexdigest() == e["sha1"], (c["name"], k)
            assert int(v.astype(np.uint64).sum()) == e["sum"] and int(v.max()) == e["max"]
            if "vector" in e:
                assert v.tolist() == e["vector"]
            if k <= 10:
                assert np.array_equal(O.cgr_count_numpy(rec, k, 16), v), (c["name"], k)
            n += 1
    assert n >= 200


def test_oracle_u32_matches_reference_machine_code(refcode_cases_u32):
    """bits_per_element = 32: the C and numpy restatements == the outputs of the reference's uint32 fill
    (generation_cgr_windows_avx2.exe VA 0x14001b360), including counts above 65535 that the uint16 build wraps."""
    n, big = 0, 0
    for name, rec, ks in refcode_cases_u32:
        for kstr, e in ks.items():
            k = int(kstr)
            v = O.cgr_count(rec, k, 32)
            assert v.dtype == np.uint32
            assert hashlib.sha1(v.astype("<u4").tobytes()).hexdigest() == e["sha1"], (name, k)
            assert int(v.astype(np.uint64).sum()) == e["sum"] and int(v.max()) == e["max"]
            if "vector" in e:
                assert v.tolist() == e["vector"]
            if k <= 9 and len(rec) <= 20000:
                assert np.array_equal(O.cgr_count_numpy(rec, k, 32), v), (name, k)
            big += e["max"] > 65535
            n += 1
    assert n >= 100 and big >= 5


def test_appendix_b_demo_values(demo_records):
    # sha1(u16 LE)[:16] at k=6 and k=9, k=1 order [C,G,A,T]
    expect = {"A1.fasta": ("81ad32aac3d4d91d", "97378f11765d317a", [1505, 2030, 3216, 1897]),
              "A6.fasta": ("d3912d54417f4f24", "366650959ba2b19a", [1598, 2181, 3259, 1970]),
              "B.fasta": ("0c142e1dbe66a50d", "137c1930b20a0c3b", [1772, 2373, 3411, 2163]),
              "C.fasta": ("d4ee1cef6348a3ce", "9afa0c5742d030b9", [1555, 2155, 3276, 1973])}
    for name, rec in demo_records:
        h6 = hashlib.sha1(O.cgr_count(rec, 6).astype("<u2").tobytes()).hexdigest()[:16]
        h9 = hashlib.sha1(O.cgr_count(rec, 9).astype("<u2").tobytes()).hexdigest()[:16]
        assert (h6, h9) == expect[name][:2]
        assert O.cgr_count(rec, 1).tolist() == expect[name][2]


def test_appendix_b_distances(demo_records):
    x = np.stack([O.cgr_count(rec, 6) for _, rec in demo_records])
    man, inf = O.dists_triangle(x, True, True)
    assert man.tolist() == [3904, 4791, 4369, 4773, 4455, 4628]
    np.testing.assert_allclose(inf, [0.22702530, 0.25923625, 0.25204274, 0.24173106, 0.25185874, 0.24215384],
                               rtol=2e-7)
    f = np.stack([O.frequencies(r) for r in x])
    np.testing.assert_allclose(O.cdist_triangle(f, "euclid"),
                               [0.0107171533, 0.0122965513, 0.0118525572, 0.0119359576, 0.0116858869,
                                0.0116338804], rtol=1e-8)
    np.testing.assert_allclose(O.cdist_triangle(x, "cosine"),
                               [0.0938612871, 0.1247486, 0.114731872, 0.121487558, 0.114964415, 0.115188912],
                               rtol=1e-7)
    np.testing.assert_allclose(O.cdist_triangle(x, "pearson"),
                               [0.156244013, 0.209079337, 0.190899155, 0.208178386, 0.195207544, 0.197113458],
                               rtol=1e-7)


def test_u32_twin_and_wrap():
    rec = b"A" * 70000
    assert O.cgr_count(rec, 1, 16)[2] == 70000 % 65536
    assert O.cgr_count(rec, 1, 32)[2] == 70000
    f = O.frequencies(O.cgr_count(rec, 1, 16))
    assert f[2] == 1.0          # the divisor is the sum of the WRAPPED counts (backend.py:49)


def test_fasta_record0():
    assert O.fasta_record0(b">h\nACGT\nAC\n\n>h2\nGG\n") == b"ACGTAC"      # Q2: first record only
    assert O.fasta_record0(b"  ACG T \r\nNN\r\n") == b"ACG TNN"              # trim both ends, keep inner
    assert O.fasta_record0(b"\n\n>x\n>y\nTT") == b"TT"
    with pytest.raises(ValueError):
        O.fasta_record0(b">only header\n\n")


def test_manhat_quirks_and_scipy():
    rng = np.random.Generator(np.random.PCG64(7))
    a = rng.integers(0, 65536, 4096).astype(np.uint16)
    b = rng.integers(0, 65536, 4096).astype(np.uint16)
    assert O.manhat(a, b) == int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum()) % 2 ** 32
    fa, fb = O.frequencies(a), O.frequencies(b)
    assert O.manhat(fa, fb) == 0                                  # Q5: float inputs truncate to 0
    from scipy.spatial.distance import jaccard
    pa, pb = (a % 3 == 0), (b % 5 == 0)
    assert abs(float(O.info(pa.astype(np.uint16), pb.astype(np.uint16))) - jaccard(pa, pb)) < 1e-6
    assert O.info(np.zeros(8, np.uint16), np.zeros(8, np.uint16)) == 0.0


def test_batch_equals_single_and_threads():
    recs = synth_records(9, 700, seed=3, dirty=0.01, jitter=200)
    chars = np.frombuffer(b"".join(recs), dtype=np.uint8)
    start = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    for bits in (16, 32):
        out = O.cgr_batch(chars, start, 5, bits, threads=3)
        for i, r in enumerate(recs):
            assert np.array_equal(out[i], O.cgr_count(r, 5, bits))


def test_pyramid_identity(demo_records):
    """A.3: count_k == 2x2 sum-pool(count_{k+1}) + one fix-up k-mer (clean input)."""
    rec = demo_records[2][1]        # B.fasta is clean
    for k in range(1, 8):
        lo = O.cgr_count(rec, k, 32).reshape(2 ** k, 2 ** k).astype(np.int64)
        hi = O.cgr_count(rec, k + 1, 32).reshape(2 ** (k + 1), 2 ** (k + 1)).astype(np.int64)
        pooled = hi.reshape(2 ** k, 2, 2 ** k, 2).sum(axis=(1, 3))
        diff = lo - pooled
        assert diff.sum() == 1 and (diff != 0).sum() == 1


def test_oracle_dists_match_reference_machine_code(refcode_dists_cases):
    """manhat (uint32) and info (float32 bit patterns) == outputs of the reference's own row lambdas,
    for all six element types, every SIMD/scalar code path, uint32 wrap and the float quirk Q5."""
    assert len(refcode_dists_cases) >= 45
    seen = set()
    for name, x, man, inf in refcode_dists_cases:
        m, f = O.dists_triangle(x, True, True)
        assert np.array_equal(m, man), name
        assert np.array_equal(f.view(np.uint32), inf.view(np.uint32)), name
        seen.add(x.dtype.str)
    assert len(seen) == 6


def test_reference_machine_code_batch_equals_port():
    """bench.py's `kind: reference` CPU baseline runs the reference's own CGR-fill machine code (loaded from
    the reference executable here, from the blob under oracle/_ref on the GPU box) on host threads: same
    vectors as the port, clean and dirty records, several k."""
    if O.refcode() is None:
        pytest.skip("oracle/_ref holds no reference machine code on this host")
    from tests.helpers import synth_records
    recs = synth_records(24, 3000, seed=4, jitter=400, dirty=0.01) + [b"", b"ACG", b"NNACGTNACGT"]
    lens = np.array([len(r) for r in recs], dtype=np.uint64)
    start = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:])
    chars = np.frombuffer(b"".join(recs), dtype=np.uint8)
    for k in (3, 6, 9):
        got = np.zeros((len(recs), 4 ** k), dtype=np.uint16)
        O.refcode_cgr_batch(chars, start, k, got, threads=3)
        assert np.array_equal(got, O.cgr_batch(chars, start, k, 16)), k


def test_reference_row_task_machine_code_threaded_equals_port():
    """bench.py's `kind: reference` distances baseline: the reference's own manhat / info row tasks (loaded from
    the executable here, from the image under oracle/_ref on the GPU box) on host threads == the port, for every
    element type."""
    if O.refdists() is None:
        pytest.skip("oracle/_ref holds no reference machine code on this host")
    rng = np.random.Generator(np.random.PCG64(3))
    for dt in O.ELEM_DTYPES:
        x = (rng.random((33, 257)) * 40).astype(dt)
        x[4] = 0
        man, inf = O.refdists_triangle(x, True, True, threads=3)
        pm, pi = O.dists_triangle(x, True, True)
        assert np.array_equal(man, pm) and np.array_equal(inf.view(np.uint32), pi.view(np.uint32)), dt
