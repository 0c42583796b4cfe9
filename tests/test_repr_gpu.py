"""GPU parity: pack + kmer_count CUDA kernels (through the C ABI) vs the oracle and the golden
vectors produced by the reference's own machine code.  Bit-exact for counts and float64
frequencies; float32 frequencies within 1e-6 relative (north_star tolerance: 1e-5)."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import HIV_P, synth_records

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def R():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200 import repr as R
    return R


def _gpu_counts(R, recs, k, bits=16, out="counts"):
    b = R.pack_records(recs)
    return R.kmer_vectors(b, k, bits, out).cpu().numpy()


def test_pack_layout(R):
    recs = [b"NACGTTGCA", b"", b"acgtACGT", b"T" * 100, b"NNNN", b"GATTACA" * 700]
    b = R.pack_records(recs)
    base = b.base_start.cpu().numpy()
    planes = b.planes.cpu().numpy().view(np.uint64)
    head = b.head_mask.cpu().numpy().view(np.uint32)
    pos = 0
    for s, rec in enumerate(recs):
        valid = [c for c in rec if c in b"ACGT"]
        assert base[s] == pos and base[s + 1] - base[s] == len(valid)
        for j, c in enumerate(valid):
            w, bit = divmod(pos + j, 32)
            x = (int(planes[w]) >> bit) & 1
            y = (int(planes[w]) >> (32 + bit)) & 1
            assert x == (c in b"GT") and y == (c in b"AT")
        hm = sum(1 << i for i, c in enumerate(rec[:32]) if c in b"ACGT")
        assert head[s] == hm
        pos += len(valid)


@pytest.fixture(params=[1, 0], ids=["single-pass", "three-launch"])
def pack_mode(request):
    from kameris_b200 import _lib
    _lib.load().kam_pack_set_fused(request.param)
    yield request.param
    _lib.load().kam_pack_set_fused(1)


def test_pack_large_many_tiles(R, pack_mode):
    """40 MB of characters (10 000 tiles: the chained scan's look-back crosses many waves of CTAs) against
    numpy: per-sequence base offsets and a checksum of the planes."""
    rng = np.random.Generator(np.random.PCG64(9))
    n, L = 4000, 10000
    alpha = np.frombuffer(b"ACGTN", dtype=np.uint8)
    allc = alpha[rng.choice(5, size=n * L, p=[0.245, 0.245, 0.245, 0.245, 0.02])]
    chars = torch.zeros(n * L + 16, dtype=torch.uint8, device="cuda")
    chars[:n * L] = torch.from_numpy(allc).cuda()
    start = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
    b = R.pack_chars(chars, start, n, n * L)
    valid = allc != ord("N")
    cum = np.concatenate([[0], np.cumsum(valid)])
    assert np.array_equal(b.base_start.cpu().numpy(), cum[np.arange(n + 1) * L])
    vb = allc[valid]
    nw = len(vb) // 32
    xb = np.isin(vb[:nw * 32], np.frombuffer(b"GT", dtype=np.uint8)).reshape(nw, 32)
    yb = np.isin(vb[:nw * 32], np.frombuffer(b"AT", dtype=np.uint8)).reshape(nw, 32)
    want = np.packbits(xb, axis=1, bitorder="little").view(np.uint32)[:, 0].astype(np.uint64) | \
        (np.packbits(yb, axis=1, bitorder="little").view(np.uint32)[:, 0].astype(np.uint64) << np.uint64(32))
    got = b.planes.cpu().numpy().view(np.uint64)[:nw]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pack_random_layouts(R, seed, pack_mode):
    """pack kernels against a numpy restatement of the planes layout: random record lengths including
    empty records (leading, trailing, consecutive), records that start exactly on 4096-character tile
    boundaries, records spanning many tiles, dirty characters, totals that are not multiples of 16."""
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGTNacgtRY-", dtype=np.uint8)
    p = np.array([0.24, 0.24, 0.24, 0.24] + [0.005] * 8)
    p /= p.sum()
    lens = []
    for _ in range(int(rng.integers(1, 200))):
        kind = rng.integers(0, 6)
        lens.append(int({0: 0, 1: rng.integers(1, 40), 2: rng.integers(100, 300), 3: 4096 * rng.integers(1, 4),
                         4: rng.integers(5000, 30000), 5: 4096 - 7}[int(kind)]))
    if seed == 1:
        lens = [0, 0, 4096, 0, 8192, 1, 0] + lens + [0, 0]
    if seed == 2:
        lens = lens + [4096 * 3 - sum(lens) % 4096]             # total a multiple of the tile size
    recs = [alpha[rng.choice(len(alpha), size=n, p=p)].tobytes() for n in lens]
    b = R.pack_records(recs)
    base = b.base_start.cpu().numpy()
    planes = b.planes.cpu().numpy().view(np.uint64)
    head = b.head_mask.cpu().numpy().view(np.uint32)
    allc = np.frombuffer(b"".join(recs), dtype=np.uint8)
    valid = np.isin(allc, np.frombuffer(b"ACGT", dtype=np.uint8))
    starts = np.concatenate([[0], np.cumsum(lens)])
    cum = np.concatenate([[0], np.cumsum(valid)])
    assert np.array_equal(base, cum[starts])
    vb = allc[valid]
    xbits = np.isin(vb, np.frombuffer(b"GT", dtype=np.uint8)).astype(np.uint64)
    ybits = np.isin(vb, np.frombuffer(b"AT", dtype=np.uint8)).astype(np.uint64)
    nw = (len(vb) + 31) // 32
    pad = nw * 32 - len(vb)
    xw = np.concatenate([xbits, np.zeros(pad, np.uint64)]).reshape(nw, 32)
    yw = np.concatenate([ybits, np.zeros(pad, np.uint64)]).reshape(nw, 32)
    sh = np.arange(32, dtype=np.uint64)
    want = (xw << sh).sum(axis=1) | ((yw << sh).sum(axis=1) << np.uint64(32))
    assert np.array_equal(planes[:nw], want)
    for s_i, rec in enumerate(recs):
        hm = sum(1 << i for i, c in enumerate(rec[:32]) if c in b"ACGT")
        assert head[s_i] == hm, s_i


def test_golden_reference_machine_code(R, refcode_cases):
    """Every golden case (quirks Q1/Q3/Q4, demo genomes, random dirty records), k <= 10."""
    ks = sorted({int(k) for c in refcode_cases for k in c["k"] if int(k) <= 10})
    for k in ks:
        cases = [c for c in refcode_cases if str(k) in c["k"]]
        recs = [c["record"].encode("latin-1") for c in cases]
        got = _gpu_counts(R, recs, k, 16)
        for i, c in enumerate(cases):
            e = c["k"][str(k)]
            assert hashlib.sha1(got[i].astype("<u2").tobytes()).hexdigest() == e["sha1"], (c["name"], k)


def test_golden_reference_machine_code_u32(R, refcode_cases_u32):
    """bits_per_element = 32 against the reference's own uint32 fill (VA 0x14001b360): every golden case,
    including counts above 65535 (paths A, L and B all meet here: 70 000 x 'A', 400 000 bases of 'AC')."""
    ks = sorted({int(k) for _, _, d in refcode_cases_u32 for k in d if int(k) <= 11})
    for k in ks:
        cases = [(n, r, d[str(k)]) for n, r, d in refcode_cases_u32 if str(k) in d]
        got = _gpu_counts(R, [r for _, r, _ in cases], k, 32)
        for i, (name, _, e) in enumerate(cases):
            assert got.dtype == np.uint32
            assert hashlib.sha1(got[i].astype("<u4").tobytes()).hexdigest() == e["sha1"], (name, k)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("bits", [16, 32])
def test_demo_and_random_vs_oracle(R, demo_records, k, bits):
    recs = [r for _, r in demo_records] + synth_records(40, 9000, seed=20261002 + k, dirty=0.002, p=HIV_P,
                                                        jitter=900)
    got = _gpu_counts(R, recs, k, bits)
    for i, r in enumerate(recs):
        assert np.array_equal(got[i], O.cgr_count(r, k, bits)), (i, k, bits)


@pytest.mark.parametrize("k", [3, 6, 9])
def test_frequencies(R, demo_records, k):
    recs = [r for _, r in demo_records] + synth_records(10, 5000, seed=5 + k, dirty=0.01)
    f64 = _gpu_counts(R, recs, k, 16, "freq64")
    f32 = _gpu_counts(R, recs, k, 16, "freq32")
    for i, r in enumerate(recs):
        ref = O.frequencies(O.cgr_count(r, k, 16))
        assert np.array_equal(f64[i], ref)                      # bit-exact numpy true division
        np.testing.assert_allclose(f32[i], ref, rtol=1e-6, atol=0)


def test_short_reads_k6(R):
    """BASELINE config 4 shape: 150 bp reads, k=6 counts (ragged, some shorter than k, some empty)."""
    recs = synth_records(3000, 150, seed=20261004, jitter=20) + [b"", b"ACG", b"NNNNNN", b"ACGTAC"]
    got = _gpu_counts(R, recs, 6, 16)
    chars = np.frombuffer(b"".join(recs), dtype=np.uint8)
    start = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
    assert np.array_equal(got, O.cgr_batch(chars, start, 6, 16))


@pytest.mark.parametrize("out,bits", [("counts", 16), ("counts", 32), ("freq32", 16), ("freq64", 16), ("freq64", 32)])
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_short_reads_warp_kernel(R, k, out, bits):
    """The warp-per-read kernel (short reads, k <= 6): ragged and dirty reads, empty reads, reads
    shorter than k, reads past 255 bases (checked uint8 counters), reads whose counters overflow and
    are redone on the spill path, a read long enough to go there directly; every output kind.  The
    CTA-per-read kernel (warp path off) must give the same bytes."""
    recs = synth_records(700, 150, seed=31 + k, jitter=40, dirty=0.01) + \
        [b"", b"ACG", b"NNNNNN", b"ACGTAC", b"NNACGTACGTAC", b"N" * 40 + b"ACGTTGCAAC", b"A" * 400, b"ACGT" * 200,
         synth_records(1, 300, seed=5)[0], synth_records(1, 3000, seed=6)[0], b"AC" * 600, b"T" * 255, b"T" * 256,
         b"G" * 261, synth_records(1, 70000 if k < 4 else 9000, seed=8)[0]] + synth_records(300, 100, seed=9, jitter=30)
    from kameris_b200 import _lib
    got = _gpu_counts(R, recs, k, bits, out)
    _lib.load().kam_kmer_set_warp_path(0)
    try:
        alt = _gpu_counts(R, recs, k, bits, out)
    finally:
        _lib.load().kam_kmer_set_warp_path(1)
    assert got.tobytes() == alt.tobytes()
    for i, r in enumerate(recs):
        c = O.cgr_count(r, k, bits)
        if out == "counts":
            assert np.array_equal(got[i], c), (k, i)
        else:
            with np.errstate(all="ignore"):
                ref = O.frequencies(c)
            if out == "freq64":
                assert np.array_equal(got[i], ref, equal_nan=True), (k, i)
            else:
                ok = np.isfinite(ref)
                np.testing.assert_allclose(got[i][ok], ref[ok], rtol=1e-6, atol=0, err_msg=str((k, i)))


@pytest.mark.parametrize("out,bits", [("counts", 16), ("counts", 32), ("freq64", 16)])
@pytest.mark.parametrize("k", [2, 5, 6, 7, 8, 9, 10])
def test_tile_kernel_scan_variants(R, k, out, bits):
    """The CTA-per-sequence kernel scans a whole plane word (32 k-mer end positions) per thread and trip;
    the first version (8 positions, kam_kmer_set_word_scan(0)) must give the same bytes, and both the
    oracle's.  Ragged genome lengths around the word / tile boundaries, dirty heads, sequences shorter
    than k and empty ones in between so every sequence starts at a different bit of a plane word."""
    recs = synth_records(40, 9000, seed=70 + k, jitter=700, dirty=0.002) + \
        [b"", b"ACG", b"ACGTACGTAC", b"N" * 5 + b"ACGTTGCAACGT", b"T" * 31, b"G" * 32, b"C" * 33, b"A" * 64,
         b"NNA" + synth_records(1, 20000, seed=3)[0], synth_records(1, 100000, seed=4)[0], b"ACGT" * 16 + b"A"] + \
        synth_records(20, 2500, seed=71, jitter=31)
    from kameris_b200 import _lib
    got = _gpu_counts(R, recs, k, bits, out)
    _lib.load().kam_kmer_set_word_scan(0)
    try:
        alt = _gpu_counts(R, recs, k, bits, out)
    finally:
        _lib.load().kam_kmer_set_word_scan(1)
    assert got.tobytes() == alt.tobytes()
    for i, r in enumerate(recs):
        c = O.cgr_count(r, k, bits)
        if out == "counts":
            assert np.array_equal(got[i], c), (k, i)
        else:
            with np.errstate(all="ignore"):
                assert np.array_equal(got[i], O.frequencies(c), equal_nan=True), (k, i)


@pytest.mark.parametrize("long_path", [2, 1, 0])
@pytest.mark.parametrize("k,bits", [(9, 16), (9, 32), (8, 16), (7, 16), (2, 16), (10, 16), (11, 32), (12, 16)])
def test_long_sequences_spill_path(R, k, bits, long_path):
    """Sequences beyond path A (length, or uint8 overflow) take path L -- for k <= 9 the one-pass kernel with
    the whole table in shared memory (long_path=2: uint16 counters below k = 9, 6-bit fields + side table at
    k = 9), else / behind it the tiled uint16 kernel (long_path=1 forces it) -- and, when even a uint16 counter
    overflows (70 000 x 'A'), the RED.ADD path B; long_path=0 sends all of them to path B directly.  The
    homopolymer and the dinucleotide repeat make the 6-bit kernel give up (every increment of a warp lands in
    one field), so the whole chain of fall-backs runs.  Dirty heads included."""
    recs = synth_records(3, 400000, seed=11 + k, dirty=0.001) + [b"ACGT" * 3000, b"A" * 70000,
                                                                 synth_records(1, 9000, seed=2)[0],
                                                                 b"NNA" + synth_records(1, 200000, seed=3)[0],
                                                                 b"AC" * 140000]
    from kameris_b200 import _lib
    _lib.load().kam_kmer_set_long_path(1 if long_path else 0)
    _lib.load().kam_kmer_set_long_one_pass(2 if long_path == 2 else 0)   # 2: however few the sequences
    try:
        got = _gpu_counts(R, recs, k, bits)
    finally:
        _lib.load().kam_kmer_set_long_path(1)
        _lib.load().kam_kmer_set_long_one_pass(1)
    for i, r in enumerate(recs):
        assert np.array_equal(got[i], O.cgr_count(r, k, bits)), (i, k, bits)


@pytest.mark.parametrize("out,bits", [("counts", 16), ("counts", 32), ("freq64", 16)])
@pytest.mark.parametrize("k", [9, 8])
def test_long_one_pass_hot_bins_and_many_sequences(R, out, bits, k):
    """The one-pass path L kernel where its 6-bit fields matter (k = 9): composition-skewed sequences whose
    hottest bins hold hundreds (every 32 increments of a bin move into the side table), more sequences than
    CTAs (persistent loop, side table and block bitmap re-used clean), ragged ends that are not multiples of
    32, a tandem repeat inside an ordinary sequence (a burst on a few bins), sequences of one plane word and
    less than k bases in between, and short sequences path A keeps for itself.  Bit-exact against the oracle,
    and the tiled kernel must give the same bytes."""
    rng = np.random.Generator(np.random.PCG64(5 + k))
    recs = []
    for i in range(170):
        L = int(rng.integers(131100, 140000)) if i % 9 else int(rng.integers(300000, 420000))
        p = [[0.4, 0.1, 0.1, 0.4], [0.15, 0.35, 0.35, 0.15], None][i % 3]
        recs.append(synth_records(1, L, seed=1000 + i, dirty=0.0005 if i % 4 == 0 else 0.0, p=p)[0])
    r = bytearray(recs[5])
    r[70000:70900] = b"ACG" * 300
    recs[5] = bytes(r)
    recs[7] = recs[7][:60000] + b"T" * 45 + recs[7][60000:]
    recs += [b"ACGTTGCA", b"", b"ACGTACGTACGTAC", synth_records(1, 5000, seed=9)[0]]
    from kameris_b200 import _lib
    got = _gpu_counts(R, recs, k, bits, out)
    _lib.load().kam_kmer_set_long_one_pass(0)
    try:
        tiled = _gpu_counts(R, recs, k, bits, out)
    finally:
        _lib.load().kam_kmer_set_long_one_pass(1)
    assert np.array_equal(got, tiled, equal_nan=True)
    for i in list(range(0, len(recs), 7)) + [5, 7] + list(range(len(recs) - 4, len(recs))):
        c = O.cgr_count(recs[i], k, bits)
        if out == "counts":
            assert np.array_equal(got[i], c), (i, k, bits)
        else:
            with np.errstate(all="ignore"):
                assert np.array_equal(got[i], O.frequencies(c), equal_nan=True), (i, k)


@pytest.mark.parametrize("out", ["freq32", "freq64"])
@pytest.mark.parametrize("k", [9, 10])
def test_long_sequences_frequencies(R, out, k):
    """path L with the fused frequency division (BASELINE config 5 shape, scaled down)."""
    recs = synth_records(2, 600000, seed=77, dirty=0.0005) + [b"A" * 70000, b"AC" * 140000, b"NNNNNNNN" + b"GATTACA" * 30000]
    for bits in (16, 32):
        got = _gpu_counts(R, recs, k, bits, out)
        for i, r in enumerate(recs):
            with np.errstate(all="ignore"):
                ref = O.frequencies(O.cgr_count(r, k, bits))
            if out == "freq64":
                assert np.array_equal(got[i], ref, equal_nan=True), (i, bits)
            else:
                np.testing.assert_allclose(got[i], ref, rtol=1e-6, atol=0)


def test_wrap_and_wrapped_frequencies(R):
    recs = [b"A" * 70000, b"T" * 65536, b"AC" * 40000]
    for k in (1, 2, 9):
        got = _gpu_counts(R, recs, k, 16)
        f = _gpu_counts(R, recs, k, 16, "freq64")
        for i, r in enumerate(recs):
            c = O.cgr_count(r, k, 16)
            assert np.array_equal(got[i], c)
            with np.errstate(all="ignore"):
                assert np.array_equal(f[i], O.frequencies(c), equal_nan=True)


def test_property_full_size_k9(R):
    """BASELINE config 2 size (10k x 9 kb, k=9) through size-independent properties: row sums equal
    the number of emitting bases, and 64 sampled rows equal the oracle."""
    n, L, k = 10000, 9000, 9
    recs = synth_records(n, L, seed=20261002, p=HIV_P)
    b = R.pack_records(recs)
    out = R.kmer_vectors(b, k, 16)
    sums = out.to(torch.int32).sum(dim=1).cpu().numpy()
    assert (sums == L - k + 1).all()
    rng = np.random.Generator(np.random.PCG64(1))
    for i in rng.choice(n, 64, replace=False):
        assert np.array_equal(out[i].cpu().numpy(), O.cgr_count(recs[i], k, 16))


def test_large_k_demo_genomes(R, refcode_cases):
    """k = 11, 12 (512 shared-memory tiles per vector) on the demo genomes vs the reference-code goldens."""
    cases = [c for c in refcode_cases if c["name"].startswith("demo-")]
    recs = [c["record"].encode("latin-1") for c in cases]
    for k in (11, 12):
        got = _gpu_counts(R, recs, k, 16)
        for i, c in enumerate(cases):
            assert hashlib.sha1(got[i].astype("<u2").tobytes()).hexdigest() == c["k"][str(k)]["sha1"], (c["name"], k)


@pytest.mark.parametrize("bits", [16, 32])
def test_cgr_pool_pyramid_and_normalize(R, demo_records, bits):
    """kam_cgr_pool: counts(k) == pool(counts(k+1)) + fix-up, for clean, dirty (Q1) and wrapping (Q3)
    records; kam_normalize == numpy true division."""
    recs = [r for _, r in demo_records] + synth_records(12, 3000, seed=77, dirty=0.02, jitter=500) + \
        [b"NNACGTACGT", b"N" * 40 + b"ACGTTGCA", b"ACG", b"", b"A" * 70000, b"ACGT" * 20000]
    b = R.pack_records(recs)
    hi = R.kmer_vectors(b, 9, bits, "counts")
    for k in range(8, 0, -1):
        lo = R.cgr_pool(b, hi, k)
        assert torch.equal(lo.view(torch.int16 if bits == 16 else torch.int32),
                           R.kmer_vectors(b, k, bits, "counts").view(torch.int16 if bits == 16 else torch.int32)), k
        hi = lo
    c = R.kmer_vectors(b, 5, bits, "counts")
    f64 = R.normalize(c, "freq64").cpu().numpy()
    cn = c.cpu().numpy()
    with np.errstate(all="ignore"):
        for i in range(len(recs)):
            assert np.array_equal(f64[i], cn[i] / cn[i].sum(), equal_nan=True)
    f32 = R.normalize(c, "freq32").cpu().numpy()
    ok = np.isfinite(f64)
    np.testing.assert_allclose(f32[ok], f64[ok], rtol=1e-6)


@pytest.mark.parametrize("out,bits", [("counts", 16), ("counts", 32), ("freq32", 16), ("freq64", 16)])
@pytest.mark.parametrize("k_hi,top", [(7, 7), (9, 7), (9, 9), (10, 10), (8, 8), (6, 6)])
def test_fused_sweep_vs_oracle(R, demo_records, out, bits, k_hi, top):
    """kam_kmer_sweep (one scan at k_hi, every lower order by in-kernel 2x2 pooling) == the oracle at
    every order, for clean and dirty records (Q1 partial k-mers at each level), records shorter than
    k, and records the fused kernel declines (> 65535 bases, uint8 overflow) and redoes per order."""
    recs = [r for _, r in demo_records] + synth_records(20, 9000, seed=400 + k_hi, dirty=0.003, p=HIV_P, jitter=2000) + \
        [b"NNACGTACGTAC", b"N" * 40 + b"ACGTTGCAAC", b"ACG", b"", b"ACGTAC", b"A" * 70000, b"ACGT" * 300,
         b"T" * 3000, synth_records(1, 90000, seed=9)[0]]
    b = R.pack_records(recs)
    from kameris_b200 import _lib
    _lib.load().kam_kmer_sweep_set_fused_top(top)      # 7 = default hybrid; 8..10 also fuse the tiled orders
    try:
        res = R.kmer_sweep(b, 1, k_hi, bits, out)
    finally:
        _lib.load().kam_kmer_sweep_set_fused_top(7)
    for k in range(1, k_hi + 1):
        got = res[k].cpu().numpy()
        for i, r in enumerate(recs):
            c = O.cgr_count(r, k, bits)
            if out == "counts":
                assert np.array_equal(got[i], c), (k, i)
            else:
                with np.errstate(all="ignore"):
                    ref = O.frequencies(c)
                if out == "freq64":
                    assert np.array_equal(got[i], ref, equal_nan=True), (k, i)
                else:
                    ok = np.isfinite(ref)
                    np.testing.assert_allclose(got[i][ok], ref[ok], rtol=1e-6, atol=0, err_msg=str((k, i)))


def test_errors(R):
    b = R.pack_records([b"ACGT"])
    with pytest.raises(ValueError):
        R.kmer_vectors(b, 33)
    with pytest.raises(ValueError):
        R.kmer_vectors(b, 4, bits=8)
