"""GPU parity of the distance kernels (through the C ABI) vs the oracle: `manhat` and `info`
bit-exact (uint32 / float32 as the reference computes them), euclid/cosine/pearson within 1e-5
relative of scipy float64 (north_star tolerance)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import HIV_P, synth_records

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200 import dist as D
    return D


def _dev(a):
    """numpy (any of the six element types) -> cuda tensor with the same bits."""
    view = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}
    tdt = {np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32, np.dtype(np.uint64): torch.uint64}
    if a.dtype in view:
        return torch.from_numpy(a.view(view[a.dtype])).cuda().view(tdt[a.dtype])
    return torch.from_numpy(a).cuda()


def _counts(n, L, k, seed, bits=16):
    recs = synth_records(n, L, seed=seed, p=HIV_P, jitter=L // 10)
    return np.stack([O.cgr_count(r, k, bits) for r in recs])


def test_golden_reference_machine_code(D, refcode_dists_cases):
    """CUDA manhat / info == outputs of the reference's own row-lambda machine code (all six element
    types; wraps, all-zero rows, float truncation quirk)."""
    for name, x, man, inf in refcode_dists_cases:
        xd = _dev(x)
        got_m = D.manhattan_tri(xd).cpu().numpy().view(np.uint32)
        got_i = D.info_tri(xd).cpu().numpy()
        assert np.array_equal(got_m, man), name
        assert np.array_equal(got_i.view(np.uint32), inf.view(np.uint32)), name


def test_demo_k6_appendix_b(D, demo_records):
    x = np.stack([O.cgr_count(r, 6) for _, r in demo_records])
    man = D.manhattan_tri(_dev(x)).cpu().numpy().view(np.uint32)
    inf = D.info_tri(_dev(x)).cpu().numpy()
    assert man.tolist() == [3904, 4791, 4369, 4773, 4455, 4628]
    ref_m, ref_i = O.dists_triangle(x, True, True)
    assert np.array_equal(man, ref_m) and np.array_equal(inf, ref_i)


@pytest.mark.parametrize("n,L,k", [(2, 500, 3), (37, 3000, 5), (130, 9000, 6), (300, 2000, 4), (129, 1000, 7)])
def test_manhat_info_vs_oracle_bytes_path(D, n, L, k):
    x = _counts(n, L, k, seed=n + k)                 # counts <= 255: VABSDIFF4 byte path
    ref_m, ref_i = O.dists_triangle(x, True, True)
    assert np.array_equal(D.manhattan_tri(_dev(x)).cpu().numpy().view(np.uint32), ref_m)
    assert np.array_equal(D.info_tri(_dev(x)).cpu().numpy(), ref_i)


@pytest.mark.parametrize("dtype", [np.uint16, np.uint32, np.uint8])
def test_manhat_wide_counts(D, dtype):
    rng = np.random.Generator(np.random.PCG64(5))
    hi = {np.uint8: 256, np.uint16: 65536, np.uint32: 2 ** 32}[dtype]
    x = rng.integers(0, hi, size=(70, 1000), dtype=np.uint64).astype(dtype)
    x[3] = 0
    ref_m, ref_i = O.dists_triangle(x, True, True)
    assert np.array_equal(D.manhattan_tri(_dev(x)).cpu().numpy().view(np.uint32), ref_m)   # wraps mod 2^32
    assert np.array_equal(D.info_tri(_dev(x)).cpu().numpy(), ref_i)


@pytest.mark.parametrize("dtype", [np.uint64, np.float32, np.float64])
def test_manhat_other_element_types_quirk_q5(D, dtype):
    rng = np.random.Generator(np.random.PCG64(6))
    if dtype == np.uint64:
        x = rng.integers(0, 2 ** 40, size=(9, 77), dtype=np.uint64)
    else:
        x = (rng.random((9, 77)) * 3.0).astype(dtype)      # |a-b| mostly < 1: truncating accumulate (Q5)
        x[0] = 0
    ref_m, ref_i = O.dists_triangle(x, True, True)
    assert np.array_equal(D.manhattan_tri(_dev(x)).cpu().numpy().view(np.uint32), ref_m)
    assert np.array_equal(D.info_tri(_dev(x)).cpu().numpy(), ref_i)
    f = np.stack([O.frequencies(r) for r in _counts(5, 800, 3, seed=1)])
    assert not D.manhattan_tri(_dev(f)).cpu().numpy().any()            # frequencies -> all zeros


def test_row_ranges_compose(D):
    x = _counts(200, 1500, 5, seed=9)
    xd = _dev(x)
    full = D.manhattan_tri(xd).cpu().numpy()
    out = torch.zeros(D.tri_len(200), dtype=torch.int32, device="cuda")
    ranges = D.balanced_row_ranges(200, 3)
    assert ranges[0][0] == 0 and ranges[-1][1] == 200
    for r in ranges:
        D.manhattan_tri(xd, rows=r, out=out)
    assert np.array_equal(out.cpu().numpy(), full)


def test_manhat_rect(D):
    x = _counts(300, 150, 6, seed=20261004)
    y = _counts(20, 150, 6, seed=3)
    got = D.manhattan_rect(_dev(x), _dev(y)).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, O.manhat_rect(x, y))
    cent = np.stack([x[i::20].mean(axis=0) for i in range(20)]).astype(np.float32)       # centroids
    gotf = D.manhattan_rect(_dev(x), torch.from_numpy(cent).cuda()).cpu().numpy()
    ref = O.cdist_rect(x, cent, "cityblock")
    np.testing.assert_allclose(gotf, ref, rtol=1e-5)


def test_symmetry_and_full_size_property(D):
    """N=1500, D=4^6: triangle equals the mirrored rectangular result; zero self-distance."""
    x = _counts(1500, 400, 6, seed=77)
    xd = _dev(x)
    tri = D.manhattan_tri(xd).cpu().numpy().view(np.uint32)
    rect = D.manhattan_rect(xd, xd).cpu().numpy().view(np.uint32)
    assert (np.diag(rect) == 0).all() and np.array_equal(rect, rect.T)
    iu = np.triu_indices(1500, 1)
    assert np.array_equal(rect[iu], tri)
    rng = np.random.Generator(np.random.PCG64(2))
    for _ in range(200):
        i, j = sorted(rng.choice(1500, 2, replace=False))
        assert tri[O.triangle_index(1500, i, j)] == O.manhat(x[i], x[j])


# ---- Gram metrics (euclid / cosine / pearson): tcgen05 tensor-core path -----------------------------
def _check_gram(D, x, rtol=1e-5, force_f16=False):
    xd = _dev(x)
    got = D.gram_tri(xd, ("euclid", "cosine", "pearson"), euclid_on_freq=True, force_f16=force_f16)
    f = x / x.sum(axis=1, keepdims=True)
    np.testing.assert_allclose(got["euclid"].cpu().numpy(), O.cdist_triangle(f, "euclid"), rtol=rtol)
    np.testing.assert_allclose(got["cosine"].cpu().numpy(), O.cdist_triangle(x, "cosine"), rtol=rtol)
    np.testing.assert_allclose(got["pearson"].cpu().numpy(), O.cdist_triangle(x, "pearson"), rtol=rtol)
    got_c = D.gram_tri(xd, ("euclid",), euclid_on_freq=False)
    np.testing.assert_allclose(got_c["euclid"].cpu().numpy(), O.cdist_triangle(x, "euclid"), rtol=rtol)


def test_gram_demo_k6_appendix_b(D, demo_records):
    x = np.stack([O.cgr_count(r, 6) for _, r in demo_records])
    got = D.gram_tri(_dev(x), ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    np.testing.assert_allclose(got["euclid"].cpu().numpy(),
                               [0.0107171533, 0.0122965513, 0.0118525572, 0.0119359576, 0.0116858869,
                                0.0116338804], rtol=1e-5)
    np.testing.assert_allclose(got["cosine"].cpu().numpy(),
                               [0.0938612871, 0.1247486, 0.114731872, 0.121487558, 0.114964415, 0.115188912],
                               rtol=1e-5)
    np.testing.assert_allclose(got["pearson"].cpu().numpy(),
                               [0.156244013, 0.209079337, 0.190899155, 0.208178386, 0.195207544, 0.197113458],
                               rtol=1e-5)


@pytest.mark.parametrize("n,L,k", [(2, 500, 3), (37, 3000, 5), (130, 9000, 6), (300, 2000, 4), (257, 1000, 7),
                                   (600, 4000, 6), (40, 9000, 9)])
def test_gram_vs_scipy(D, n, L, k):
    x = _counts(n, L, k, seed=100 + n + k)
    _check_gram(D, x)                      # uint8 x uint8 -> int32 tensor path (counts <= 255)
    _check_gram(D, x, force_f16=True)      # fp16 -> fp32 tensor path


@pytest.mark.parametrize("n,L,k", [(3, 500, 3), (257, 1000, 7), (700, 4000, 6), (520, 9000, 8)])
def test_gram_kernel_variants_agree_exactly(D, n, L, k):
    """cta_group::2 pair kernel (default), cluster-multicast cta_group::1 kernel and the no-cluster kernel
    all produce the exact integer Gram matrix: identical float32 outputs, both operand kinds."""
    from kameris_b200 import _lib
    lib = _lib.load()
    x = _dev(_counts(n, L, k, seed=n + k))
    res = []
    try:
        for pair, cl in ((1, 2), (0, 2), (0, 1)):
            lib.kam_gram_set_pair(pair)
            lib.kam_gram_set_cluster(cl)
            for f16 in (False, True):
                res.append(D.gram_tri(x, ("euclid", "cosine", "pearson"), euclid_on_freq=True, force_f16=f16))
    finally:
        lib.kam_gram_set_pair(1)
        lib.kam_gram_set_cluster(2)
    for r in res[1:]:
        for m in r:
            assert torch.equal(r[m], res[0][m])


def test_gram_i8_and_f16_paths_agree_exactly(D):
    """both tensor paths produce the exact integer Gram matrix, hence identical float32 outputs."""
    x = _dev(_counts(300, 5000, 6, seed=5))
    a = D.gram_tri(x)
    b = D.gram_tri(x, force_f16=True)
    for m in a:
        assert torch.equal(a[m], b[m])


def test_gram_f16_path_mid_counts(D):
    rng = np.random.Generator(np.random.PCG64(12))
    x = rng.integers(0, 1500, size=(200, 12)).astype(np.uint16)       # 255 < max <= 2048, G_ii < 2^24
    _check_gram(D, x)


def test_gram_ragged_d_and_u32(D):
    rng = np.random.Generator(np.random.PCG64(8))
    x = rng.integers(0, 50, size=(150, 1000)).astype(np.uint32)       # d not a multiple of 64
    _check_gram(D, x)
    _check_gram(D, x[:, :4].copy() + 1)                                 # k=1-like: d=4


@pytest.mark.parametrize("limb", [1, 0])
def test_gram_outside_fp16_envelope(D, limb):
    """counts > 2048 or sum of squares >= 2^24: the limb-split tensor path (c = 256 hi + lo, three uint8 Grams
    combined exactly; limb=1) or the exact integer Gram on the CUDA cores (limb=0) -- same results."""
    from kameris_b200 import _lib
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(9))
    lib.kam_gram_set_limb(limb)
    try:
        x = rng.integers(0, 60000, size=(70, 300)).astype(np.uint16)
        _check_gram(D, x)
        x2 = rng.integers(100, 2000, size=(50, 4096)).astype(np.uint16)     # <= 2048 but G_ii ~ 5e9 >= 2^24
        _check_gram(D, x2)
        x3 = rng.integers(0, 70000, size=(33, 500)).astype(np.uint32)       # above 65535: never the limb path
        _check_gram(D, x3)
    finally:
        lib.kam_gram_set_limb(1)


def test_gram_limb_path_many_tiles_and_its_limit(D):
    """limb-split path on a config-5-like shape (mean count ~19 with a heavy tail above 255, D = 4^9: the sum of
    squares of a row is ~1e8, far beyond the fp16 envelope) over several tiles and row blocks, against a float64
    Gram; identical bits to the CUDA-core integer Gram; and rows whose low bytes alone would overflow an int32
    accumulator (all 255 over 65536 bins) must fall through to the integer path and still be right."""
    from kameris_b200 import _lib
    lib = _lib.load()
    n, d = 900, 4 ** 9
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    x = torch.poisson(torch.full((n, d), 19.0, device="cuda"), generator=g)
    x[:, ::977] *= 40                                     # ~270 bins per row with counts ~760: hi limbs in use
    assert float(x.max()) > 255 and float(x.max()) < 65536
    xd = x.to(torch.int32).to(torch.uint16)
    ref = _ref_metrics_f64(x.to(torch.float64))
    iu = torch.triu_indices(n, n, 1, device="cuda")
    got = D.gram_tri(xd, ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    for m in got:
        want = ref[m][iu[0], iu[1]]
        err = ((got[m].double() - want).abs() / want.abs().clamp(min=1e-30)).max().item()
        assert err < 2e-6, (m, err)
    part = D.gram_tri(xd, ("cosine",), rows=(300, 601))["cosine"]           # a row range that starts inside a tile
    lo, hi = 300 * n - 300 * 301 // 2, 601 * n - 601 * 602 // 2
    assert torch.equal(part[lo:hi], got["cosine"][lo:hi])
    lib.kam_gram_set_limb(0)
    try:
        slow = D.gram_tri(xd[:200], ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    finally:
        lib.kam_gram_set_limb(1)
    fast = D.gram_tri(xd[:200], ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    for m in slow:
        assert torch.equal(slow[m], fast[m]), m
    y = np.full((40, 65536), 255, dtype=np.uint16)         # sum lo^2 = 4.26e9 >= 2^31: limb path must decline
    y[:, ::3] = 300
    y[np.arange(40), np.arange(40)] = 7
    _check_gram(D, y)


def test_gram_row_ranges_compose(D):
    x = _counts(500, 1500, 5, seed=19)
    xd = _dev(x)
    full = D.gram_tri(xd, ("cosine",))["cosine"].cpu().numpy()
    acc = np.zeros_like(full)
    for r in D.balanced_row_ranges(500, 3):
        part = D.gram_tri(xd, ("cosine",), rows=r)["cosine"].cpu().numpy()
        assert not (np.logical_and(acc != 0, part != 0)).any()
        acc += part
    assert np.array_equal(acc, full)


def test_gram_ring_single_rank_equals_triangle(D):
    """parallel.gram_ring without a process group: one triangular block == the packed triangle."""
    from kameris_b200 import parallel
    x = _counts(333, 2000, 6, seed=21)
    xd = _dev(x)
    blocks, counts = parallel.gram_ring(xd, ("cosine", "euclid"))
    assert counts == [333] and len(blocks) == 1 and blocks[0]["tri"]
    ref = D.gram_tri(xd, ("cosine", "euclid"))
    iu = np.triu_indices(333, 1)
    for m in ("cosine", "euclid"):
        assert np.array_equal(blocks[0][m].cpu().numpy()[iu], ref[m].cpu().numpy())


def test_gram_block_rectangular(D):
    """kam_gram_block on two different row sets (dense output) vs scipy."""
    from kameris_b200 import _lib
    lib = _lib.load()
    xa, xb = _counts(150, 2000, 5, seed=31), _counts(70, 2000, 5, seed=32)
    d = xa.shape[1]
    dpad = int(lib.kam_gram_dpad(d))
    st = torch.cuda.current_stream().cuda_stream
    ops, S, G = [], [], []
    for x in (xa, xb):
        xd = _dev(x)
        op = torch.empty((x.shape[0], dpad), dtype=torch.uint8, device="cuda")
        s = torch.zeros(x.shape[0], dtype=torch.int64, device="cuda")
        g = torch.zeros(x.shape[0], dtype=torch.int64, device="cuda")
        fl = torch.zeros(8, dtype=torch.int64, device="cuda")
        _lib.check(lib.kam_gram_prepare(xd.data_ptr(), 1, x.shape[0], d, d, 0, op.data_ptr(), s.data_ptr(), g.data_ptr(),
                                        fl.data_ptr(), st))
        assert int(fl[0]) == int(x.max()) and int(fl[1]) == int((x.astype(np.int64) ** 2).sum(axis=1).max())
        assert np.array_equal(s.cpu().numpy(), x.sum(axis=1))
        ops.append(op), S.append(s), G.append(g)
    out = torch.zeros((150, 70), dtype=torch.float32, device="cuda")
    ws = torch.empty(lib.kam_gram_block_workspace_bytes(150, 70), dtype=torch.uint8, device="cuda")
    _lib.check(lib.kam_gram_block(ops[0].data_ptr(), S[0].data_ptr(), G[0].data_ptr(), 150, ops[1].data_ptr(),
                                  S[1].data_ptr(), G[1].data_ptr(), 70, 0, d, 0, 0, 150, _lib.KAM_M_PEARSON, 0, None, None,
                                  out.data_ptr(), 70, 0, ws.data_ptr(), ws.numel(), st))
    np.testing.assert_allclose(out.cpu().numpy(), O.cdist_rect(xa, xb, "pearson"), rtol=1e-5)


# ---- Gram metrics on floating-point rows (a `frequencies` .mm-repr) ----------------------------------
def _freq(x, dtype):
    """cgr / cgr.sum() as backend.py:49 computes it (float64); float32 as kmer_count.cu's freq32 kind."""
    with np.errstate(all="ignore"):
        f = x / x.sum(axis=1, keepdims=True).astype(np.float64)
    return f.astype(dtype)


def _set_recover(on):
    from kameris_b200 import _lib
    _lib.load().kam_gram_set_float_recover(on)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,L,k", [(37, 3000, 5), (130, 9000, 6), (257, 1000, 7), (40, 9000, 9), (65, 2000, 4)])
def test_gram_on_frequency_rows_recovers_counts(D, n, L, k, dtype):
    """float rows f = c / sum(c): the pre-pass recovers the integer counts (verified bit for bit), so the
    tensor-core path gives exactly the triangles of the count matrix with euclid on frequencies; the
    float64 CUDA-core kernel (recovery off) agrees to rounding; both match scipy on the stored rows."""
    x = _counts(n, L, k, seed=300 + n + k)
    f = _freq(x, dtype)
    got = D.gram_tri(_dev(f), ("euclid", "cosine", "pearson"))
    ref = D.gram_tri(_dev(x), ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    recovered = all(int(r[r > 0].min()) <= 8 for r in x)
    if recovered:
        for m in got:
            assert torch.equal(got[m], ref[m]), m
    _set_recover(0)
    try:
        alt = D.gram_tri(_dev(f), ("euclid", "cosine", "pearson"))
    finally:
        _set_recover(1)
    f64 = f.astype(np.float64)
    for m in ("euclid", "cosine", "pearson"):
        want = O.cdist_triangle(f64, m)
        tol = 1e-5 if dtype == np.float64 else 2e-4      # float32 rows: the INPUT is only good to 6e-8 per element
        np.testing.assert_allclose(alt[m].cpu().numpy(), want, rtol=tol, err_msg=m)
        np.testing.assert_allclose(got[m].cpu().numpy(), want, rtol=tol, err_msg=m)


def test_gram_on_float_rows_general(D):
    """rows that are not frequency vectors (arbitrary positive / signed floats, ragged d), an all-NaN row
    (the 0/0 of an empty sequence) and an all-zero row: the float64 Gram between the rows as stored."""
    rng = np.random.Generator(np.random.PCG64(77))
    for n, d in ((5, 7), (70, 100), (130, 1000), (64, 64)):
        x = rng.standard_normal((n, d))
        got = D.gram_tri(_dev(x), ("euclid", "cosine", "pearson"))
        for m in got:
            np.testing.assert_allclose(got[m].cpu().numpy(), O.cdist_triangle(x, m), rtol=1e-5, atol=1e-6, err_msg=m)
        x32 = np.abs(x).astype(np.float32)
        got = D.gram_tri(_dev(x32), ("euclid", "cosine", "pearson"))
        for m in got:
            np.testing.assert_allclose(got[m].cpu().numpy(), O.cdist_triangle(x32.astype(np.float64), m), rtol=1e-5,
                                       atol=1e-6, err_msg=m)
    # frequency rows + one row of NaN: still recovered; pairs with the NaN row are NaN like scipy's
    c = _counts(20, 3000, 7, seed=9)
    f = _freq(c, np.float64)
    f[3] = np.nan
    got = D.gram_tri(_dev(f), ("euclid", "cosine", "pearson"))
    with np.errstate(all="ignore"):
        for m in got:
            want = O.cdist_triangle(f, m)
            g = got[m].cpu().numpy()
            assert np.array_equal(np.isnan(g), np.isnan(want)), m
            ok = ~np.isnan(want)
            np.testing.assert_allclose(g[ok], want[ok], rtol=1e-5, err_msg=m)
    # a mixed row (some NaN) or a zero row is not a frequency vector: float64 path, same NaN pattern as scipy
    f = _freq(c, np.float64)
    f[5, :10] = np.nan
    f[7] = 0.0
    got = D.gram_tri(_dev(f), ("euclid", "cosine"))
    with np.errstate(all="ignore"):
        for m in got:
            want = O.cdist_triangle(f, m)
            g = got[m].cpu().numpy()
            assert np.array_equal(np.isnan(g), np.isnan(want)), m
            ok = ~np.isnan(want)
            np.testing.assert_allclose(g[ok], want[ok], rtol=1e-5, err_msg=m)


def test_gram_on_frequency_rows_row_ranges_and_f16(D):
    """row blocks compose on both float paths; counts above 255 take the fp16 operand after recovery."""
    c = _counts(150, 9000, 6, seed=41)
    f = _freq(c, np.float64)
    whole = D.gram_tri(_dev(f), ("cosine", "euclid"))
    for rec in (1, 0):
        _set_recover(rec)
        try:
            full = D.gram_tri(_dev(f), ("cosine", "euclid"))
            parts = [D.gram_tri(_dev(f), ("cosine", "euclid"), rows=r) for r in ((0, 40), (40, 97), (97, 150))]
        finally:
            _set_recover(1)
        for m in full:
            assert torch.equal(sum(p[m] for p in parts), full[m]), (rec, m)
            np.testing.assert_allclose(full[m].cpu().numpy(), whole[m].cpu().numpy(), rtol=1e-6)
    rng = np.random.Generator(np.random.PCG64(5))
    c = rng.integers(0, 600, size=(90, 48)).astype(np.uint16)
    c[:, 0] = 1                                           # smallest positive count 1: recoverable
    f = _freq(c, np.float64)
    got = D.gram_tri(_dev(f), ("euclid", "cosine", "pearson"))
    ref = D.gram_tri(_dev(c), ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    for m in got:
        assert torch.equal(got[m], ref[m]), m


def test_gram_ring_on_frequency_rows(D):
    """the block interface (kam_gram_prepare) recovers the counts of float64 frequency rows too: the ring's
    block equals the packed triangle of the count matrix (euclid between frequency vectors whatever the
    caller's flag); rows that are not frequency vectors are an error there (no float64 block kernel)."""
    from kameris_b200 import parallel
    x = _counts(150, 3000, 7, seed=23)
    f = _dev(_freq(x, np.float64))
    blocks, counts = parallel.gram_ring(f, ("cosine", "euclid"), euclid_on_freq=False)
    ref = D.gram_tri(_dev(x), ("cosine", "euclid"), euclid_on_freq=True)
    iu = np.triu_indices(150, 1)
    for m in ("cosine", "euclid"):
        assert np.array_equal(blocks[0][m].cpu().numpy()[iu], ref[m].cpu().numpy())
    bad = np.abs(np.random.Generator(np.random.PCG64(1)).standard_normal((20, 64)))
    with pytest.raises(ValueError):
        parallel.gram_ring(_dev(bad), ("cosine",))


# ---- Gram kernels past one tile per CTA pair: the persistent loop, TMEM double buffering, stage-ring wrap ----
def _ref_metrics_f64(xd64):
    """float64 reference of all pairs on the device: exact Gram (sums < 2^53) + the scipy formulas"""
    g = xd64 @ xd64.T
    s = xd64.sum(dim=1)
    d = xd64.shape[1]
    gii = torch.diagonal(g)
    cos = 1.0 - g / torch.sqrt(gii[:, None] * gii[None, :])
    cov = g - s[:, None] * s[None, :] / d
    var = gii - s * s / d
    pea = 1.0 - cov / torch.sqrt(var[:, None] * var[None, :])
    f2 = gii / (s * s)
    euc = torch.sqrt(torch.clamp(f2[:, None] + f2[None, :] - 2.0 * g / (s[:, None] * s[None, :]), min=0.0))
    return {"euclid": euc, "cosine": cos, "pearson": pea}


def test_gram_many_waves_every_pair_every_variant(D):
    """N = 6450, D = 4096: 26 x 26 / 2 = 351 tiles of 256 x 256 on 74 CTA pairs (4.7 waves) -- every CTA pair
    runs the persistent loop several times, so the TMEM accumulator hand-over (tfull / tempty phases), the
    wrap of the shared-memory stage ring across tiles and the strip schedule are all exercised.  EVERY pair is
    compared with a float64 Gram (exact) + the scipy formulas, for the three kernel variants and both operand
    kinds; the variants must also agree bit for bit."""
    from kameris_b200 import _lib
    lib = _lib.load()
    n, d = 6450, 4096
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.poisson(torch.full((n, d), 2.2, device="cuda"), generator=g)      # counts 0..~14
    x[17] *= 9                                                                   # a row with larger counts (<= 255)
    xd = x.to(torch.int32).to(torch.uint16)
    ref = _ref_metrics_f64(x.to(torch.float64))
    iu = torch.triu_indices(n, n, 1, device="cuda")
    want = {m: ref[m][iu[0], iu[1]] for m in ref}
    del ref
    first = None
    try:
        for pair, cl in ((1, 2), (0, 2), (0, 1)):
            lib.kam_gram_set_pair(pair)
            lib.kam_gram_set_cluster(cl)
            for f16 in (False, True):
                got = D.gram_tri(xd, ("euclid", "cosine", "pearson"), euclid_on_freq=True, force_f16=f16)
                for m in got:
                    err = ((got[m].double() - want[m]).abs() / want[m].abs().clamp(min=1e-30)).max().item()
                    assert err < 2e-6, (pair, cl, f16, m, err)
                if first is None:
                    first = got
                else:
                    for m in got:
                        assert torch.equal(got[m], first[m]), (pair, cl, f16, m)
    finally:
        lib.kam_gram_set_pair(1)
        lib.kam_gram_set_cluster(2)


def test_gram_block_halves_and_rectangles_many_waves(D):
    """kam_gram_block as the ring uses it, at sizes with several waves of tiles: a full rectangular block, its
    `lo` / `hi` row halves (they must tile the full block exactly) and a triangular block with dense output,
    every entry against the float64 reference."""
    from kameris_b200 import _lib
    lib = _lib.load()
    na, nb, d = 3300, 3100, 4096
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    xa = torch.poisson(torch.full((na, d), 1.7, device="cuda"), generator=g)
    xb = torch.poisson(torch.full((nb, d), 2.9, device="cuda"), generator=g)
    ref = _ref_metrics_f64(torch.cat([xa, xb]).to(torch.float64))
    st = torch.cuda.current_stream().cuda_stream
    dpad = int(lib.kam_gram_dpad(d))
    ops = []
    for x in (xa, xb):
        xd = x.to(torch.int32).to(torch.uint16)
        op = torch.empty((x.shape[0], dpad), dtype=torch.uint8, device="cuda")
        s = torch.zeros(x.shape[0], dtype=torch.int64, device="cuda")
        gg = torch.zeros(x.shape[0], dtype=torch.int64, device="cuda")
        fl = torch.zeros(8, dtype=torch.int64, device="cuda")
        _lib.check(lib.kam_gram_prepare(xd.data_ptr(), 1, x.shape[0], d, d, 0, op.data_ptr(), s.data_ptr(), gg.data_ptr(),
                                        fl.data_ptr(), st))
        ops.append((op, s, gg, x.shape[0]))

    def block(a, b, tri, r0, r1, metric_bit):
        (oa, sa, ga, n_a), (ob, sb, gb, n_b) = ops[a], ops[b]
        out = torch.full((n_a, n_b), -7.0, dtype=torch.float32, device="cuda")
        ws = torch.empty(lib.kam_gram_block_workspace_bytes(n_a, n_b), dtype=torch.uint8, device="cuda")
        p = out.data_ptr()
        _lib.check(lib.kam_gram_block(oa.data_ptr(), sa.data_ptr(), ga.data_ptr(), n_a, ob.data_ptr(), sb.data_ptr(),
                                      gb.data_ptr(), n_b, 0, d, int(tri), r0, r1, metric_bit, 1,
                                      p if metric_bit == 4 else None, p if metric_bit == 8 else None,
                                      p if metric_bit == 16 else None, n_b, 0, ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        return out

    def close(got, want):
        return ((got.double() - want).abs() / want.abs().clamp(min=1e-30)).max().item() < 2e-6

    for name, bit in (("cosine", 8), ("pearson", 16), ("euclid", 4)):
        want = ref[name][:na, na:]
        full = block(0, 1, False, 0, na, bit)
        assert close(full, want), name
        h = na // 2
        lo, hi = block(0, 1, False, 0, h, bit), block(0, 1, False, h, na, bit)
        assert torch.equal(lo[:h], full[:h]) and torch.equal(hi[h:], full[h:]), name
        assert bool((lo[h:] == -7.0).all()) and bool((hi[:h] == -7.0).all()), "rows outside the range were written"
    tri = block(1, 1, True, 0, nb, 8)
    iu = torch.triu_indices(nb, nb, 1, device="cuda")
    assert close(tri[iu[0], iu[1]], ref["cosine"][na:, na:][iu[0], iu[1]])
    il = torch.tril_indices(nb, nb, 0, device="cuda")
    assert bool((tri[il[0], il[1]] == -7.0).all()), "a triangular block wrote on or below its diagonal"
