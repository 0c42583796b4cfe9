"""GPU parity of the sparse return path: kam_kmer_count_sparse (sorted (bin << 8 | count) lists) against the
oracle, and kam_kmers_host with the host-side expansion against its own dense path -- identical bytes for
uint16 / uint32 counts and float32 / float64 frequencies, including chunks the sparse kernel declines."""
import threading

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


def _mixed_records(seed, n=40, length=4000):
    recs = synth_records(n, length, seed=seed, dirty=0.003, jitter=length - 50, p=HIV_P)
    recs += [b"", b"NNNNNNNN", b"ACG", b"ACGTACGTA", b"NACGTTGCAAGT", b"acgtACGTACGTTTGA" * 9]
    return recs


@pytest.mark.parametrize("k", [7, 8, 9, 10])
def test_sparse_lists_match_oracle(R, k):
    recs = _mixed_records(k)
    recs.insert(3, b"A" * 700)                       # one bin counted 700-k+1 times: above 255 -> declined
    recs.insert(9, b"ACGT" * 40000)                  # 160 000 bases: longer than the tiles take -> declined
    b = R.pack_records(recs)
    chars, start = R.concat_records(recs)
    ent, nnz = R.kmer_sparse(b, k, torch.from_numpy(start.astype(np.int64)).cuda())
    ent, nnz = ent.cpu().numpy().view(np.uint32), nnz.cpu().numpy().view(np.uint32)
    for i, r in enumerate(recs):
        want = O.cgr_count(r, k, 32)
        if i in (3, 9):
            assert nnz[i] == 0xffffffff, i
            continue
        assert nnz[i] != 0xffffffff and nnz[i] <= len(r), i
        e = ent[int(start[i]):int(start[i]) + int(nnz[i])]
        bins, cnt = e >> 8, e & 0xff
        assert np.all(np.diff(bins.astype(np.int64)) > 0), "entries must ascend by bin"
        got = np.zeros(4 ** k, dtype=np.uint32)
        got[bins] = cnt
        assert np.array_equal(got, want), i


def _host(lib, _lib, chars, start, n, k, bits, kind, np_dt, pinned):
    out = torch.empty((n, 4 ** k), dtype={np.uint16: torch.uint16, np.uint32: torch.uint32, np.float32: torch.float32,
                                          np.float64: torch.float64}[np_dt])
    if pinned:
        out = out.pin_memory()
    _lib.check(lib.kam_kmers_host(chars.ctypes.data, start.ctypes.data, n, k, bits, kind, out.data_ptr()))
    return out.numpy().copy() if np_dt not in (np.uint16, np.uint32) else \
        out.view(torch.int16 if np_dt == np.uint16 else torch.int32).numpy().view(np_dt).copy()


@pytest.mark.parametrize("k,n,length", [(7, 2600, 300), (9, 300, 5000)])
def test_kmers_host_sparse_equals_dense(R, k, n, length):
    """more sequences than one sparse chunk holds (1024 rows), ragged lengths, dirty heads, empty records;
    forced-sparse, automatic and dense runs must write the same bytes"""
    from kameris_b200 import _lib
    lib = _lib.load()
    recs = synth_records(n, length, seed=31 + k, dirty=0.003, jitter=length - 20, p=HIV_P) + [b"", b"NN", b"ACGTA"]
    chars, start = R.concat_records(recs)
    chars = np.concatenate([chars, np.zeros(16, np.uint8)])
    m = len(recs)
    ref = np.stack([O.cgr_count(r, k, 16) for r in recs[:64]])
    for bits, kind, dt in ((16, _lib.KAM_OUT_COUNTS, np.uint16), (32, _lib.KAM_OUT_COUNTS, np.uint32),
                           (16, _lib.KAM_OUT_FREQ_F32, np.float32), (16, _lib.KAM_OUT_FREQ_F64, np.float64)):
        res = {}
        try:
            for mode in (0, 2, 1):
                lib.kam_kmers_host_set_sparse(mode)
                res[mode] = _host(lib, _lib, chars, start, m, k, bits, kind, dt, pinned=(mode != 2))
        finally:
            lib.kam_kmers_host_set_sparse(1)
        if kind == _lib.KAM_OUT_COUNTS:
            assert np.array_equal(res[0][:64], ref.astype(dt))
        for mode in (2, 1):
            assert res[mode].tobytes() == res[0].tobytes() or \
                np.array_equal(res[mode], res[0], equal_nan=True), (bits, kind, mode)
        # NaN rows (0/0 of the empty records) aside, the bytes are the same
        ok = ~np.isnan(res[0]).any(axis=1) if dt in (np.float32, np.float64) else np.ones(m, bool)
        assert res[2][ok].tobytes() == res[0][ok].tobytes()


def test_kmers_host_declined_chunk_is_redone_densely(R):
    from kameris_b200 import _lib
    lib = _lib.load()
    k = 8
    recs = synth_records(50, 2000, seed=5, p=HIV_P)
    recs[17] = b"T" * 3000                            # count 2993 in one bin
    recs[40] = b"GATTACA" * 30000                     # 210 000 bases
    chars, start = R.concat_records(recs)
    chars = np.concatenate([chars, np.zeros(16, np.uint8)])
    try:
        lib.kam_kmers_host_set_sparse(2)
        got = _host(lib, _lib, chars, start, len(recs), k, 16, _lib.KAM_OUT_COUNTS, np.uint16, pinned=False)
        f64 = _host(lib, _lib, chars, start, len(recs), k, 16, _lib.KAM_OUT_FREQ_F64, np.float64, pinned=True)
    finally:
        lib.kam_kmers_host_set_sparse(1)
    for i, r in enumerate(recs):
        c = O.cgr_count(r, k, 16)
        assert np.array_equal(got[i], c), i
        assert np.array_equal(f64[i], O.frequencies(c)), i


def test_kmers_host_is_safe_to_call_from_several_threads(R):
    """the header promises serialised calls from any thread (ctypes releases the GIL)"""
    from kameris_b200 import _lib
    lib = _lib.load()
    sets = []
    for t in range(4):
        recs = synth_records(200, 1500, seed=100 + t, jitter=700, dirty=0.002, p=HIV_P)
        chars, start = R.concat_records(recs)
        sets.append((recs, np.concatenate([chars, np.zeros(16, np.uint8)]), start))
    out, err = [None] * 4, []

    def work(t):
        try:
            recs, chars, start = sets[t]
            for _ in range(3):
                out[t] = _host(lib, _lib, chars, start, len(recs), 7 + (t & 1), 16, _lib.KAM_OUT_COUNTS, np.uint16,
                               pinned=False)
        except Exception as e:   # noqa: BLE001
            err.append(e)
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not err, err
    for t in range(4):
        recs = sets[t][0]
        for i in (0, 57, 199):
            assert np.array_equal(out[t][i], O.cgr_count(recs[i], 7 + (t & 1), 16)), (t, i)
