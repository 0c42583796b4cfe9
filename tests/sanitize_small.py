"""Small end-to-end exercise of every kernel family (for compute-sanitizer memcheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from kameris_b200 import repr as R, dist as D, _lib
from tests.helpers import synth_records, HIV_P

recs = synth_records(20, 1500, seed=1, p=HIV_P, dirty=0.01, jitter=300) + [b"", b"ACG", b"A" * 70000, synth_records(1, 200000, seed=2)[0]]
b = R.pack_records(recs)
for k, bits, out in [(1, 16, "counts"), (3, 32, "counts"), (6, 16, "freq64"), (7, 16, "counts"), (9, 16, "freq32"), (9, 32, "counts"), (10, 16, "counts")]:
    got = R.kmer_vectors(b, k, bits, out).cpu().numpy()
    for i, r in enumerate(recs):
        c = O.cgr_count(r, k, bits)
        with np.errstate(all="ignore"):
            ref = c if out == "counts" else O.frequencies(c).astype(got.dtype)
        assert np.allclose(got[i], ref, rtol=1e-6, atol=0, equal_nan=True), (k, bits, out, i)
x = np.stack([O.cgr_count(r, 5) for r in recs[:20]])
xd = torch.from_numpy(x.view(np.int16)).cuda().view(torch.uint16)
m = D.manhattan_tri(xd).cpu().numpy().view(np.uint32); i = D.info_tri(xd).cpu().numpy()
rm, ri = O.dists_triangle(x, True, True)
assert np.array_equal(m, rm) and np.array_equal(i, ri)
wide = (x.astype(np.uint32) * 700).astype(np.uint16)
assert np.array_equal(D.manhattan_tri(torch.from_numpy(wide.view(np.int16)).cuda().view(torch.uint16)).cpu().numpy().view(np.uint32), O.dists_triangle(wide, True, False)[0])
lib = _lib.load()
for cl in (1, 2):
    lib.kam_gram_set_cluster(cl)
    for f16 in (False, True):
        g = D.gram_tri(xd, force_f16=f16)
        assert np.allclose(g["cosine"].cpu().numpy(), O.cdist_triangle(x, "cosine"), rtol=1e-5)
g = D.gram_tri(torch.from_numpy(wide.view(np.int16)).cuda().view(torch.uint16))      # integer fallback
assert np.allclose(g["pearson"].cpu().numpy(), O.cdist_triangle(wide, "pearson"), rtol=1e-5)
cent = x[:5].astype(np.float32) * 0.5
r = D.manhattan_rect(xd, torch.from_numpy(cent).cuda()).cpu().numpy()
assert np.allclose(r, O.cdist_rect(x, cent, "cityblock"), rtol=1e-5)
out = np.zeros((len(recs), 4 ** 4), dtype=np.float64)
chars = np.frombuffer(b"".join(recs), dtype=np.uint8); start = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
_lib.check(lib.kam_kmers_host(chars.ctypes.data, start.ctypes.data, len(recs), 4, 16, 2, out.ctypes.data))
with np.errstate(all="ignore"):
    assert all(np.array_equal(out[i], O.frequencies(O.cgr_count(r, 4)), equal_nan=True) for i, r in enumerate(recs))
torch.cuda.synchronize()
print("sanitize_small: all checks passed")
