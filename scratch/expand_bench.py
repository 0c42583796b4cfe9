"""host-side list expansion rate (kam_sparse_expand), no GPU: n vectors of k=9 with ~8992 non-zero bins"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from kameris_b200 import _lib
lib = _lib.load()
n, k, nnz1 = 1500, 9, 8800
rng = np.random.default_rng(1)
ent = np.empty(n * 9000, dtype=np.uint32)
start = np.arange(n, dtype=np.uint64) * 9000
nnz = np.full(n, nnz1, dtype=np.uint32)
for i in range(n):
    b = np.sort(rng.choice(4 ** k, nnz1, replace=False)).astype(np.uint32)
    ent[i * 9000:i * 9000 + nnz1] = (b << 8) | 1
for kind, bits, dt in ((0, 16, np.uint16), (1, 16, np.float32), (2, 16, np.float64)):
    out = np.empty((n, 4 ** k), dtype=dt)
    for rep in range(3):
        t0 = time.perf_counter()
        rc = lib.kam_sparse_expand(ent.ctypes.data, start.ctypes.data, nnz.ctypes.data, n, k, bits, kind, out.ctypes.data)
        dt_s = time.perf_counter() - t0
    assert rc == 0
    print(dt.__name__, "threads", lib.kam_host_threads(), "%.1f GB/s  %.0f vectors/s" % (out.nbytes / dt_s / 1e9, n / dt_s))
