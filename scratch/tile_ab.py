"""A/B of the two scans of cgr_tile_kernel (whole plane word vs groups of 8) on the bench shape."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import repr as R, _lib
from kameris_b200.repr_bench import _time
dev = torch.device("cuda", 0)
n, L = 10000, 9000
g = torch.Generator(device=dev); g.manual_seed(2)
p = torch.tensor([0.36, 0.18, 0.24, 0.22], device=dev)
alpha = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
chars = torch.zeros(n * L + 16, dtype=torch.uint8, device=dev)
chars[:n * L] = alpha[torch.multinomial(p, n * L, replacement=True, generator=g)]
b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
res = {}
for k in (9, 8, 7, 6):
    for kind, dt, sz in (("counts", torch.uint16, 2), ("freq32", torch.float32, 4), ("freq64", torch.float64, 8)):
        o = torch.empty((n, 4 ** k), dtype=dt, device=dev)
        ws = R.kmer_workspace(n, k, dev)
        row = {}
        for mode in (1, 0):
            _lib.load().kam_kmer_set_word_scan(mode)
            t = _time(lambda: R.kmer_vectors(b, k, 16, kind, out_tensor=o, workspace=ws), 20)
            row["word" if mode else "group8"] = {"ms": round(t * 1e3, 4), "GBps": round(n * (2250 + 4 ** k * sz) / t / 1e9, 1)}
        _lib.load().kam_kmer_set_word_scan(1)
        res["k%d_%s" % (k, kind)] = row
        del o
print(json.dumps(res))
