"""short reads (config 4 shape): warp-per-read kernel vs CTA-per-read kernel, resident planes, preallocated output."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import repr as R, _lib
from tests.suite import gen_chars, pack
dev = torch.device("cuda:0")
m, L = 1_000_000, 150
chars = gen_chars(dev, m, L, 20261004)
b = pack(dev, chars, m, L)
lib = _lib.load()
res = {}
for k in (6, 5, 4):
    for out, dt, nb in (("counts", torch.uint16, 2), ("freq32", torch.float32, 4)):
        o = torch.empty((m, 4 ** k), dtype=dt, device=dev)
        ws = R.kmer_workspace(m, k, dev)
        for warp in (1, 0):
            lib.kam_kmer_set_warp_path(warp)
            for _ in range(3):
                R.kmer_vectors(b, k, 16, out, out_tensor=o, workspace=ws)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                R.kmer_vectors(b, k, 16, out, out_tensor=o, workspace=ws)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            byt = m * ((L + 3) // 4 + 4 ** k * nb)
            res["k%d_%s_%s" % (k, out, "warp" if warp else "cta")] = {"ms": round(ms, 4), "vec_per_s": m / ms * 1e3, "GBps": byt / ms / 1e6, "frac": byt / ms / 1e6 / 6545}
        del o
lib.kam_kmer_set_warp_path(1)
print(json.dumps(res, indent=1))
