"""cta_group::2 Gram kernel vs the cluster-multicast kernel: identical outputs, then timing at N=8192, D=4^9."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kameris_b200 import dist as D, _lib
from kameris_b200.dist_bench import synth_counts, _time
lib = _lib.load()
dev = torch.device("cuda:0")
ok = True
for (n, L, k) in [(2, 500, 3), (130, 9000, 6), (257, 1000, 7), (600, 4000, 6), (1000, 9000, 8)]:
    x = synth_counts(torch, dev, n, L, k, 100 + n)
    for f16 in (False, True):
        lib.kam_gram_set_pair(0)
        a = D.gram_tri(x, ("euclid", "cosine", "pearson"), euclid_on_freq=True, force_f16=f16)
        lib.kam_gram_set_pair(1)
        b = D.gram_tri(x, ("euclid", "cosine", "pearson"), euclid_on_freq=True, force_f16=f16)
        torch.cuda.synchronize()
        same = all(torch.equal(a[m], b[m]) for m in a)
        ok &= same
        print("n=%d k=%d f16=%s identical=%s" % (n, k, f16, same), flush=True)
n, k = 8192, 9
d = 4 ** k
x = synth_counts(torch, dev, n, 9000, k, 20261003)
pairs = n * (n - 1) // 2
tri = {m: torch.empty(pairs, dtype=torch.float32, device=dev) for m in ("euclid", "cosine", "pearson")}
ws = torch.empty(lib.kam_gram_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
def gram():
    _lib.check(lib.kam_gram_tri(x.data_ptr(), 1, n, d, x.stride(0), 0, n, 4 | 8 | 16, 1, tri["euclid"].data_ptr(),
                                tri["cosine"].data_ptr(), tri["pearson"].data_ptr(), ws.data_ptr(), ws.numel(), stream))
res = {}
for pair in (0, 1):
    lib.kam_gram_set_pair(pair)
    for f16 in (0, 1):
        lib.kam_gram_force_f16(f16)
        t = _time(torch, gram, 5)
        res["pair%d_%s" % (pair, "f16" if f16 else "i8")] = {"ms": t * 1e3, "TF": 2.0 * d * pairs / t / 1e12}
        if pair == 0:
            keep = {m: tri[m].clone() for m in tri} if f16 == 0 else keep
    lib.kam_gram_force_f16(0)
ok &= all(torch.equal(keep[m], tri[m]) for m in tri) if False else ok
print(json.dumps(res, indent=1)); print("all identical:", ok)
