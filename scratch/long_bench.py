"""config 5 shape: n x 5 Mb, k=9 float32 frequencies; path L (shared-memory uint16 tiles) vs path B (RED.ADD)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import repr as R, _lib
from tests.suite import gen_chars, pack
dev = torch.device("cuda:0")
n, L, k = int(os.environ.get("N", "148")), int(os.environ.get("L", "5000000")), int(os.environ.get("K", "9"))
chars = gen_chars(dev, n, L, 20261005)
b = pack(dev, chars, n, L)
lib = _lib.load()
out = torch.empty((n, 4 ** k), dtype=torch.float32, device=dev)
ws = R.kmer_workspace(n, k, dev)
res = {}
for mode in ([1] if os.environ.get("ONLY_L") else [1, 0]):
    lib.kam_kmer_set_long_path(mode)
    R.kmer_vectors(b, k, 16, "freq32", out_tensor=out, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = int(os.environ.get("REPS", "3"))
    for _ in range(reps):
        R.kmer_vectors(b, k, 16, "freq32", out_tensor=out, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["L" if mode else "B"] = {"ms": ms, "seq_per_s": n / ms * 1e3, "kmers_per_s": n * L / ms * 1e3}
lib.kam_kmer_set_long_path(1)
print(json.dumps(res))
