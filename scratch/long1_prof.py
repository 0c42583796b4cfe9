"""scratch: one launch set of the one-pass path L kernel for ncu (148 x 5 Mb, k = 9, float32)."""
import sys, torch
sys.path.insert(0, ".")
from kameris_b200 import repr as R
from kameris_b200.repr_bench import _gen
dev = torch.device("cuda:0")
sms = torch.cuda.get_device_properties(dev).multi_processor_count
n, L, k = sms, 5_000_000, 9
chars = _gen(dev, n, L, 20261005)
b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
o = torch.empty((n, 4 ** k), dtype=torch.float32, device=dev)
ws = R.kmer_workspace(n, k, dev)
for _ in range(3):
    R.kmer_vectors(b, k, 16, "freq32", out_tensor=o, workspace=ws)
torch.cuda.synchronize()
