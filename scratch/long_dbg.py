import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import _lib, repr as R
from kameris_b200.repr_bench import _gen
dev = torch.device("cuda", 0)
lib = _lib.load()
n, L, k = 148, 5_000_000, 9
chars = _gen(dev, n, L, 20261005)
b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
o = torch.empty((n, 4 ** k), dtype=torch.float32, device=dev)
ws = R.kmer_workspace(n, k, dev)
for mode in (2, 1, 3):
    lib.kam_kmer_set_long_path(mode)
    R.kmer_vectors(b, k, 16, "freq32", out_tensor=o, workspace=ws)
    torch.cuda.synchronize()
    print(mode, "hdr words (queue, spill_count, maxlen lo/hi, queue2, spill2_count, maxlen2 lo/hi):", ws[:32].view(torch.int32).tolist())
