"""one launch each of the kernels added after the first profile pass, for `ncu --set full`"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import repr as R, dist as D, _lib
from kameris_b200.repr_bench import _gen
from kameris_b200.dist_bench import synth_counts
dev = torch.device("cuda:0")
lib = _lib.load()
# short reads: cgr_warp_kernel<uint16>
n, L, k = 1_000_000, 150, 6
b = R.pack_chars(_gen(dev, n, L, 20261004), torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
o = R.kmer_vectors(b, k, 16, "counts")
torch.cuda.synchronize(); del b, o; torch.cuda.empty_cache()
# megabase sequences: cgr_long_kernel<float>
n, L, k = 148, 5_000_000, 9
b = R.pack_chars(_gen(dev, n, L, 20261005), torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
o = R.kmer_vectors(b, k, 16, "freq32")
torch.cuda.synchronize(); del b, o; torch.cuda.empty_cache()
# Gram: gram_pair_kernel<i8>, <f16>
x = synth_counts(torch, dev, 8192, 9000, 9, 20261003)
g = D.gram_tri(x, ("euclid", "cosine", "pearson")); del g
g = D.gram_tri(x, ("euclid", "cosine", "pearson"), force_f16=True); del g
torch.cuda.synchronize()
print("ok")
