"""one launch each of the kernels changed or added in round 2, for `ncu --set full` (profiles/r2_kernels.*):
the bench step (pack_fused_kernel, cgr_tile_kernel<uint8,float>), cgr_sparse_kernel (the host-return path of
kam_kmers_host), gram_pair_kernel<i8> with the staged epilogue at D = 4^9 (N = 8192) and at D = 4096
(N = 16384, three metrics), the limb-split Gram (prepass + three launches + combine, N = 2048, D = 4^9)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kameris_b200 import repr as R, dist as D, _lib
from kameris_b200.dist_bench import synth_counts
import bench
dev = torch.device("cuda:0")
lib = _lib.load()
chars, start = bench.synth_chars(bench.N_SEQS, bench.SEQ_LEN, bench.SEED)
d_chars = torch.zeros(len(chars) + 16, dtype=torch.uint8, device=dev)
d_chars[:len(chars)] = torch.from_numpy(chars.copy()).to(dev)
d_start = torch.from_numpy(start.astype(np.int64)).to(dev)
b = R.pack_chars(d_chars, d_start, bench.N_SEQS, len(chars))
o = R.kmer_vectors(b, 9, 16, "freq32"); torch.cuda.synchronize(); del o
ent, nnz = R.kmer_sparse(b, 9, d_start); torch.cuda.synchronize()
assert int(nnz.min()) > 8000 and int(nnz.max()) <= 8992
del ent, nnz, b, d_chars
torch.cuda.empty_cache()
x = synth_counts(torch, dev, 8192, 9000, 9, 20261003)
g = D.gram_tri(x, ("euclid", "cosine", "pearson")); torch.cuda.synchronize(); del g, x
x = synth_counts(torch, dev, 16384, 9000, 6, 20261008)
g = D.gram_tri(x, ("euclid", "cosine", "pearson")); torch.cuda.synchronize(); del g, x
gen = torch.Generator(device=dev); gen.manual_seed(3)
xl = torch.poisson(torch.full((2048, 4 ** 9), 19.0, device=dev), generator=gen)
xl[:, ::977] *= 40
xl = xl.to(torch.int32).to(torch.uint16)
g = D.gram_tri(xl, ("cosine", "pearson")); torch.cuda.synchronize()
print("ok")
