"""one launch each of the kernels changed or added at the end of round 1, for `ncu --set full`
(profiles/r1c_kernels.*): cgr_tile_kernel with the whole-word scan (float32 and uint16 outputs, bench
workload), pack_fused_kernel, the frequency-row Gram pre-pass, the float64 Gram, the column-statistics
kernels of the classify pre-processing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kameris_b200 import repr as R, dist as D, preprocess as P, _lib
from kameris_b200.dist_bench import synth_counts
import bench
dev = torch.device("cuda:0")
lib = _lib.load()
# bench workload: 10 000 x 9 kb, k=9 (pack_fused_kernel, cgr_tile_kernel<uint8,float>, <uint8,uint16>)
chars, start = bench.synth_chars(bench.N_SEQS, bench.SEQ_LEN, bench.SEED)
d_chars = torch.zeros(len(chars) + 16, dtype=torch.uint8, device=dev)
d_chars[:len(chars)] = torch.from_numpy(chars.copy()).to(dev)
b = R.pack_chars(d_chars, torch.from_numpy(start.astype(np.int64)).to(dev), bench.N_SEQS, len(chars))
o = R.kmer_vectors(b, 9, 16, "freq32"); torch.cuda.synchronize(); del o
o = R.kmer_vectors(b, 9, 16, "counts"); torch.cuda.synchronize(); del o, b, d_chars
torch.cuda.empty_cache()
# Gram on frequency rows: gram_prepass_float_kernel<double,uint8> (+ gram_pair_kernel<i8>)
xf = synth_counts(torch, dev, 2048, 9000, 9, 20261003, kind="freq64")
g = D.gram_tri(xf, ("cosine",)); torch.cuda.synchronize(); del g, xf
# float64 Gram on arbitrary rows: gram_f64_kernel<double>
gen = torch.Generator(device=dev); gen.manual_seed(7)
xr = torch.rand((4096, 4096), dtype=torch.float64, device=dev, generator=gen)
g = D.gram_tri(xr, ("cosine",)); torch.cuda.synchronize(); del g, xr
# classify pre-processing: col_sum / col_dev / col_scale kernels on 10 000 x 4096 uint16
x = synth_counts(torch, dev, 10000, 9000, 6, 20261007)
mean, var, cnt = P.column_stats(x)
t = P.scale_columns(x, P.scale_from_variance(var, mean, cnt))
torch.cuda.synchronize()
print("ok")
