"""one launch sequence each of the tensor-core manhat / info paths (round 2, profiles/r2_thermo_*): the rectangle
131 072 reads x 1000 centroids (D = 4096), the `info` triangle of 16 384 genomes (D = 4096), the manhat triangle of
4096 genomes at k = 9 (D = 4^9).  Run plain, then under `ncu --metrics gpu__time_duration.sum` (launch list) and
`ncu --set full -k regex:thermo|gram_pair` (captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import dist as D
from kameris_b200.dist_bench import synth_counts
dev = torch.device("cuda:0")
reads = synth_counts(torch, dev, 131072, 150, 6, 20261004)
cent = synth_counts(torch, dev, 1000, 150, 6, 3)
o = D.manhattan_rect(reads, cent); torch.cuda.synchronize(); del o, reads, cent
x = synth_counts(torch, dev, 16384, 9000, 6, 20261009)
o = D.info_tri(x); torch.cuda.synchronize(); del o, x
x = synth_counts(torch, dev, 4096, 9000, 9, 20261012)
o = D.manhattan_tri(x); torch.cuda.synchronize(); del o, x
print("ok")
