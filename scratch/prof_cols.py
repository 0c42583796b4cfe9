"""one launch of the column kernels of the classify pre-processing on a matrix larger than L2, for ncu"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import preprocess as P
from kameris_b200.dist_bench import synth_counts
dev = torch.device("cuda:0")
x = synth_counts(torch, dev, 10000, 9000, 6, 20261007).repeat(10, 1)
mean, var, cnt = P.column_stats(x)
t = P.scale_columns(x, P.scale_from_variance(var, mean, cnt))
torch.cuda.synchronize()
print("ok")
