"""gram_pair_kernel<i8>, N = 8192 and N = 20000, D = 4^9: ms per launch sequence for several strip heights"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import dist as D, _lib
from kameris_b200.dist_bench import synth_counts, _time
lib = _lib.load()
dev = torch.device("cuda:0")
for n in (8192, 20000):
    x = synth_counts(torch, dev, n, 9000, 9, 20261003)
    res = {}
    for strip in (8, 12, 16, 24, 32, 48):
        lib.kam_gram_set_strip(strip)
        t = _time(torch, lambda: D.gram_tri(x, ("cosine", "pearson")), 3)
        res[strip] = round(t * 1e3, 3)
    print(json.dumps({"n": n, "ms_by_strip": res}), flush=True)
    del x
    torch.cuda.empty_cache()
