"""Gram metrics at D = 4096 (N = 16384), three metrics then cosine only -- for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import _lib
from kameris_b200.dist_bench import synth_counts
dev = torch.device("cuda", 0)
lib = _lib.load()
n, d = 16384, 4096
x = synth_counts(torch, dev, n, 9000, 6, 20261008)
p = n * (n - 1) // 2
tri = [torch.empty(p, dtype=torch.float32, device=dev) for _ in range(3)]
ws = torch.empty(lib.kam_gram_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
for mask in (4 | 8 | 16, 8, 4 | 8 | 16, 8):
    _lib.check(lib.kam_gram_tri(x.data_ptr(), 1, n, d, d, 0, n, mask, 1, tri[0].data_ptr(), tri[1].data_ptr(),
                                tri[2].data_ptr(), ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
print("done")
