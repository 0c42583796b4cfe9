import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import repr as R
from tests.suite import gen_chars
dev = torch.device("cuda:0")
n, L = 200000, 9000
chars = gen_chars(dev, n, L, 7)
start = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
for _ in range(3):
    b = R.pack_chars(chars, start, n, n * L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
b = R.pack_chars(chars, start, n, n * L)
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1), "GB/s", n * L / e0.elapsed_time(e1) / 1e6)
