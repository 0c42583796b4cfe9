import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import mds as M
dev = torch.device("cuda:0")
n = 20000
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
B = A  # any matrix: the product kernel does not need symmetry
Q = torch.randn((n, 16), dtype=torch.float64, device=dev, generator=g)
for _ in range(3):
    Y = M.symm_block(B, Q)
torch.cuda.synchronize()
print("ok")
