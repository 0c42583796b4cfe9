"""mds step at n = 10 000 (a float32 cosine-like triangle of synthetic points): kernel times + iterations"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from kameris_b200 import mds as M
dev = torch.device("cuda:0")
n, p, dim = int(os.environ.get("N", "10000")), 40, int(os.environ.get("DIM", "10"))
g = torch.Generator(device=dev); g.manual_seed(1)
pts = torch.randn((n, p), dtype=torch.float64, device=dev, generator=g) * torch.linspace(3.0, 0.3, p, dtype=torch.float64, device=dev)
d = torch.cdist(pts, pts)
iu = torch.triu_indices(n, n, 1, device=dev)
tri = d[iu[0], iu[1]].to(torch.float32).contiguous()
del d
def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
t_c, B = ev(lambda: M.center(tri, n))
Q = torch.randn((n, 16), dtype=torch.float64, device=dev, generator=g)
t_s, _ = ev(lambda: M.symm_block(B, Q), 10)
from kameris_b200 import _lib
_lib.load().kam_symm_set_tma(0)
t_s_reg, _ = ev(lambda: M.symm_block(B, Q), 10)
_lib.load().kam_symm_set_tma(1)
t_mm, _ = ev(lambda: B @ Q, 10)
stats = {}
M.top_eigenpairs(B, dim)                     # warm-up: cuSOLVER handles, lazy module loads
torch.cuda.synchronize()
t0 = time.perf_counter(); w, X = M.top_eigenpairs(B, dim, stats=stats); torch.cuda.synchronize(); t_e = time.perf_counter() - t0
print(json.dumps({"n": n, "center_ms": t_c, "center_GBps": (n * n * 8 * 4 + tri.numel() * 4) / t_c / 1e6,
                  "symm_block_ms": t_s, "symm_regs_kernel_ms": t_s_reg, "symm_GBps": n * n * 8 / t_s / 1e6, "torch_matmul_ms": t_mm,
                  "eig_s": t_e, **stats, "w": [float(x) for x in w[:4]]}))
