"""Secondary benchmark line of the `mds` step for bench.py: n points of a 40-dimensional Euclidean
configuration (float32 `.mm-dist` triangle resident on the device), dimensions = 10.

  center      : kam_mds_center, 4 sweeps of n^2 x 8 B (HBM-bound)
  symm_block  : one B Q product (n x n by n x 16, float64).  B is streamed once (n^2 x 8 B), but at 16
                columns the product needs 2 DFMA per byte: the FP64 pipe, not HBM, is the bound (cuBLAS
                dgemm on the same shapes is reported beside it).
  step        : centring + eigen-solve + points (warm)
"""
import time

import torch


def _ev(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3, r


def run(dev, hbm_peak, n=10000, dim=10):
    from . import mds as M
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    pts = torch.randn((n, 40), dtype=torch.float64, device=dev, generator=g) * \
        torch.linspace(3.0, 0.3, 40, dtype=torch.float64, device=dev)
    d = torch.cdist(pts, pts)
    iu = torch.triu_indices(n, n, 1, device=dev)
    tri = d[iu[0], iu[1]].to(torch.float32).contiguous()
    del d, iu
    t_c, B = _ev(lambda: M.center(tri, n), 3)
    Q = torch.randn((n, M.BLOCK), dtype=torch.float64, device=dev, generator=g)
    t_s, _ = _ev(lambda: M.symm_block(B, Q), 10)
    t_mm, _ = _ev(lambda: B @ Q, 10)
    del B
    M.mds_from_triangle(tri, n, dim)                       # warm-up (cuSOLVER handles)
    torch.cuda.synchronize()
    stats = {}
    t0 = time.perf_counter()
    out = M.mds_from_triangle(tri, n, dim, stats=stats)
    t_step = time.perf_counter() - t0
    sample = tri.cpu().numpy()
    return {"metric": "mds_points_per_sec", "value": n / t_step, "unit": "points/s", "ms": t_step * 1e3,
            "config": {"workload": "n=%d points, float32 triangle on the device, dimensions=%d" % (n, dim),
                       "sweeps": stats.get("iterations"), "residual": stats.get("residual"),
                       "points_shape": list(out.shape)},
            "center": {"ms": t_c * 1e3, "GBps": (4 * n * n * 8 + tri.numel() * 4) / t_c / 1e9,
                       "frac": (4 * n * n * 8 + tri.numel() * 4) / t_c / 1e9 / hbm_peak},
            "symm_block": {"ms": t_s * 1e3, "GBps": n * n * 8 / t_s / 1e9, "frac": n * n * 8 / t_s / 1e9 / hbm_peak,
                           "dfma_per_clk_per_sm": n * n * 16 / t_s / 148 / 1.965e9, "cublas_dgemm_ms": t_mm * 1e3}}, sample
