"""Secondary benchmark line of the classify pre-processing for bench.py: n count vectors of D = 4^6 (9 kb
genomes, uint16), the two transformers of kameris/job_steps/classify.py:29-50.

  col_stats  : kam_col_stats, two sweeps of n x d x 2 B (HBM-bound; the matrix fits L2 only for small n)
  col_scale  : kam_col_scale, n x d x (2 + 8) B
  svd        : randomized TruncatedSVD (n_components = ceil(avg_nnz / 10), 5 LU-normalised power iterations):
               library GEMM / LU / QR / SVD calls on the device
  step       : scaler fit + transform + SVD fit_transform through the sklearn-compatible transformers, host
               uint16 matrix in, host float64 reduced features out
"""
import time

import numpy as np
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


def run(torch_mod, dev, hbm_peak, n=10000, k=6):
    from . import preprocess as P
    from .dist_bench import synth_counts
    from .job_steps import classify as C
    x = synth_counts(torch_mod, dev, n, 9000, k, 20261007)
    d = x.shape[1]
    # the two kernels on a matrix that does not fit L2 (10 copies of the rows: 100 000 x 4096 uint16 = 0.8 GB)
    big = x.repeat(10, 1)
    nb = big.shape[0]
    t_stats, (mean, var, cnt) = _ev(lambda: P.column_stats(big), 10)
    out = torch.empty((nb, d), dtype=torch.float64, device=dev)
    t_scale, _ = _ev(lambda: P.scale_columns(big, P.scale_from_variance(var, mean, cnt), out=out), 5)
    del big, out
    mean, var, cnt = P.column_stats(x)
    scale = P.scale_from_variance(var, mean, cnt)
    out = P.scale_columns(x, scale)
    host = x.cpu().numpy()
    ncomp = int(np.ceil(C.avg_num_nonzero_entries(host[:256]) * 0.1))
    t_svd, _ = _ev(lambda: P.truncated_svd_fit(out, ncomp, random_state=0), 2)

    def step():
        sc = C.StandardScaler(with_mean=False, device_output=True)
        return C.TruncatedSVD(n_components=ncomp, random_state=0).fit_transform(sc.fit_transform(host))
    step()
    t0 = time.perf_counter()
    red = step()
    t_step = time.perf_counter() - t0
    b_stats, b_scale = 2.0 * nb * d * 2, nb * d * (2.0 + 8.0)
    return {"metric": "classify_preprocess_vectors_per_sec", "value": n / t_step, "unit": "vectors/s", "ms": t_step * 1e3,
            "config": {"workload": "%d count vectors (9 kb genomes, k=%d, D=%d, uint16), n_components=%d" % (n, k, d, ncomp),
                       "reduced_shape": list(red.shape)},
            "kernels_on": "%d x %d uint16 (0.8 GB, larger than L2)" % (nb, d),
            "col_stats": {"ms": t_stats * 1e3, "GBps": b_stats / t_stats / 1e9, "frac": b_stats / t_stats / 1e9 / hbm_peak},
            "col_scale": {"ms": t_scale * 1e3, "GBps": b_scale / t_scale / 1e9, "frac": b_scale / t_scale / 1e9 / hbm_peak},
            "svd": {"ms": t_svd * 1e3, "what": "library GEMM / LU / QR / SVD (torch on the device), float64"}}, host
