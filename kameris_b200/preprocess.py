"""Device-level classify pre-processing: the two transformers kameris puts in front of every classifier
(kameris/job_steps/classify.py:29-50, build_pipeline): `StandardScaler(with_mean=False)` and
`TruncatedSVD(n_components=ceil(avg_nnz * dim_reduce_fraction))`.

The arithmetic of both lives in scikit-learn (setup.py:85 pins 0.19.1; the copy that can run here, and
that the tests compare with, is the installed one):
  * column statistics and the transform of the scaler: hand-written kernels through the C ABI
    (kameris_b200/csrc/colstats.cu: kam_col_stats, kam_col_scale) over the count / frequency matrix as it
    is stored (uint16 / uint32 / float32 / float64), float64 results;
  * the randomized truncated SVD (sklearn.utils.extmath._randomized_svd, Halko et al. Algorithm 4.3 with
    LU-normalised power iterations) is tall-skinny GEMMs + thin LU / QR / SVD: plain library calls
    (torch.matmul / torch.linalg on the device), with the Gaussian test matrix drawn on the host from
    numpy's RandomState exactly like sklearn does, so the same `random_state` gives the same subspace.
torch owns the memory; there is no CPU fallback (CUDA tensors only).
"""
import numpy as np
import torch

from . import _lib
from ._lib import check
from .dist import _ELEM

_STAT_TYPES = (torch.uint8, torch.uint16, torch.uint32, torch.int16, torch.int32, torch.float32, torch.float64)


def _require_cuda(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError("kameris_b200.preprocess works on CUDA tensors only (no CPU fallback)")
    if x.dim() != 2:
        raise ValueError("expected a 2-D feature matrix")
    if x.dtype not in _STAT_TYPES:
        raise TypeError("unsupported element type %s" % x.dtype)
    return x if x.stride(1) == 1 else x.contiguous()


def column_stats(x):
    """(mean_, var_, n_samples_seen_ per column) of sklearn's StandardScaler on an (n, d) device matrix:
    float64 mean and population variance over the non-NaN entries of every column."""
    lib = _lib.load()
    x = _require_cuda(x)
    n, d = x.shape
    if n < 1 or d < 1:
        raise ValueError("empty feature matrix")
    dev = x.device
    mean = torch.empty(d, dtype=torch.float64, device=dev)
    var = torch.empty(d, dtype=torch.float64, device=dev)
    cnt = torch.empty(d, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.kam_colstats_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_col_stats(x.data_ptr(), _ELEM[x.dtype], n, d, x.stride(0), mean.data_ptr(), var.data_ptr(),
                                cnt.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return mean, var, cnt


def scale_from_variance(var, mean, count):
    """scale_ = sqrt(var_) with the near-constant columns set to 1 (sklearn/preprocessing/_data.py
    _is_constant_feature + _handle_zeros_in_scale: var <= n eps var + (n mean eps)^2)."""
    eps = float(np.finfo(np.float64).eps)
    n = count.to(torch.float64)
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale = torch.sqrt(var)
    scale[constant] = 1.0
    return scale


def scale_columns(x, scale, out=None):
    """x / scale_ per column, float64 (StandardScaler.transform with with_mean=False)."""
    lib = _lib.load()
    x = _require_cuda(x)
    n, d = x.shape
    dev = x.device
    if scale.shape != (d,) or scale.dtype != torch.float64 or scale.device != dev:
        raise ValueError("scale must be a float64 vector of d entries on the same device")
    if out is None:
        out = torch.empty((n, d), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_col_scale(x.data_ptr(), _ELEM[x.dtype], n, d, x.stride(0), scale.contiguous().data_ptr(),
                                out.data_ptr(), out.stride(0), torch.cuda.current_stream(dev).cuda_stream))
    return out


def random_state_of(seed):
    """sklearn.utils.check_random_state."""
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, np.random.RandomState):
        return seed
    return np.random.RandomState(seed)


def svd_flip_v(u, vt):
    """sklearn.utils.extmath.svd_flip(u, v, u_based_decision=False): the entry of largest magnitude of every
    row of vt becomes positive."""
    idx = vt.abs().argmax(dim=1)
    signs = torch.sign(vt[torch.arange(vt.shape[0], device=vt.device), idx])
    if u is not None:
        u = u * signs[None, :]
    return u, vt * signs[:, None]


def _lu_normalizer(y):
    """scipy.linalg.lu(y, permute_l=True)[0] = P L  (cuSOLVER getrf: MAGMA's batched LU prints a banner on
    stdout for matrices of this size)"""
    prev = torch.backends.cuda.preferred_linalg_library()
    torch.backends.cuda.preferred_linalg_library("cusolver")
    try:
        p, l, _ = torch.linalg.lu(y)
    finally:
        torch.backends.cuda.preferred_linalg_library(prev)
    return p @ l


def randomized_svd(a, n_components, n_iter=5, n_oversamples=10, random_state=None):
    """sklearn.utils.extmath._randomized_svd(M, n_components, n_iter=n_iter, flip_sign=False,
    power_iteration_normalizer='auto', transpose='auto') on a float64 device matrix: (U, s, Vt)."""
    if not (a.is_cuda and a.dtype == torch.float64 and a.dim() == 2):
        raise RuntimeError("randomized_svd takes a float64 CUDA matrix")
    n_samples, n_features = a.shape
    n_random = n_components + n_oversamples
    transpose = n_samples < n_features
    m = a.T if transpose else a
    rs = random_state_of(random_state)
    q = torch.from_numpy(rs.normal(size=(m.shape[1], n_random))).to(a.device)
    normalizer = _lu_normalizer if n_iter > 2 else (lambda y: y)
    for _ in range(n_iter):
        q = normalizer(m @ q)
        q = normalizer(m.T @ q)
    q, _ = torch.linalg.qr(m @ q, mode="reduced")
    b = q.T @ m
    uhat, s, vt = torch.linalg.svd(b, full_matrices=False)
    u = q @ uhat
    if transpose:
        return vt[:n_components].T, s[:n_components], u[:, :n_components].T
    return u[:, :n_components], s[:n_components], vt[:n_components]


def truncated_svd_fit(x, n_components, n_iter=5, n_oversamples=10, random_state=None):
    """TruncatedSVD(algorithm='randomized').fit_transform on a float64 device matrix: dict with components_
    (n_components, d), singular_values_, explained_variance_, explained_variance_ratio_ and the transformed
    matrix X components_^T."""
    if n_components > x.shape[1]:
        raise ValueError("n_components(%d) must be <= n_features(%d)." % (n_components, x.shape[1]))
    u, s, vt = randomized_svd(x, n_components, n_iter, n_oversamples, random_state)
    _, vt = svd_flip_v(None, vt)
    vt = vt.contiguous()
    xt = x @ vt.T
    exp_var = xt.var(dim=0, unbiased=False)
    full_var = x.var(dim=0, unbiased=False).sum()
    return {"components_": vt, "singular_values_": s, "explained_variance_": exp_var,
            "explained_variance_ratio_": exp_var / full_var, "transformed": xt}
