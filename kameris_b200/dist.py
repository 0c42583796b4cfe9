"""Device-level `distances` path: pairwise distance triangles over count / frequency vectors.

Host mirror of what generation_dists computes (`manhat` uint32, `info`
float32, packed strict upper triangle) plus the Gram-matrix metrics the north star adds
(euclid, cosine, pearson; float32).  The arithmetic runs in kameris_b200/csrc/{dist_tile,gram}.cu
through the C ABI; torch only owns device memory and streams.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check

_ELEM = {torch.uint8: 0, torch.uint16: 1, torch.uint32: 2, torch.uint64: 3, torch.float32: 4, torch.float64: 5,
         torch.int16: 1, torch.int32: 2, torch.int64: 3}   # signed views of the same bits are accepted

METRICS = ("manhat", "info", "euclid", "cosine", "pearson")


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _prep(x):
    if x.dim() != 2:
        raise ValueError("Vectors must have the same length")      # generation_dists @0x10002809c
    if x.stride(1) != 1:
        x = x.contiguous()
    if x.dtype not in _ELEM:
        raise TypeError("unsupported element type %s" % x.dtype)
    return x


def tri_len(n):
    return n * (n - 1) // 2


def manhattan_tri(x, rows=None, out=None, levels=None):
    """`manhat` of every pair i<j: uint32 packed triangle (as int32 storage).  levels: size the workspace for the
    tensor path up to column maxima summing to `levels` x d (default: 4 x d; see kam_manhattan_workspace_bytes_levels)."""
    lib = _lib.load()
    x = _prep(x)
    n, d = x.shape
    dev = x.device
    r0, r1 = rows if rows is not None else (0, n)
    if out is None:
        out = torch.zeros(max(tri_len(n), 1), dtype=torch.int32, device=dev)
    nbytes = lib.kam_manhattan_workspace_bytes(n, d) if levels is None else \
        lib.kam_manhattan_workspace_bytes_levels(n, d, int(levels))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_manhattan_tri(_ptr(x), _ELEM[x.dtype], n, d, x.stride(0), r0, r1, _ptr(out), _ptr(ws),
                                    ws.numel(), _stream(dev)))
    return out[:tri_len(n)]


def manhattan_rect(x, y, out=None):
    """uint32 manhat(x_i, y_j) for integer rows; float32 L1 when y is float32 (centroids)."""
    lib = _lib.load()
    x, y = _prep(x), _prep(y)
    m, d = x.shape
    c = y.shape[0]
    if y.shape[1] != d:
        raise ValueError("Vectors must have the same length")
    dev = x.device
    ws = torch.empty(lib.kam_manhattan_workspace_bytes(m, d) + lib.kam_manhattan_workspace_bytes(c, d),
                     dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if y.dtype == torch.float32:
            if out is None:
                out = torch.empty((m, c), dtype=torch.float32, device=dev)
            check(lib.kam_manhattan_rect_f32(_ptr(x), _ELEM[x.dtype], _ptr(y), m, c, d, x.stride(0), y.stride(0),
                                             _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
        else:
            if y.dtype != x.dtype:
                raise TypeError("x and y must have the same element type")
            if out is None:
                out = torch.empty((m, c), dtype=torch.int32, device=dev)
            check(lib.kam_manhattan_rect(_ptr(x), _ptr(y), _ELEM[x.dtype], m, c, d, x.stride(0), y.stride(0),
                                         _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
    return out


def info_tri(x, rows=None, out=None):
    """`info` (Jaccard distance on presence) of every pair i<j: float32 packed triangle."""
    lib = _lib.load()
    x = _prep(x)
    n, d = x.shape
    dev = x.device
    r0, r1 = rows if rows is not None else (0, n)
    if out is None:
        out = torch.zeros(max(tri_len(n), 1), dtype=torch.float32, device=dev)
    ws = torch.empty(lib.kam_info_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_info_tri(_ptr(x), _ELEM[x.dtype], n, d, x.stride(0), r0, r1, _ptr(out), _ptr(ws), ws.numel(),
                               _stream(dev)))
    return out[:tri_len(n)]


def gram_tri(x, metrics=("euclid", "cosine", "pearson"), euclid_on_freq=True, rows=None, force_f16=False):
    """euclid / cosine / pearson of every pair i<j from one tensor-core Gram pass: dict of float32
    packed triangles.  x: uint8/uint16/uint32 count vectors, or float32/float64 frequency vectors (the
    counts behind them are recovered and verified on the device; euclid is then always between the
    frequency vectors)."""
    lib = _lib.load()
    x = _prep(x)
    n, d = x.shape
    dev = x.device
    r0, r1 = rows if rows is not None else (0, n)
    bits = {"euclid": _lib.KAM_M_EUCLID, "cosine": _lib.KAM_M_COSINE, "pearson": _lib.KAM_M_PEARSON}
    mask = 0
    outs = {}
    for m in metrics:
        mask |= bits[m]
        outs[m] = torch.zeros(max(tri_len(n), 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        lib.kam_gram_force_f16(int(force_f16))         # test hook: fp16 tensor path instead of uint8
        try:
            for size_fn in (lib.kam_gram_workspace_bytes, lib.kam_gram_workspace_bytes_int):
                ws = torch.empty(max(size_fn(n, d), 256), dtype=torch.uint8, device=dev)
                rc = lib.kam_gram_tri(_ptr(x), _ELEM[x.dtype], n, d, x.stride(0), r0, r1, mask, int(euclid_on_freq),
                                      _ptr(outs.get("euclid")), _ptr(outs.get("cosine")), _ptr(outs.get("pearson")),
                                      _ptr(ws), ws.numel(), _stream(dev))
                if rc != _lib.KAM_E_WORKSPACE:     # the integer fallback asks for its (larger) workspace
                    break
                del ws
            check(rc)
        finally:
            lib.kam_gram_force_f16(0)
    return {m: t[:tri_len(n)] for m, t in outs.items()}


def balanced_row_ranges(n, parts):
    """Contiguous row blocks [r0, r1) with (nearly) equal numbers of pairs (i, j>i): row i owns
    n-1-i pairs, so the cut points follow the triangle's area, not the row count."""
    total = tri_len(n)
    cuts = [0]
    for p in range(1, parts):
        target = total * p / parts
        # pairs in rows [0, r) = r*n - r(r+1)/2  -> solve for r
        r = int(n - 0.5 - np.sqrt(max((n - 0.5) ** 2 - 2 * target, 0.0)))
        cuts.append(min(max(r, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(parts)]


def bench_lines(torch_mod, dev, hbm_peak, tf_peak):
    """Secondary bench lines (pairs/s) -- see bench.py."""
    from . import dist_bench
    return dist_bench.run(torch_mod, dev, hbm_peak, tf_peak)
