"""Device-level `mds` step: classical multidimensional scaling of a `.mm-dist` triangle.

Host mirror of mds(delta, dim), kameris/job_steps/mds.py:10-35.  The two O(n^2) parts run in
hand-written kernels through the C ABI (kameris_b200/csrc/mds.cu): the double-centring of the squared
distances (kam_mds_center) and the matrix-block product B Q (kam_symm_block) of a block subspace
iteration with Rayleigh-Ritz extraction, which replaces scipy's ARPACK call eigsh(B, k=dim) of
mds.py:26 (same selection: the `dim` eigenvalues of largest magnitude).  The O(n * 16^2) pieces
(thin QR, the 16 x 16 Ritz problem) are torch calls on the device; torch otherwise only owns memory.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check

BLOCK = 16          # KAM_MDS_BLOCK: columns of the iteration block


def center(tri, n):
    """packed triangle on the device (int32/uint32 manhat bits or float32) -> B (n, n) float64."""
    lib = _lib.load()
    dev = tri.device
    elem = _lib.KAM_F32 if tri.dtype == torch.float32 else _lib.KAM_U32
    if tri.dtype not in (torch.float32, torch.int32, torch.uint32):
        raise TypeError("a .mm-dist triangle holds uint32 or float32 values")
    if tri.numel() != n * (n - 1) // 2:
        raise ValueError("triangle length does not match n")
    B = torch.empty((n, n), dtype=torch.float64, device=dev)
    ws = torch.empty(lib.kam_mds_workspace_bytes(n), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_mds_center(tri.data_ptr() if tri.numel() else None, elem, n, B.data_ptr(), ws.data_ptr(),
                                 ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return B


def symm_workspace(n, device):
    return torch.empty(_lib.load().kam_symm_workspace_bytes(n), dtype=torch.uint8, device=device)


def symm_block(B, Q, out=None, workspace=None):
    """Y = B Q for symmetric B (n, n) float64 and Q (n, 16) float64, both contiguous on the device."""
    lib = _lib.load()
    n = B.shape[0]
    assert B.dtype == torch.float64 and Q.dtype == torch.float64 and Q.shape == (n, BLOCK)
    assert B.is_contiguous() and Q.is_contiguous()
    if out is None:
        out = torch.empty_like(Q)
    ws = workspace if workspace is not None else symm_workspace(n, B.device)
    with torch.cuda.device(B.device):
        check(lib.kam_symm_block(B.data_ptr(), n, Q.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream(B.device).cuda_stream))
    return out


def top_eigenpairs(B, dim, tol=1e-12, max_iter=5000, seed=0, stats=None, check_every=4):
    """The `dim` eigenpairs of largest |eigenvalue| of the symmetric matrix B (what eigsh(B, k=dim) with its
    default which='LM' returns, mds.py:26), eigenvalues in descending algebraic order.

    Block subspace iteration: Q <- orth(B Q) with a 16-column block; every `check_every` sweeps the Ritz
    pairs of the 16 x 16 projected matrix Q^T B Q are formed and the iteration stops when every wanted
    pair has |B x - theta x| <= tol * max|theta| (the only host synchronisation).  Convergence of pair i is
    geometric with ratio |lambda_17 / lambda_i|."""
    n = B.shape[0]
    if not 1 <= dim < n:
        raise ValueError("k must be between 1 and the order of the matrix minus 1")     # scipy eigsh's condition
    if dim > BLOCK - 2 and n > BLOCK:
        raise ValueError("dimensions > %d not supported by this build" % (BLOCK - 2))
    dev = B.device
    b = min(BLOCK, n)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    Q = torch.zeros((n, BLOCK), dtype=torch.float64, device=dev)
    Q[:, :b] = torch.linalg.qr(torch.randn((n, b), dtype=torch.float64, device=dev, generator=g))[0]
    Y = torch.empty_like(Q)
    ws = symm_workspace(n, dev)
    # orthonormalisation of B Q: Cholesky QR twice (three small launches each, no host round trip); a
    # rank-deficient block (B of rank < 16: exactly embeddable inputs) makes the factorisation fail, which
    # is noticed at the next convergence check and answered by Householder QR from the last good block
    householder = False
    bad = torch.zeros((), dtype=torch.int32, device=dev)
    Q_good = Q.clone()

    def orth(Yb):
        nonlocal bad
        if householder:
            return torch.linalg.qr(Yb)[0]
        Z = Yb
        for _ in range(2):
            L, info = torch.linalg.cholesky_ex(Z.T @ Z)
            bad = bad | (info != 0).to(torch.int32)
            Z = torch.linalg.solve_triangular(L, Z.T, upper=False).T
        return Z

    it = 0
    while True:
        symm_block(B, Q, out=Y, workspace=ws)
        it += 1
        if b == n or it % check_every == 0 or it >= max_iter:
            if not householder and (int(bad) != 0 or not bool(torch.isfinite(Y).all())):
                householder = True                         # redo the last sweeps from the last good block
                Q.copy_(Q_good)
                bad.zero_()
                symm_block(B, Q, out=Y, workspace=ws)
            T = Q[:, :b].T @ Y[:, :b]
            theta, S = torch.linalg.eigh((T + T.T) * 0.5)
            order = torch.argsort(theta.abs(), descending=True)[:dim]
            X = Q[:, :b] @ S[:, order]
            R = Y[:, :b] @ S[:, order] - X * theta[order]
            res = torch.linalg.vector_norm(R, dim=0)
            scale = theta.abs().max().clamp_min(1e-300)
            if bool((res <= tol * scale).all()) or b == n or it >= max_iter:
                break
            Q_good.copy_(Q)
        Q[:, :b] = orth(Y[:, :b])
    converged = bool((res <= tol * scale).all()) or b == n
    if not converged:
        import logging
        logging.getLogger('kameris').warning(
            'mds: the subspace iteration stopped after %d sweeps with relative residual %.3g (tolerance %.3g): '
            'the embedding is approximate', it, float((res / scale).max()), tol)
    if stats is not None:
        stats.update({"iterations": it, "residual": float((res / scale).max()), "block": b, "converged": converged,
                      "orthonormalisation": "householder" if householder else "cholesky-qr2"})
    w = theta[order]
    desc = torch.argsort(w, descending=True)
    return w[desc], X[:, desc]


def mds_from_triangle(tri, n, dim, stats=None):
    """points (n, <= dim) float64 on the host, columns in descending eigenvalue order; columns whose
    eigenvalue is negative are dropped (NaN columns, mds.py:33-35) after a warning like mds.py:27-30."""
    import logging
    B = center(tri, n)
    w, X = top_eigenpairs(B, dim, stats=stats)
    if bool((w < 0).any()):
        logging.getLogger('kameris.mds').warning('some eigenvalues were negative')
    keep = w >= 0
    pts = X[:, keep] * torch.sqrt(w[keep])
    return pts.cpu().numpy()


def mds(delta, dim, device="cuda:0"):
    """mds(delta, dim) of the reference for a symmetric distance matrix (ndarray, any numeric dtype): the
    upper triangle goes to the device as float32 (uint32 for integer matrices), exactly what a `.mm-dist`
    file holds."""
    delta = np.asarray(delta)
    n = delta.shape[0]
    if delta.shape != (n, n):
        raise ValueError("delta must be square")
    iu = np.triu_indices(n, 1)
    tri = delta[iu]
    tri = tri.astype(np.uint32).view(np.int32) if tri.dtype.kind in "ui" else tri.astype(np.float32)
    return mds_from_triangle(torch.from_numpy(np.ascontiguousarray(tri)).to(device), n, dim)
