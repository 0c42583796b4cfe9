"""ctypes front-end of oracle/kameris_oracle.c plus the float64 definitions of the metrics the
reference does not have.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (kameris_b200/) never does.

Reference anchors: generation_cgr_mac_avx2 @0x10000d330 (CGR fill),
generation_dists_mac_avx2 @0x10001b8c0 (row lambda), kameris/job_steps/backend.py:42-56
(frequencies).  Pinning: CGR counts against tests/golden/refcode_cgr.json.gz (outputs of the
reference's own machine code, oracle/refcode/); manhat/info against
tests/golden/refcode_dists.json.gz (outputs of the reference's own row-lambda machine code, all six
element types); euclid/cosine/pearson do not
exist in the reference and are defined as scipy.spatial.distance.cdist in float64.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkameris_oracle.so")
_lib = None

ELEM_DTYPES = [np.uint8, np.uint16, np.uint32, np.uint64, np.float32, np.float64]  # A.4 element_type
_SFX = ["u8", "u16", "u32", "u64", "f32", "f64"]


def _host_tag():
    """identity of the host CPU: the .so is built -march=native and must not run elsewhere."""
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        import hashlib
        return hashlib.sha1("\n".join(model).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    """(Re)build libkameris_oracle.so when missing, stale, or built on a different host CPU."""
    src = os.path.join(_HERE, "kameris_oracle.c")
    tagf = _SO + ".host"
    tag = _host_tag()
    old = open(tagf).read().strip() if os.path.exists(tagf) else ""
    if (force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src) or old != tag):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libkameris_oracle.so"])
        with open(tagf, "w") as f:
            f.write(tag)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        L = _lib
        L.orc_fasta_record0.restype = ctypes.c_int64
        L.orc_fasta_record0.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        for s in ("u16", "u32"):
            f = getattr(L, "orc_cgr_count_" + s)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
            g = getattr(L, "orc_freq_" + s)
            g.restype = None
            g.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        for s in _SFX:
            f = getattr(L, "orc_manhat_" + s)
            f.restype = ctypes.c_uint32
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
            g = getattr(L, "orc_info_" + s)
            g.restype = ctypes.c_float
            g.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.orc_cgr_batch.restype = ctypes.c_int
        L.orc_cgr_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_cgr_freq_batch.restype = ctypes.c_int
        L.orc_cgr_freq_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.orc_dists_triangle.restype = ctypes.c_int
        L.orc_dists_triangle.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_manhat_rect.restype = ctypes.c_int
        L.orc_manhat_rect.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                      ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return _lib


# ---- the reference's own machine code (oracle/refcode/refcode.c) -----------------------------------------
_REF_SO = os.path.join(_HERE, "_ref", "librefcode.so")
REF_BLOB = os.path.join(_HERE, "_ref", "cgr_fill_u16.bin")
REF_PE = "/root/reference/kameris/scripts/generation_cgr_windows_avx2.exe"
_ref = False


def _has_bmi2():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " bmi2" in line
    except OSError:
        pass
    return False


def refcode_extract_blob():
    """Build step (container with /root/reference): write the CGR fill function's 0x168 bytes + the order
    string into oracle/_ref/cgr_fill_u16.bin, a build output that travels to the GPU box like the .so files."""
    lib = ctypes.CDLL(_REF_SO)
    rc = lib.refcode_extract(REF_PE.encode(), REF_BLOB.encode())
    if rc != 0:
        raise RuntimeError("refcode_extract failed: %d" % rc)
    return REF_BLOB


def refcode():
    """Runner for the reference's own CGR-fill machine code: loaded from the reference executable where
    /root/reference is mounted, else from the blob the build extracted into oracle/_ref/.  None when
    neither is there (or the host has no BMI2: the code uses shlx)."""
    global _ref
    if _ref is False:
        _ref = None
        if os.path.exists(_REF_SO) and _has_bmi2():
            lib = ctypes.CDLL(_REF_SO)
            if os.path.exists(REF_PE):
                rc = lib.refcode_load(REF_PE.encode())
            elif os.path.exists(REF_BLOB):
                rc = lib.refcode_load_blob(REF_BLOB.encode())
            else:
                rc = -1
            if rc == 0:
                lib.refcode_order.restype = ctypes.c_char_p
                lib.refcode_cgr_u16_batch.restype = ctypes.c_int
                lib.refcode_cgr_u16_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64,
                                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
                if lib.refcode_order() == b"ACGT":
                    _ref = lib
    return _ref


_REFD_SO = os.path.join(_HERE, "_ref", "librefdists.so")
REFD_BLOB = os.path.join(_HERE, "_ref", "dists_image.bin")
REFD_PE = "/root/reference/kameris/scripts/generation_dists_windows_avx2.exe"
_refd = False


def refdists_extract_blob():
    """Build step (container with /root/reference): the loaded image of generation_dists_windows_avx2.exe
    (sections at their virtual addresses) -> oracle/_ref/dists_image.bin, a build output that travels to the
    GPU box like the .so files."""
    lib = ctypes.CDLL(_REFD_SO)
    rc = lib.refdists_extract(REFD_PE.encode(), REFD_BLOB.encode())
    if rc != 0:
        raise RuntimeError("refdists_extract failed: %d" % rc)
    return REFD_BLOB


def _has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " avx2" in line
    except OSError:
        pass
    return False


def refdists():
    """Runner for the reference's own manhat / info row-task machine code (oracle/refcode/refdists.c), from the
    executable where /root/reference is mounted, else from the image the build extracted into oracle/_ref/.
    None when neither is there, the host has no AVX2, or the image's preferred address range is taken."""
    global _refd
    if _refd is False:
        _refd = None
        if os.path.exists(_REFD_SO) and _has_avx2():
            lib = ctypes.CDLL(_REFD_SO)
            if os.path.exists(REFD_PE):
                rc = lib.refdists_load(REFD_PE.encode())
            elif os.path.exists(REFD_BLOB):
                rc = lib.refdists_load_blob(REFD_BLOB.encode())
            else:
                rc = -1
            if rc == 0:
                lib.refdists_run_threaded.restype = ctypes.c_int
                lib.refdists_run_threaded.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
                _refd = lib
    return _refd


def refdists_triangle(x, want_manhat=True, want_info=False, threads=0):
    """(manhat uint32 triangle | None, info float32 triangle | None) of the rows of x by the REFERENCE'S row-task
    machine code on `threads` host threads (what generation_dists' executor runs)."""
    lib = refdists()
    if lib is None:
        raise RuntimeError("the reference's machine code is not available here")
    x = np.ascontiguousarray(x)
    n, d = x.shape
    elem = [np.dtype(t) for t in ELEM_DTYPES].index(x.dtype)
    pairs = n * (n - 1) // 2
    man = np.zeros(pairs, dtype=np.uint32) if want_manhat else None
    inf = np.zeros(pairs, dtype=np.float32) if want_info else None
    rc = lib.refdists_run_threaded(elem, _ptr(x), n, d, _ptr(man) if want_manhat else None,
                                   _ptr(inf) if want_info else None, threads or host_threads())
    if rc != 0:
        raise RuntimeError("refdists_run_threaded failed: %d" % rc)
    return man, inf


def refcode_cgr_batch(chars, start, k, out, threads=0):
    """uint16 CGR count vectors of a batch of records by the REFERENCE'S machine code on `threads` host
    threads (what generation_cgr's executor does per file, minus the file reads)."""
    lib = refcode()
    if lib is None:
        raise RuntimeError("the reference's machine code is not available here")
    n = len(start) - 1
    assert out.dtype == np.uint16 and out.shape == (n, 4 ** k) and out.flags.c_contiguous
    rc = lib.refcode_cgr_u16_batch(_ptr(chars), _ptr(start), n, k, _ptr(out), threads or host_threads())
    if rc != 0:
        raise RuntimeError("refcode_cgr_u16_batch failed: %d" % rc)
    return out


def host_threads():
    """std::thread::hardware_concurrency() of the reference (cgr@0x100015641, dists@0x1000017ae)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fasta_record0(data: bytes) -> bytes:
    """A.1: first FASTA record of a file's bytes.  Raises ValueError when there is none."""
    out = ctypes.create_string_buffer(max(len(data), 1))
    n = lib().orc_fasta_record0(data, len(data), out)
    if n < 0:
        raise ValueError("no FASTA record in input")
    return out.raw[:n]


def cgr_count(record: bytes, k: int, bits: int = 16) -> np.ndarray:
    """A.2: CGR count vector (4^k,) of uint16/uint32 for one record."""
    dt = np.uint16 if bits == 16 else np.uint32
    out = np.zeros(4 ** k, dtype=dt)
    fn = lib().orc_cgr_count_u16 if bits == 16 else lib().orc_cgr_count_u32
    if fn(record, len(record), k, _ptr(out)) != 0:
        raise ValueError("k is too large to fit in the index")
    return out


def cgr_count_numpy(record: bytes, k: int, bits: int = 16) -> np.ndarray:
    """Independent numpy restatement of A.2 (vectorised; used to cross-check the C restatement)."""
    raw = np.frombuffer(record, dtype=np.uint8)
    valid = (raw == 65) | (raw == 67) | (raw == 71) | (raw == 84)
    pos = np.nonzero(valid)[0]
    c = raw[pos]
    xb = ((c == 71) | (c == 84)).astype(np.uint64)
    yb = ((c == 65) | (c == 84)).astype(np.uint64)
    n = len(c)
    x = np.zeros(n, dtype=np.uint64)
    y = np.zeros(n, dtype=np.uint64)
    for t in range(min(k, n)):               # bit (k-1-t) of the window ending at j is base j-t
        sh = np.uint64(k - 1 - t)
        x[t:] |= xb[: n - t] << sh
        y[t:] |= yb[: n - t] << sh
    j = np.arange(n)
    part = j < k - 1                         # centre bit still inside the window (Q1)
    centre = np.zeros(n, dtype=np.uint64)
    centre[part] = (np.int64(1) << (np.int64(k - 2) - j[part])).astype(np.uint64)
    x |= centre
    y |= centre
    emit = pos >= k - 1
    idx = ((y << np.uint64(k)) | x)[emit]
    cnt = np.bincount(idx.astype(np.int64), minlength=4 ** k).astype(np.uint64)
    return cnt.astype(np.uint16 if bits == 16 else np.uint32)   # astype wraps mod 2^bits (Q3)


def frequencies(count: np.ndarray) -> np.ndarray:
    """backend.py:49: cgr / cgr.sum() -> float64 (0/0 -> nan like numpy)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return count / count.sum()


def cgr_batch(chars: np.ndarray, start: np.ndarray, k: int, bits: int = 16, threads: int = 0) -> np.ndarray:
    """run<T>: threaded batch over concatenated records (uint8 chars, uint64 start[n+1])."""
    n = len(start) - 1
    out = np.empty((n, 4 ** k), dtype=np.uint16 if bits == 16 else np.uint32)
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    start = np.ascontiguousarray(start, dtype=np.uint64)
    rc = lib().orc_cgr_batch(_ptr(chars), _ptr(start), n, k, bits, _ptr(out), threads or host_threads())
    if rc != 0:
        raise ValueError("bad k/bits")
    return out


def cgr_freq_batch(chars, start, k, bits=16, f64=False, threads=0, out=None):
    """`kmers` step with mode=frequencies (run<T> + backend.py:42-56), threaded; float32/float64."""
    n = len(start) - 1
    if out is None:
        out = np.empty((n, 4 ** k), dtype=np.float64 if f64 else np.float32)
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    start = np.ascontiguousarray(start, dtype=np.uint64)
    rc = lib().orc_cgr_freq_batch(_ptr(chars), _ptr(start), n, k, bits, int(f64), _ptr(out),
                                  threads or host_threads())
    if rc != 0:
        raise ValueError("bad k/bits")
    return out


def _elem(x):
    for i, dt in enumerate(ELEM_DTYPES):
        if x.dtype == np.dtype(dt):
            return i
    raise TypeError("unsupported element type %s" % x.dtype)


def manhat(a: np.ndarray, b: np.ndarray) -> int:
    e = _elem(a)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    return int(getattr(lib(), "orc_manhat_" + _SFX[e])(_ptr(a), _ptr(b), a.size))


def info(a: np.ndarray, b: np.ndarray) -> np.float32:
    e = _elem(a)
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    return np.float32(getattr(lib(), "orc_info_" + _SFX[e])(_ptr(a), _ptr(b), a.size))


def dists_triangle(x: np.ndarray, want_manhat=True, want_info=False, threads: int = 0):
    """generation_dists: packed strict upper triangle(s) for rows of x (n, d)."""
    x = np.ascontiguousarray(x)
    n, d = x.shape
    m = n * (n - 1) // 2
    man = np.zeros(m, dtype=np.uint32) if want_manhat else None
    inf = np.zeros(m, dtype=np.float32) if want_info else None
    rc = lib().orc_dists_triangle(_ptr(x), n, d, _elem(x), _ptr(man) if want_manhat else None,
                                  _ptr(inf) if want_info else None, threads or host_threads())
    assert rc == 0
    return man, inf


def manhat_rect(x: np.ndarray, y: np.ndarray, threads: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    out = np.zeros((x.shape[0], y.shape[0]), dtype=np.uint32)
    rc = lib().orc_manhat_rect(_ptr(x), _ptr(y), x.shape[0], y.shape[0], x.shape[1], _elem(x), _ptr(out),
                               threads or host_threads())
    assert rc == 0
    return out


def triangle_index(n: int, i: int, j: int) -> int:
    """dists@0x10001bf7e: position of (i<j) in the packed strict upper triangle."""
    return n * i - i * (i + 3) // 2 + j - 1


# ---- metrics absent from the reference: float64 scipy definitions ---------------
_SCIPY = {"euclid": "euclidean", "cosine": "cosine", "pearson": "correlation", "cityblock": "cityblock"}


def cdist_triangle(x: np.ndarray, metric: str) -> np.ndarray:
    """float64 packed strict upper triangle of scipy's pdist for 'euclid' | 'cosine' | 'pearson'."""
    from scipy.spatial.distance import pdist
    return pdist(np.asarray(x, dtype=np.float64), _SCIPY[metric])


def cdist_rect(x: np.ndarray, y: np.ndarray, metric: str) -> np.ndarray:
    from scipy.spatial.distance import cdist
    return cdist(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), _SCIPY[metric])


# ---- mds job step (kameris/job_steps/mds.py:10-35) ------------------------------------------------------
def full_from_triangle(tri: np.ndarray, n: int) -> np.ndarray:
    """what kameris_formats.dist_reader.read_matrix hands to mds (mds.py:39): the symmetric N x N matrix."""
    m = np.zeros((n, n), dtype=tri.dtype)
    iu = np.triu_indices(n, 1)
    m[iu] = tri
    m.T[iu] = tri
    return m


def mds_reference(delta: np.ndarray, dim: int) -> np.ndarray:
    """Restatement of mds(delta, dim), kameris/job_steps/mds.py:10-35: classical MDS.
    B = -0.5 (D^2 - colmean - rowmean + mean) (mds.py:14-24); the `dim` eigenpairs of largest magnitude
    (scipy eigsh default which='LM', mds.py:26), in DEscending algebraic order after the fliplr of
    mds.py:32; points = V sqrt(lambda); columns with a negative eigenvalue become NaN and are dropped
    (mds.py:34).  Deterministic (full symmetric eigendecomposition instead of ARPACK); eigenvector signs
    are arbitrary in both.  Pinned against tests/golden/ref_mds.npz (outputs of the reference function
    itself, oracle/refpy/gen_golden_mds.py)."""
    delta = np.asarray(delta).astype(float)
    n = delta.shape[0]
    sq = delta ** 2
    tot = sq.sum(axis=0) / n
    b = -0.5 * (sq - tot[None, :] - tot[:, None] + tot.sum() / n)
    w, v = np.linalg.eigh((b + b.T) / 2)
    pick = np.argsort(-np.abs(w), kind="stable")[:dim]
    pick = pick[np.argsort(-w[pick], kind="stable")]              # descending algebraic order
    with np.errstate(invalid="ignore"):
        pts = v[:, pick] * np.sqrt(w[pick])[None, :]
    return pts[:, ~np.isnan(pts).any(axis=0)]


def same_up_to_column_signs(a: np.ndarray, b: np.ndarray, rtol: float) -> bool:
    """eigenvector signs are arbitrary: compare column by column with the better of the two signs;
    the tolerance is relative to the column's largest entry."""
    if a.shape != b.shape:
        return False
    for j in range(a.shape[1]):
        scale = np.abs(b[:, j]).max()
        err = min(np.abs(a[:, j] - b[:, j]).max(), np.abs(a[:, j] + b[:, j]).max())
        if not err <= rtol * scale:
            return False
    return True


def mds_reference_arpack(delta: np.ndarray, dim: int) -> np.ndarray:
    """The reference's own algorithm for timing (bench.py's cpu_baseline of the mds line): the
    statements of kameris/job_steps/mds.py:10-35 with scipy's ARPACK eigsh, as the reference runs them."""
    import scipy.sparse.linalg as linalg
    delta = delta.astype(float)
    n = delta.shape[0]
    deltasq = delta ** 2
    deltatotals = np.sum(deltasq, axis=0) / n
    total = np.sum(deltatotals) / n
    b = deltasq
    b -= deltatotals
    b = b.transpose()
    b -= deltatotals
    b = b.transpose()
    b += total
    b *= -0.5
    vals, vecs = linalg.eigsh(b, k=dim)
    with np.errstate(invalid="ignore"):
        pts = np.fliplr(np.dot(vecs, np.diag(np.sqrt(vals))))
    return pts[:, ~np.isnan(pts).any(axis=0)]
