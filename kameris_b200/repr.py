"""Device-level `kmers` path: FASTA records -> planes -> 4^k CGR count / frequency vectors.

Host mirror of what generation_cgr computes per file with the work done by
the CUDA kernels of kameris_b200/csrc/{pack,kmer_count}.cu through the C ABI.  torch is used
only to own device memory and streams.
"""
import numpy as np
import torch

from . import _lib
from ._lib import KAM_OUT_COUNTS, KAM_OUT_FREQ_F32, KAM_OUT_FREQ_F64, check

_OUT_KINDS = {"counts": KAM_OUT_COUNTS, "freq32": KAM_OUT_FREQ_F32, "freq64": KAM_OUT_FREQ_F64}


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class PackedBatch:
    """A batch of sequences in the planes layout (include/kameris_b200.h)."""

    def __init__(self, planes, base_start, head_mask, n_seqs, total_chars):
        self.planes, self.base_start, self.head_mask = planes, base_start, head_mask
        self.n_seqs, self.total_chars = n_seqs, total_chars

    @property
    def device(self):
        return self.planes.device


def concat_records(records):
    """list of bytes -> (uint8 array of all characters, uint64 offsets[n+1])."""
    lens = np.fromiter((len(r) for r in records), dtype=np.uint64, count=len(records))
    start = np.zeros(len(records) + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:])
    chars = np.frombuffer(b"".join(records), dtype=np.uint8)
    return chars, start


def pack_chars(chars_dev, char_start_dev, n_seqs, total_chars):
    """kam_pack_ascii on device tensors (uint8 chars, int64/uint64 offsets)."""
    lib = _lib.load()
    dev = chars_dev.device
    planes = torch.empty(lib.kam_planes_words(total_chars), dtype=torch.int64, device=dev)
    base_start = torch.empty(n_seqs + 1, dtype=torch.int64, device=dev)
    head_mask = torch.empty(max(n_seqs, 1), dtype=torch.int32, device=dev)
    ws_bytes = lib.kam_pack_workspace_bytes(total_chars, n_seqs)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.kam_pack_ascii(_ptr(chars_dev), _ptr(char_start_dev), n_seqs, total_chars, _ptr(planes),
                                 _ptr(base_start), _ptr(head_mask), _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return PackedBatch(planes, base_start, head_mask, n_seqs, total_chars)


def pack_records(records, device="cuda:0"):
    """Host records (list of bytes, record[0] of each file) -> PackedBatch on `device`."""
    chars, start = concat_records(records)
    dev = torch.device(device)
    # pad so the 16-byte vector loads of the pack kernel stay inside the allocation
    chars_t = torch.zeros(max(len(chars), 1) + 16, dtype=torch.uint8)
    chars_t[:len(chars)] = torch.from_numpy(chars.copy())
    chars_dev = chars_t.to(dev)
    start_dev = torch.from_numpy(start.astype(np.int64)).to(dev)
    return pack_chars(chars_dev, start_dev, len(records), int(start[-1]))


def kmer_workspace(n_seqs, k, device):
    lib = _lib.load()
    nbytes = lib.kam_kmer_workspace_bytes(n_seqs, k)
    if nbytes == 0:
        raise ValueError("k is too large to fit in the index" if k > 32 else "unsupported k=%d" % k)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def kmer_vectors(batch, k, bits=16, out="counts", out_tensor=None, workspace=None):
    """CGR vectors of every sequence of `batch`: tensor (n_seqs, 4^k).

    out: 'counts' (uint16/uint32 per `bits`, wrapping like the reference), 'freq32', 'freq64'
    (count / sum; freq64 is bit-exact numpy true division, backend.py:49).
    """
    if k < 1 or 2 * k > 64:
        raise ValueError("k is too large to fit in the index")
    if bits not in (16, 32):
        raise ValueError("bits_per_element must be 16 or 32")
    lib = _lib.load()
    dev = batch.device
    d = 4 ** k
    dtype = {"counts": torch.uint16 if bits == 16 else torch.uint32, "freq32": torch.float32,
             "freq64": torch.float64}[out]
    if out_tensor is None:
        out_tensor = torch.empty((batch.n_seqs, d), dtype=dtype, device=dev)
    assert out_tensor.dtype == dtype and out_tensor.is_contiguous()
    ws = workspace if workspace is not None else kmer_workspace(batch.n_seqs, k, dev)
    with torch.cuda.device(dev):
        check(lib.kam_kmer_count(_ptr(batch.planes), _ptr(batch.base_start), _ptr(batch.head_mask), batch.n_seqs,
                                 batch.total_chars, k, bits, _OUT_KINDS[out], _ptr(out_tensor), out_tensor.stride(0) if batch.n_seqs else d,
                                 _ptr(ws), ws.numel(), _stream_ptr(dev)))
    return out_tensor


def cgr_pool(batch, counts_k1, k):
    """order-(k+1) count vectors of `batch` -> order-k count vectors (2x2 pooling + fix-up, A.3)."""
    lib = _lib.load()
    bits = {torch.uint16: 16, torch.uint32: 32}[counts_k1.dtype]
    assert counts_k1.shape == (batch.n_seqs, 4 ** (k + 1)) and counts_k1.is_contiguous()
    out = torch.empty((batch.n_seqs, 4 ** k), dtype=counts_k1.dtype, device=counts_k1.device)
    with torch.cuda.device(out.device):
        check(lib.kam_cgr_pool(_ptr(batch.planes), _ptr(batch.base_start), _ptr(batch.head_mask), batch.n_seqs, k, bits,
                               _ptr(counts_k1), counts_k1.stride(0) if batch.n_seqs else 4 ** (k + 1), _ptr(out),
                               4 ** k, _stream_ptr(out.device)))
    return out


def normalize(counts, out="freq64"):
    """count matrix on the device -> frequencies (each row divided by its sum, backend.py:49)."""
    lib = _lib.load()
    bits = {torch.uint16: 16, torch.uint32: 32}[counts.dtype]
    n, d = counts.shape
    res = torch.empty((n, d), dtype=torch.float64 if out == "freq64" else torch.float32, device=counts.device)
    with torch.cuda.device(counts.device):
        check(lib.kam_normalize(_ptr(counts), bits, n, d, counts.stride(0) if n else d, _OUT_KINDS[out], _ptr(res), d,
                                _stream_ptr(counts.device)))
    return res


def kmer_sweep(batch, k_lo, k_hi, bits=16, out="counts", out_tensors=None, workspace=None):
    """CGR vectors of every order k_lo..k_hi: dict k -> tensor (n_seqs, 4^k); one fused kernel for
    8 <= k_hi <= 10 (scan once, 2x2 pooling for the lower orders), see kam_kmer_sweep."""
    import ctypes
    lib = _lib.load()
    dev = batch.device
    dtype = {"counts": torch.uint16 if bits == 16 else torch.uint32, "freq32": torch.float32,
             "freq64": torch.float64}[out]
    if out_tensors is None:
        out_tensors = {k: torch.empty((batch.n_seqs, 4 ** k), dtype=dtype, device=dev) for k in range(k_lo, k_hi + 1)}
    ptrs = (ctypes.c_void_p * (k_hi - k_lo + 1))(*[out_tensors[k].data_ptr() for k in range(k_lo, k_hi + 1)])
    ws = workspace if workspace is not None else kmer_workspace(batch.n_seqs, k_hi, dev)
    with torch.cuda.device(dev):
        check(lib.kam_kmer_sweep(_ptr(batch.planes), _ptr(batch.base_start), _ptr(batch.head_mask), batch.n_seqs,
                                 batch.total_chars, k_lo, k_hi, bits, _OUT_KINDS[out], ptrs, _ptr(ws), ws.numel(),
                                 _stream_ptr(dev)))
    return out_tensors


def kmer_sparse(batch, k, char_start_dev, workspace=None):
    """Sparse form of kmer_vectors (kam_kmer_count_sparse): (entries int32[total_chars], nnz int32[n_seqs]);
    vector s is entries[char_start[s] : char_start[s] + nnz[s]], each (bin << 8 | count), ascending by bin.
    nnz == -1 (KAM_SPARSE_DECLINED as int32): the sequence needs kmer_vectors."""
    lib = _lib.load()
    dev = batch.device
    entries = torch.empty(max(batch.total_chars, 1), dtype=torch.int32, device=dev)
    nnz = torch.empty(max(batch.n_seqs, 1), dtype=torch.int32, device=dev)
    ws = workspace if workspace is not None else kmer_workspace(batch.n_seqs, k, dev)
    with torch.cuda.device(dev):
        check(lib.kam_kmer_count_sparse(_ptr(batch.planes), _ptr(batch.base_start), _ptr(batch.head_mask),
                                        batch.n_seqs, k, _ptr(char_start_dev), _ptr(entries), _ptr(nnz), _ptr(ws),
                                        ws.numel(), _stream_ptr(dev)))
    return entries, nnz[:batch.n_seqs]


def sparse_expand(entries, entry_start, nnz, k, bits=16, out="counts"):
    """Host expansion of kmer_sparse lists (numpy arrays) into a dense (n, 4^k) ndarray (kam_sparse_expand)."""
    lib = _lib.load()
    n = len(nnz)
    dt = {"counts": np.uint16 if bits == 16 else np.uint32, "freq32": np.float32, "freq64": np.float64}[out]
    res = np.empty((n, 4 ** k), dtype=dt)
    entries = np.ascontiguousarray(entries, dtype=np.uint32)
    entry_start = np.ascontiguousarray(entry_start, dtype=np.uint64)
    nnz = np.ascontiguousarray(nnz, dtype=np.uint32)
    check(lib.kam_sparse_expand(entries.ctypes.data if entries.size else None, entry_start.ctypes.data,
                                nnz.ctypes.data, n, k, bits, _OUT_KINDS[out], res.ctypes.data))
    return res
