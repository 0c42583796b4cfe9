"""Drop-in for kameris/job_steps/backend.py: same function names, option keys and error behaviour,
with the two executables replaced by libkameris_b200.so (no subprocess, no CPU fallback).

    run_backend_kmers(options, exp_options)   backend.py:34-56
        options: fasta_output_dir, output_file, k, bits_per_element (16|32),
                 mode ('counts'|'frequencies'), disable_avx (ignored)
    run_backend_dists(options, exp_options)   backend.py:59-66
        options: input_file, output_prefix, distances (list), disable_avx (ignored)
        distances: 'manhat', 'info' (reference) and 'euclid', 'cosine', 'pearson' (new).

Outputs: `.mm-repr` (uint16/uint32 counts, or float64 frequencies shaped 1 x 4^k exactly like the
numpy rewrite of backend.py:45-56) and `<prefix>-<name>.mm-dist`.  Failures raise
subprocess.CalledProcessError like _command.py:17-18 so callers see the same exception type.
"""
from __future__ import absolute_import, division, unicode_literals

import ctypes
import logging
import os
import subprocess
import time

import numpy as np

from .. import _lib, formats

REFERENCE_METRICS = ('manhat', 'info')
GRAM_METRICS = ('euclid', 'cosine', 'pearson')


def _log():
    return logging.getLogger('kameris')


def list_input_files(directory):
    """Every entry of the directory, full paths, byte-wise sorted (generation_cgr @0x1000017c9-
    0x100001b20: directory_iterator + stable_sort(std::less<std::string>), no regular-file filter)."""
    paths = [os.path.join(directory, n) for n in os.listdir(directory)]
    return sorted(paths, key=lambda p: os.fsencode(p))


def _read_record(path):
    lib = _lib.load()
    with open(path, 'rb') as f:
        data = f.read()
    out = ctypes.create_string_buffer(max(len(data), 1))
    n = lib.kam_fasta_record0(data, len(data), out)     # ctypes releases the GIL during the call
    if n < 0:
        raise ValueError('no FASTA record in "%s"' % path)
    return out.raw[:n]


def read_records(paths, threads=None):
    """record[0] of every file (A.1), in the order of `paths`.  A file without a record raises (the
    reference reads begin() of an empty vector there).  Files are read by a small thread pool: both
    the file reads and the record extraction run outside the GIL."""
    if len(paths) < 8:
        return [_read_record(p) for p in paths]
    from concurrent.futures import ThreadPoolExecutor
    workers = threads or min(32, (os.cpu_count() or 4))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(_read_record, paths, chunksize=64))


def kmers_from_records(records, k, bits_per_element, mode):
    """Host records -> (n, 4^k) ndarray through kam_kmers_host: counts (uint16/uint32) or float64
    frequencies (bit-exact cgr/cgr.sum())."""
    import torch
    lib = _lib.load()
    if not isinstance(k, int) or k < 1 or 2 * k > 64:
        raise ValueError('k is too large to fit in the index')
    if bits_per_element not in (16, 32):
        raise ValueError('bits per element must be 16 or 32')
    n = len(records)
    lens = np.fromiter((len(r) for r in records), dtype=np.uint64, count=n)
    start = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:])
    chars = np.frombuffer(b''.join(records), dtype=np.uint8)
    if mode == 'frequencies':
        kind, dt = _lib.KAM_OUT_FREQ_F64, torch.float64
    else:
        kind, dt = _lib.KAM_OUT_COUNTS, (torch.uint16 if bits_per_element == 16 else torch.uint32)
    out = torch.empty((n, 4 ** k), dtype=dt)
    try:
        out = out.pin_memory()                    # direct DMA target; pageable also works (staged)
    except RuntimeError:
        pass
    _lib.check(lib.kam_kmers_host(chars.ctypes.data if chars.size else None, start.ctypes.data, n, k,
                                  bits_per_element, kind, out.data_ptr()))
    np_dt = {torch.float64: np.float64, torch.uint16: np.uint16, torch.uint32: np.uint32}[dt]
    return np.frombuffer((ctypes.c_char * (out.numel() * out.element_size())).from_address(out.data_ptr()),
                         dtype=np_dt).reshape(n, 4 ** k), out


CHUNK_BYTES = 256 << 20        # pinned staging per buffer; two buffers alternate


def kmers_to_file(records, k, bits_per_element, mode, output_file):
    """records -> `.mm-repr`, streamed: the batch is cut into chunks of rows; while chunk c is being
    written to the file from one pinned buffer, chunk c+1 is produced (H2D, kernels, D2H inside
    kam_kmers_host, which releases the GIL) into the other.  No buffer of the whole result exists."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    lib = _lib.load()
    if not isinstance(k, int) or k < 1 or 2 * k > 64:
        raise ValueError('k is too large to fit in the index')
    if bits_per_element not in (16, 32):
        raise ValueError('bits per element must be 16 or 32')
    n, d = len(records), 4 ** k
    lens = np.fromiter((len(r) for r in records), dtype=np.uint64, count=n)
    start = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:])
    chars = np.frombuffer(b''.join(records), dtype=np.uint8)
    if mode == 'frequencies':
        kind, tdt, ndt = _lib.KAM_OUT_FREQ_F64, torch.float64, np.float64
    else:
        kind = _lib.KAM_OUT_COUNTS
        tdt, ndt = (torch.uint16, np.uint16) if bits_per_element == 16 else (torch.uint32, np.uint32)
    rows = max(1, min(n, CHUNK_BYTES // (d * np.dtype(ndt).itemsize)))
    bufs = []
    for _ in range(2 if n > rows else 1):
        t = torch.empty((rows, d), dtype=tdt)
        try:
            t = t.pin_memory()                    # direct DMA target; pageable also works (staged)
        except RuntimeError:
            pass
        bufs.append(t)
    writer = formats.repr_writer(output_file, np.zeros((1, d), dtype=ndt), n, create_file=True)

    def view(t, m):
        raw = (ctypes.c_char * (m * d * t.element_size())).from_address(t.data_ptr())
        return np.frombuffer(raw, dtype=ndt).reshape(m, d)

    try:
        with ThreadPoolExecutor(max_workers=1) as wpool:
            futs = [None, None]
            for c, i0 in enumerate(range(0, n, rows)):
                i1 = min(i0 + rows, n)
                slot = c % len(bufs)
                if futs[slot] is not None:
                    futs[slot].result()           # the write that used this buffer has finished
                sub = (start[i0:i1 + 1] - start[i0]).copy()
                base = chars.ctypes.data + int(start[i0]) if chars.size else None
                _lib.check(lib.kam_kmers_host(base, sub.ctypes.data, i1 - i0, k, bits_per_element, kind,
                                              bufs[slot].data_ptr()))
                futs[slot] = wpool.submit(lambda a=view(bufs[slot], i1 - i0): a.tofile(writer.file))
            for f in futs:
                if f is not None:
                    f.result()
    finally:
        writer.file.close()


def run_backend_kmers(options, exp_options):
    log = _log()
    t0 = time.time()
    try:
        paths = list_input_files(options['fasta_output_dir'])
        log.info('%d files, using B200 backend', len(paths))
        records = read_records(paths)
        kmers_to_file(records, options['k'], options['bits_per_element'], options['mode'], options['output_file'])
    except (_lib.KamError, ValueError, OSError) as e:
        log.error('kmers step failed: %s', e)
        raise subprocess.CalledProcessError(1, 'generation_cgr (kameris_b200)', str(e))
    log.info('kmers: %d vectors of %d bins in %.2f seconds', len(records), 4 ** options['k'], time.time() - t0)


def requested_metrics(names):
    """generation_dists substring-searches argv[3] for each metric name (@0x100000f7e, @0x100000fdf);
    the new names are matched the same way."""
    joined = ','.join(names)
    unknown = [n for n in names if n not in REFERENCE_METRICS + GRAM_METRICS]
    if unknown:
        raise ValueError('unknown distance(s): %s' % ', '.join(unknown))
    return [m for m in REFERENCE_METRICS + GRAM_METRICS if m in joined]


def dists_from_matrix(x, metrics):
    """(n, d) host matrix -> {metric: packed triangle ndarray} through kam_dists_host."""
    lib = _lib.load()
    x = np.ascontiguousarray(x)
    n, d = x.shape
    elem = formats.element_type_of(x.dtype)
    pairs = n * (n - 1) // 2
    bits = {'manhat': _lib.KAM_M_MANHAT, 'info': _lib.KAM_M_INFO, 'euclid': _lib.KAM_M_EUCLID,
            'cosine': _lib.KAM_M_COSINE, 'pearson': _lib.KAM_M_PEARSON}
    out, mask = {}, 0
    for m in metrics:
        out[m] = np.zeros(pairs, dtype=np.uint32 if m == 'manhat' else np.float32)
        mask |= bits[m]
    if any(m in GRAM_METRICS for m in metrics) and elem > 2:
        raise ValueError('euclid/cosine/pearson need integer count vectors (mode: counts)')

    def p(m):
        return out[m].ctypes.data if m in out else None
    _lib.check(lib.kam_dists_host(x.ctypes.data, elem, n, d, mask, p('manhat'), p('info'), p('euclid'), p('cosine'),
                                  p('pearson')))
    return out


def run_backend_dists(options, exp_options):
    log = _log()
    t0 = time.time()
    try:
        metrics = requested_metrics(options['distances'])
        with formats.repr_reader(options['input_file']) as reader:
            x = reader.read_all()                 # raises "Sparse matrices are not supported"
        n = x.shape[0]
        res = dists_from_matrix(x, metrics)
        for m in metrics:
            w = formats.dist_writer('%s-%s.mm-dist' % (options['output_prefix'], m), res[m].dtype, n)
            w.write_triangle(res[m])
            w.close()
    except (_lib.KamError, ValueError, OSError, TypeError) as e:
        log.error('distances step failed: %s', e)
        raise subprocess.CalledProcessError(1, 'generation_dists (kameris_b200)', str(e))
    log.info('distances: %s over %d vectors in %.2f seconds', ','.join(metrics), n, time.time() - t0)
