"""Drop-in for kameris/job_steps/backend.py: same function names, option keys and error behaviour,
with the two executables replaced by libkameris_b200.so (no subprocess, no CPU fallback).

    run_backend_kmers(options, exp_options)   backend.py:34-56
        options: fasta_output_dir, output_file, k, bits_per_element (16|32),
                 mode ('counts'|'frequencies'), disable_avx (ignored)
    run_backend_dists(options, exp_options)   backend.py:59-66
        options: input_file, output_prefix, distances (list), disable_avx (ignored)
        distances: 'manhat', 'info' (reference) and 'euclid', 'cosine', 'pearson' (new; on a `counts`
        file between the count vectors, on a `frequencies` file between the frequency vectors).

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


def read_records_packed(paths, threads=0):
    """record[0] of every file (A.1), in the order of `paths`, as ONE buffer: (chars uint8[total], start
    uint64[n+1]) with record i at chars[start[i]:start[i+1]] -- the layout kam_kmers_host takes.  The files are
    read and the records extracted by the library's host thread pool (kam_fasta_read_files, all cores by
    default), not by Python.  A file without a record raises (the reference reads begin() of an empty vector
    there), and so does an entry that is not a readable regular file."""
    lib = _lib.load()
    n = len(paths)
    start = np.zeros(n + 1, dtype=np.uint64)
    if n == 0:
        return np.zeros(0, dtype=np.uint8), start
    enc = [os.fsencode(p) for p in paths]
    arr = (ctypes.c_char_p * n)(*enc)
    sizes = np.zeros(n, dtype=np.uint64)
    if lib.kam_fasta_file_sizes(arr, n, sizes.ctypes.data, threads) != 0:
        raise OSError(lib.kam_last_error().decode(errors='replace'))
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(sizes, out=offs[1:])
    chars = np.empty(int(offs[-1]) + 16, dtype=np.uint8)
    bad = ctypes.c_uint32(0)
    if lib.kam_fasta_read_files(arr, n, offs.ctypes.data, chars.ctypes.data, start.ctypes.data, threads,
                                ctypes.byref(bad)) != 0:
        msg = lib.kam_last_error().decode(errors='replace')
        raise (ValueError if msg.startswith('no FASTA record') else OSError)(msg)
    return chars[:int(start[-1])], start


def read_records(paths, threads=0):
    """record[0] of every file as a list of bytes objects (see read_records_packed)."""
    chars, start = read_records_packed(paths, threads)
    raw = chars.tobytes()
    return [raw[int(start[i]):int(start[i + 1])] for i in range(len(paths))]


def _pack_records(records):
    """list of bytes -> (chars, start) like read_records_packed"""
    n = len(records)
    lens = np.fromiter((len(r) for r in records), dtype=np.uint64, count=n)
    start = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:])
    return np.frombuffer(b''.join(records), dtype=np.uint8), start


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
    chars, start = _pack_records(records)
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


def kmers_to_file(records, k, bits_per_element, mode, output_file, row_offset=0, total_rows=None, create=True):
    """records (a list of bytes, or the (chars, start) pair of read_records_packed) -> `.mm-repr`, streamed: the batch is cut into chunks of rows; while chunk c is being
    written to the file from one pinned buffer, chunk c+1 is produced (H2D, kernels, D2H inside
    kam_kmers_host, which releases the GIL) into the other.  No buffer of the whole result exists.

    Multi-GPU runs (run_backend_kmers under torch.distributed): `records` is this rank's shard of the
    file list; its rows go to rows [row_offset, row_offset + len(records)) of a file of `total_rows`
    matrices that rank 0 created (create=False: open in place, write nothing but the rows)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    lib = _lib.load()
    if not isinstance(k, int) or k < 1 or 2 * k > 64:
        raise ValueError('k is too large to fit in the index')
    if bits_per_element not in (16, 32):
        raise ValueError('bits per element must be 16 or 32')
    chars, start = records if isinstance(records, tuple) else _pack_records(records)
    n, d = len(start) - 1, 4 ** k
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
    if create:
        writer = formats.repr_writer(output_file, np.zeros((1, d), dtype=ndt), n if total_rows is None else total_rows,
                                     create_file=True)
    else:
        writer = formats.repr_appender(output_file)
    writer.file.flush()
    fd, row_bytes = writer.file.fileno(), d * np.dtype(ndt).itemsize
    base_off = formats.REPR_HEADER + row_offset * row_bytes

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
                # positional, multi-threaded write of the chunk's rows (formats.pwrite_from)
                futs[slot] = wpool.submit(formats.pwrite_from, fd, view(bufs[slot], i1 - i0), base_off + i0 * row_bytes)
            for f in futs:
                if f is not None:
                    f.result()
    finally:
        writer.file.close()


def _world():
    """(torch.distributed module, rank, world) when the step runs under an initialised process group
    with more than one rank (one process per GPU), else (None, 0, 1)."""
    try:
        import torch.distributed as dist
    except ImportError:
        return None, 0, 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _value_dtype(bits_per_element, mode):
    if mode == 'frequencies':
        return np.dtype(np.float64)
    return np.dtype(np.uint16 if bits_per_element == 16 else np.uint32)


def kmers_to_file_sharded(paths, k, bits_per_element, mode, output_file, dist, produce=None):
    """The `kmers` step on all ranks of a process group: the sorted file list is cut into
    contiguous shards with nearly equal numbers of input bytes, every rank turns ITS files into vectors
    and writes them straight into its row range of the one `.mm-repr` (no gather, no data-path
    collective; two barriers order the header before and the close after the row writes).

    `produce(records, row_offset, total_rows)` defaults to the CUDA streaming writer; the CPU tests inject
    a host function to exercise the sharding and the file offsets under gloo."""
    from ..parallel import shard_by_bases
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [os.path.getsize(p) for p in paths]
    b, e = shard_by_bases(sizes, world)[rank]
    n, d = len(paths), 4 ** k
    if rank == 0:
        w = formats.repr_writer(output_file, np.zeros((1, d), dtype=_value_dtype(bits_per_element, mode)), n,
                                create_file=True)
        w.file.truncate(formats.REPR_HEADER + n * d * w.dtype.itemsize)
        w.file.close()
    dist.barrier()
    if produce is None:
        kmers_to_file(read_records_packed(paths[b:e]), k, bits_per_element, mode, output_file, row_offset=b,
                      total_rows=n, create=False)
    else:
        produce(read_records(paths[b:e]), b, n)
    dist.barrier()
    return b, e


def run_backend_kmers(options, exp_options):
    log = _log()
    t0 = time.time()
    try:
        paths = list_input_files(options['fasta_output_dir'])
        log.info('%d files, using B200 backend', len(paths))
        dist, rank, world = _world()
        if dist is not None:
            b, e = kmers_to_file_sharded(paths, options['k'], options['bits_per_element'], options['mode'],
                                         options['output_file'], dist)
            log.info('kmers: rank %d/%d wrote vectors [%d, %d) of %d in %.2f seconds', rank, world, b, e, len(paths),
                     time.time() - t0)
            return
        records = read_records_packed(paths)
        kmers_to_file(records, options['k'], options['bits_per_element'], options['mode'], options['output_file'])
    except (_lib.KamError, ValueError, OSError) as e:
        log.error('kmers step failed: %s', e)
        raise subprocess.CalledProcessError(1, 'generation_cgr (kameris_b200)', str(e))
    log.info('kmers: %d vectors of %d bins in %.2f seconds', len(paths), 4 ** options['k'], time.time() - t0)


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
    if any(m in GRAM_METRICS for m in metrics) and elem == 3:
        raise ValueError('euclid/cosine/pearson are not defined for 64-bit integer vectors')

    def p(m):
        return out[m].ctypes.data if m in out else None
    _lib.check(lib.kam_dists_host(x.ctypes.data, elem, n, d, mask, p('manhat'), p('info'), p('euclid'), p('cosine'),
                                  p('pearson')))
    return out


def dists_to_files_sharded(input_file, output_prefix, metrics, dist, device=None, compute=None):
    """The `distances` step on all ranks of a process group: every rank reads a contiguous
    block of rows of the `.mm-repr`, the rows are all-gathered (the one exchange step), rank r computes
    the block of triangle rows that holds 1/world of the pairs -- a CONTIGUOUS byte range of every
    `.mm-dist` -- and writes it in place into the files rank 0 created.  `compute` as in
    parallel.distances_sharded (CPU tests)."""
    import torch
    from .. import parallel
    rank, world = dist.get_rank(), dist.get_world_size()
    with formats.repr_reader(input_file) as reader:
        n = reader.count
        b, e = rank * n // world, (rank + 1) * n // world
        x = reader.read_rows(b, e)
    if any(m in GRAM_METRICS for m in metrics) and formats.element_type_of(x.dtype) == 3:
        raise ValueError('euclid/cosine/pearson are not defined for 64-bit integer vectors')
    t = torch.from_numpy(x)
    if device is not None:
        t = t.to(device)
    slices, (r0, r1), n_all = parallel.distances_sharded(t, metrics, compute=compute, euclid_on_freq=False)
    assert n_all == n
    off0 = parallel.tri_row_offset(n, r0)
    for m in metrics:
        path = '%s-%s.mm-dist' % (output_prefix, m)
        ndt = np.dtype(np.uint32 if m == 'manhat' else np.float32)
        if rank == 0:
            w = formats.dist_writer(path, ndt, n)
            w.file.truncate(formats.DIST_HEADER + (n * (n - 1) // 2) * 4)
            w.close()
        dist.barrier()
        part = slices[m].cpu().numpy()
        part = part.view(np.uint32) if m == 'manhat' else part
        if part.size:
            with open(path, 'r+b') as f:
                formats.pwrite_from(f.fileno(), np.ascontiguousarray(part, dtype=ndt), formats.DIST_HEADER + off0 * 4)
        dist.barrier()
    return n, (r0, r1)


def run_backend_dists(options, exp_options):
    log = _log()
    t0 = time.time()
    try:
        metrics = requested_metrics(options['distances'])
        dist, rank, world = _world()
        if dist is not None:
            import torch
            n, rows = dists_to_files_sharded(options['input_file'], options['output_prefix'], metrics, dist,
                                             device=torch.device('cuda', torch.cuda.current_device()))
            log.info('distances: rank %d/%d wrote triangle rows [%d, %d) of %d (%s) in %.2f seconds', rank, world,
                     rows[0], rows[1], n, ','.join(metrics), time.time() - t0)
            return
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
