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

# ---- hand-over between the steps of one job ------------------------------------------------------------------
# run-job runs the steps of an experiment back to back (kameris/subcommands/run_job.py:175-180): `kmers` writes a
# .mm-repr, `distances` reads it again.  The count matrix the kmers step computed stays on the device (when it
# fits RESIDENT_MAX_BYTES), remembered under the output file's path, size and mtime; a distances step whose
# input file still matches takes it from there -- no re-read of the file, no second upload.  The files written
# are the same either way.
RESIDENT_MAX_BYTES = int(float(os.environ.get('KAM_RESIDENT_GB', '16')) * (1 << 30))
_resident = {}
resident_hits = 0


def set_resident_limit(nbytes):
    global RESIDENT_MAX_BYTES
    RESIDENT_MAX_BYTES = int(nbytes)
    if nbytes <= 0:
        release_resident()


def release_resident():
    _resident.clear()


def _remember(path, counts, mode):
    st = os.stat(path)
    _resident.clear()                                   # one matrix at a time
    _resident[os.path.realpath(path)] = {'size': st.st_size, 'mtime_ns': st.st_mtime_ns, 'counts': counts, 'mode': mode}


def _recall(path):
    e = _resident.get(os.path.realpath(path))
    if e is None:
        return None
    try:
        st = os.stat(path)
    except OSError:
        st = None
    if st is None or st.st_size != e['size'] or st.st_mtime_ns != e['mtime_ns']:
        _resident.clear()                               # the file changed since: forget the matrix
        return None
    return e


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
    keep = None
    cdt = torch.uint16 if bits_per_element == 16 else torch.uint32
    if create and row_offset == 0 and total_rows is None and n > 1 and torch.cuda.is_available() and \
            n * d * (bits_per_element // 8) <= RESIDENT_MAX_BYTES:
        try:
            keep = torch.empty((n, d), dtype=cdt, device=torch.device('cuda', torch.cuda.current_device()))
        except RuntimeError:                              # no room on the device: the hand-over is optional
            keep = None
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
                dk = keep.data_ptr() + i0 * d * (bits_per_element // 8) if keep is not None else None
                _lib.check(lib.kam_kmers_host_keep(base, sub.ctypes.data, i1 - i0, k, bits_per_element, kind,
                                                   bufs[slot].data_ptr(), dk, d))
                # positional, multi-threaded write of the chunk's rows (formats.pwrite_from)
                futs[slot] = wpool.submit(formats.pwrite_from, fd, view(bufs[slot], i1 - i0), base_off + i0 * row_bytes)
            for f in futs:
                if f is not None:
                    f.result()
    finally:
        writer.file.close()
    if keep is not None:
        _remember(output_file, keep, mode)


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


def _agree(dist, err, what):
    """Every rank learns here whether ANY rank failed, instead of a bare barrier: a rank that raised between
    two collectives would leave the others waiting until the NCCL timeout.  Re-raises the local error, or an
    OSError on the ranks that were fine."""
    import torch
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        if err is not None:
            raise err
        raise OSError('%s failed on another rank' % what)


def _unlink_quietly(paths):
    for p in paths:
        try:
            os.unlink(p)
        except OSError:
            pass


def kmers_to_file_sharded(paths, k, bits_per_element, mode, output_file, dist, produce=None):
    """The `kmers` step on all ranks of a process group: the sorted file list is cut into
    contiguous shards with nearly equal numbers of input bytes, every rank turns ITS files into vectors
    and writes them straight into its row range of the one `.mm-repr` (no gather, no data-path
    collective; two agreement points order the header before and the close after the row writes, and make
    every rank fail -- and rank 0 remove the partial file -- when any rank failed).

    `produce(records, row_offset, total_rows)` defaults to the CUDA streaming writer; the CPU tests inject
    a host function to exercise the sharding and the file offsets under gloo."""
    from ..parallel import shard_by_bases
    rank, world = dist.get_rank(), dist.get_world_size()
    n, d = len(paths), 4 ** k
    err, b, e = None, 0, 0
    try:
        sizes = [os.path.getsize(p) for p in paths]
        b, e = shard_by_bases(sizes, world)[rank]
        if rank == 0:
            w = formats.repr_writer(output_file, np.zeros((1, d), dtype=_value_dtype(bits_per_element, mode)), n,
                                    create_file=True)
            w.file.truncate(formats.REPR_HEADER + n * d * w.dtype.itemsize)
            w.file.close()
    except Exception as ex:   # noqa: BLE001 -- whatever it is, the other ranks must hear of it
        err = ex
    try:
        _agree(dist, err, 'kmers step (creating %s)' % output_file)
        try:
            if produce is None:
                kmers_to_file(read_records_packed(paths[b:e]), k, bits_per_element, mode, output_file, row_offset=b,
                              total_rows=n, create=False)
            else:
                produce(read_records(paths[b:e]), b, n)
        except Exception as ex:   # noqa: BLE001
            err = ex
        _agree(dist, err, 'kmers step (vectors of a shard)')
    except Exception:
        if rank == 0:
            _unlink_quietly([output_file])
        raise
    return b, e


def _kmers_single(options):
    """the step in one process (what run_backend_kmers does outside a process group)"""
    paths = list_input_files(options['fasta_output_dir'])
    records = read_records_packed(paths)
    kmers_to_file(records, options['k'], options['bits_per_element'], options['mode'], options['output_file'])
    return len(paths)


def run_backend_kmers(options, exp_options):
    log = _log()
    t0 = time.time()
    try:
        dist, rank, world = _world()
        if dist is not None:
            paths = list_input_files(options['fasta_output_dir'])
            log.info('%d files, using B200 backend', len(paths))
            b, e = kmers_to_file_sharded(paths, options['k'], options['bits_per_element'], options['mode'],
                                         options['output_file'], dist)
            log.info('kmers: rank %d/%d wrote vectors [%d, %d) of %d in %.2f seconds', rank, world, b, e, len(paths),
                     time.time() - t0)
            return
        n = _kmers_single(options)
    except (_lib.KamError, ValueError, OSError) as e:
        log.error('kmers step failed: %s', e)
        raise subprocess.CalledProcessError(1, 'generation_cgr (kameris_b200)', str(e))
    log.info('kmers: %d vectors of %d bins in %.2f seconds', n, 4 ** options['k'], time.time() - t0)


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


def _write_ring_blocks(blocks, counts, metrics, paths, n):
    """Dense blocks of parallel.gram_ring -> their places in the packed `.mm-dist` triangles (A.5 order,
    generation_dists @0x10001bf7e).  Row i of a block covers a run of consecutive columns j > i, i.e. ONE
    contiguous segment of the file; a block whose column shard lies before its row shard is the transposed
    case (its pairs are stored under the column's row).  Segments of different blocks are disjoint, so all ranks
    write into the same files without coordination."""
    from concurrent.futures import ThreadPoolExecutor
    from ..parallel import tri_row_offset
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    with ThreadPoolExecutor(max_workers=8) as pool:
        for blk in blocks:
            a, c = blk['row_shard'], blk['col_shard']
            r0, r1 = blk['rows']
            for m in metrics:
                host = blk[m].cpu().numpy()
                if blk['tri']:
                    segs = [(int(offs[a]) + r, int(offs[a]) + r + 1, host[r - r0, r + 1:]) for r in range(r0, r1)]
                elif a < c:
                    segs = [(int(offs[a]) + r, int(offs[c]), host[r - r0]) for r in range(r0, r1)]
                else:            # pairs (i in shard a, j in shard c < a) live in row j of the triangle
                    ht = np.ascontiguousarray(host.T)
                    segs = [(int(offs[c]) + j, int(offs[a]) + r0, ht[j]) for j in range(ht.shape[0])]
                with open(paths[m], 'r+b') as f:
                    fd = f.fileno()

                    def put(seg):
                        i, j0, vals = seg
                        if vals.size:
                            os.pwrite(fd, np.ascontiguousarray(vals).data,        # float32, or manhat's uint32 bits
                                      formats.DIST_HEADER + 4 * (tri_row_offset(n, i) + j0 - i - 1))
                    list(pool.map(put, segs))


def dists_to_files_sharded(input_file, output_prefix, metrics, dist, device=None, compute=None):
    """The `distances` step on all ranks of a process group: every rank reads a contiguous block
    of rows of the `.mm-repr`.
      * euclid / cosine / pearson: parallel.gram_ring -- block-cyclic pair ownership, shards of tensor-core
        operand rows travel around the ring while the block before is computed; every rank writes the rows of
        its dense blocks into their (disjoint) segments of the `.mm-dist` files;
      * manhat / info (and Gram metrics outside the tensor-exact envelope, or with an injected `compute`): the
        rows are all-gathered, rank r computes the triangle rows that hold 1/world of the pairs -- a CONTIGUOUS
        byte range of every `.mm-dist` -- and writes it in place.
    Rank 0 creates the files; agreement points replace bare barriers so that a failure on any rank fails every
    rank and removes the partial outputs.  `compute` as in parallel.distances_sharded (CPU tests)."""
    import torch
    from .. import parallel
    rank, world = dist.get_rank(), dist.get_world_size()
    paths = {m: '%s-%s.mm-dist' % (output_prefix, m) for m in metrics}
    err, n, x = None, 0, None
    try:
        with formats.repr_reader(input_file) as reader:
            n = reader.count
            b, e = rank * n // world, (rank + 1) * n // world
            x = reader.read_rows(b, e)
        if any(m in GRAM_METRICS for m in metrics) and formats.element_type_of(x.dtype) == 3:
            raise ValueError('euclid/cosine/pearson are not defined for 64-bit integer vectors')
        if rank == 0:
            for m in metrics:
                w = formats.dist_writer(paths[m], np.dtype(np.uint32 if m == 'manhat' else np.float32), n)
                w.file.truncate(formats.DIST_HEADER + (n * (n - 1) // 2) * 4)
                w.close()
    except Exception as ex:   # noqa: BLE001 -- the other ranks must hear of it
        err = ex
    rows = (0, 0)
    try:
        _agree(dist, err, 'distances step (reading %s / creating the outputs)' % input_file)
        try:
            t = torch.from_numpy(x)
            if device is not None:
                t = t.to(device)
            rest = list(metrics)
            gram = [m for m in metrics if m in GRAM_METRICS]
            if gram and compute is None and t.is_cuda and x.dtype != np.uint64:
                try:
                    blocks, counts = parallel.gram_ring(t, gram, euclid_on_freq=False)
                    _write_ring_blocks(blocks, counts, gram, paths, n)
                    del blocks
                    rest = [m for m in metrics if m not in gram]
                except (NotImplementedError, ValueError):
                    pass          # decided from all-reduced flags: every rank takes the all-gather path below
            # manhat / info ride the same schedule on the tensor cores (thermometer-coded rows) where the counts
            # allow it; the decision comes from all-reduced column maxima, so every rank decides alike
            ref_metrics = [m for m in rest if m in ('manhat', 'info')]
            if ref_metrics and compute is None and t.is_cuda and x.dtype in (np.uint8, np.uint16, np.uint32) \
                    and world > 1 and os.environ.get('KAM_RING_FUSED', '1') == '1':
                for m in ref_metrics:
                    try:
                        blocks, counts = parallel.thermo_ring(t, (m,))
                    except NotImplementedError:
                        continue
                    _write_ring_blocks(blocks, counts, [m], paths, n)
                    del blocks
                    rest = [r for r in rest if r != m]
            if rest:
                slices, rows, n_all = parallel.distances_sharded(t, rest, compute=compute, euclid_on_freq=False)
                assert n_all == n
                off0 = parallel.tri_row_offset(n, rows[0])
                for m in rest:
                    ndt = np.dtype(np.uint32 if m == 'manhat' else np.float32)
                    part = slices[m].cpu().numpy()
                    part = part.view(np.uint32) if m == 'manhat' else part
                    if part.size:
                        with open(paths[m], 'r+b') as f:
                            formats.pwrite_from(f.fileno(), np.ascontiguousarray(part, dtype=ndt),
                                                formats.DIST_HEADER + off0 * 4)
        except Exception as ex:   # noqa: BLE001
            err = ex
        _agree(dist, err, 'distances step (a rank\'s share of the pairs)')
    except Exception:
        if rank == 0:
            _unlink_quietly(paths.values())
        raise
    return n, rows


def dists_from_device_counts(counts, metrics, euclid_on_freq):
    """device count matrix (uint16/uint32) -> {metric: packed triangle ndarray} through kam_dists_device"""
    lib = _lib.load()
    n, d = counts.shape
    pairs = n * (n - 1) // 2
    bits = {'manhat': _lib.KAM_M_MANHAT, 'info': _lib.KAM_M_INFO, 'euclid': _lib.KAM_M_EUCLID,
            'cosine': _lib.KAM_M_COSINE, 'pearson': _lib.KAM_M_PEARSON}
    out, mask = {}, 0
    for m in metrics:
        out[m] = np.zeros(pairs, dtype=np.uint32 if m == 'manhat' else np.float32)
        mask |= bits[m]

    def p(m):
        return out[m].ctypes.data if m in out else None
    import torch
    with torch.cuda.device(counts.device):
        _lib.check(lib.kam_dists_device(counts.data_ptr(), _lib.KAM_U16 if counts.element_size() == 2 else _lib.KAM_U32,
                                        n, d, counts.stride(0), mask, int(euclid_on_freq), p('manhat'), p('info'),
                                        p('euclid'), p('cosine'), p('pearson')))
    return out


def _dists_single(options, metrics):
    """the step in one process (what run_backend_dists does outside a process group)"""
    global resident_hits
    res, n = {}, None
    held = _recall(options['input_file'])
    todo = list(metrics)
    if held is not None:
        # counts file: everything from the resident matrix.  frequencies file: the rows are count / sum, so
        # cosine / pearson / info are those of the counts and euclid is euclid between the frequency vectors;
        # only `manhat` depends on the float rows themselves (it truncates per addition, quirk Q5).
        mine = todo if held['mode'] != 'frequencies' else [m for m in todo if m != 'manhat']
        if mine:
            n = held['counts'].shape[0]
            res.update(dists_from_device_counts(held['counts'], mine, held['mode'] == 'frequencies'))
            todo = [m for m in todo if m not in mine]
            resident_hits += 1
    if todo:
        with formats.repr_reader(options['input_file']) as reader:
            x = reader.read_all()                 # raises "Sparse matrices are not supported"
        n = x.shape[0]
        res.update(dists_from_matrix(x, todo))
    for m in metrics:
        w = formats.dist_writer('%s-%s.mm-dist' % (options['output_prefix'], m), res[m].dtype, n)
        w.write_triangle(res[m])
        w.close()
    return n


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
            log.info('distances: rank %d/%d wrote its share of the pairs of %d vectors (%s) in %.2f seconds', rank,
                     world, n, ','.join(metrics), time.time() - t0)
            return
        n = _dists_single(options, metrics)
    except (_lib.KamError, ValueError, OSError, TypeError) as e:
        log.error('distances step failed: %s', e)
        raise subprocess.CalledProcessError(1, 'generation_dists (kameris_b200)', str(e))
    log.info('distances: %s over %d vectors in %.2f seconds', ','.join(metrics), n, time.time() - t0)
