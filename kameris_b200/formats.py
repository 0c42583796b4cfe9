"""`.mm-repr` / `.mm-dist` readers and writers, byte-exact to the layouts the reference uses.

Re-provides the parts of the external `kameris-formats` package that kameris calls
(kameris/job_steps/backend.py:45-56, job_steps/classify.py:219-226, job_steps/mds.py:39,
subcommands/classify.py:51-54) and the C++ twins inside the executables (mmg::repr_writer
generation_cgr @0x100009fc0, mmg::repr_reader generation_dists @0x100019c10, mmg::dist_writer
@0x100019720).  Layouts: SURVEY.md A.4 / A.5 (little-endian, no padding).
"""
import os

import numpy as np

ELEMENT_TYPES = [np.dtype("<u1"), np.dtype("<u2"), np.dtype("<u4"), np.dtype("<u8"), np.dtype("<f4"),
                 np.dtype("<f8")]           # element_type codes 0..5
REPR_MAGIC = b"MMREPR\0"
DIST_MAGIC = b"MMDIST\0"
REPR_HEADER = 34
DIST_HEADER = 16


# ---- bulk file I/O ------------------------------------------------------------------------------------------
# numpy's fromfile / tofile move ~1 GB/s through one thread (and fault in a fresh buffer on the way); the
# matrices here are gigabytes.  Positional reads / writes of disjoint slices from a few threads (the system
# calls release the GIL) reach the page cache's bandwidth instead (20 GB/s for a 2 GB read on 8 cores).
PARALLEL_IO_MIN_BYTES = 32 << 20
PARALLEL_IO_THREADS = 8


def _io_slices(nbytes, threads):
    chunk = -(-nbytes // threads)
    chunk = -(-chunk // 4096) * 4096                      # page-aligned pieces
    return [(o, min(chunk, nbytes - o)) for o in range(0, nbytes, chunk)]


def pread_into(fd, arr, offset, threads=None):
    """fill the C-contiguous ndarray `arr` from file descriptor `fd` at byte `offset`; returns bytes read
    (short only at end of file)."""
    nbytes = arr.nbytes
    if nbytes == 0:
        return 0
    mv = memoryview(arr.reshape(-1).view(np.uint8))

    def one(o, ln):
        got = 0
        while got < ln:
            r = os.preadv(fd, [mv[o + got:o + ln]], offset + o + got)
            if r <= 0:
                break
            got += r
        return got
    t = min(threads or PARALLEL_IO_THREADS, os.cpu_count() or 1)
    if nbytes < PARALLEL_IO_MIN_BYTES or t < 2:
        return one(0, nbytes)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=t) as pool:
        return sum(pool.map(lambda s: one(*s), _io_slices(nbytes, t)))


def pwrite_from(fd, arr, offset, threads=None):
    """write the C-contiguous ndarray `arr` to file descriptor `fd` at byte `offset` (the file grows as needed)."""
    nbytes = arr.nbytes
    if nbytes == 0:
        return
    mv = memoryview(arr.reshape(-1).view(np.uint8))

    def one(o, ln):
        done = 0
        while done < ln:
            done += os.pwritev(fd, [mv[o + done:o + ln]], offset + o + done)
    t = min(threads or PARALLEL_IO_THREADS, os.cpu_count() or 1)
    if nbytes < PARALLEL_IO_MIN_BYTES or t < 2:
        one(0, nbytes)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=t) as pool:
        list(pool.map(lambda s: one(*s), _io_slices(nbytes, t)))


def element_type_of(dtype):
    dt = np.dtype(dtype).newbyteorder("<") if np.dtype(dtype).byteorder == ">" else np.dtype(dtype)
    for i, e in enumerate(ELEMENT_TYPES):
        if dt == e:
            return i
    raise TypeError("unsupported element type %s" % dtype)


class repr_reader(object):
    """Reads a dense `.mm-repr`.  Attributes: count, rows, cols, is_sparse, value_type, dtype, file."""

    def __init__(self, filename):
        self.file = open(filename, "rb")
        head = self.file.read(REPR_HEADER)
        if len(head) < REPR_HEADER or head[:7] != REPR_MAGIC:
            self.file.close()
            raise ValueError("not an .mm-repr file: %s" % filename)
        self.is_sparse = bool(head[7])
        self.key_type = head[8]
        self.value_type = head[9]
        self.count, self.rows, self.cols = (int(v) for v in np.frombuffer(head, dtype="<u8", count=3, offset=10))
        if self.value_type > 5:
            raise ValueError("bad value type %d" % self.value_type)
        self.dtype = ELEMENT_TYPES[self.value_type]

    def _require_dense(self):
        if self.is_sparse:
            raise ValueError("Sparse matrices are not supported")       # generation_dists @0x100027e78

    def read_matrix(self, i, flatten=False):
        self._require_dense()
        if not 0 <= i < self.count:
            raise IndexError(i)
        n = self.rows * self.cols
        self.file.seek(REPR_HEADER + i * n * self.dtype.itemsize)
        m = np.fromfile(self.file, dtype=self.dtype, count=n)
        if m.size != n:
            raise ValueError("truncated .mm-repr file")
        return m if flatten else m.reshape(self.rows, self.cols)

    def read_all(self):
        """All matrices as one (count, rows*cols) array (what generation_dists loads into RAM)."""
        return self.read_rows(0, self.count)

    def read_rows(self, begin, end):
        """Matrices [begin, end) as one (end-begin, rows*cols) array (a rank's shard in multi-GPU runs)."""
        self._require_dense()
        if not 0 <= begin <= end <= self.count:
            raise IndexError((begin, end))
        n = self.rows * self.cols
        m = np.empty((end - begin, n), dtype=self.dtype)
        if pread_into(self.file.fileno(), m, REPR_HEADER + begin * n * self.dtype.itemsize) != m.nbytes:
            raise ValueError("truncated .mm-repr file")
        return m

    def close(self):
        self.file.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class repr_writer(object):
    """Writes a dense `.mm-repr`; the header is taken from `example_matrix` (shape, dtype) and
    `count` like kameris_formats.repr_writer(path, cgrs[0], len(cgrs), create_file=True)."""

    def __init__(self, filename, example_matrix, count, create_file=True):
        ex = np.asarray(example_matrix)
        if ex.ndim == 1:
            ex = ex.reshape(1, -1)
        self.rows, self.cols = ex.shape
        self.value_type = element_type_of(ex.dtype)
        self.dtype = ELEMENT_TYPES[self.value_type]
        self.count = int(count)
        self.written = 0
        self.file = open(filename, "wb" if create_file else "r+b")
        head = REPR_MAGIC + bytes([0, 3, self.value_type]) + \
            np.array([self.count, self.rows, self.cols], dtype="<u8").tobytes()
        assert len(head) == REPR_HEADER
        self.file.write(head)
        self.file.flush()                                  # positional writes (pwrite_from) may follow

    def write_matrix(self, m):
        m = np.ascontiguousarray(m, dtype=self.dtype)
        if m.size != self.rows * self.cols:
            raise ValueError("matrix shape does not match the header")
        m.tofile(self.file)
        self.written += 1

    def write_all(self, mats):
        """(count, rows*cols) array in one write."""
        mats = np.ascontiguousarray(mats, dtype=self.dtype)
        if mats.size != self.count * self.rows * self.cols:
            raise ValueError("matrix block does not match the header")
        self.file.flush()
        pwrite_from(self.file.fileno(), mats, REPR_HEADER)
        self.file.seek(REPR_HEADER + mats.nbytes)
        self.written = self.count

    def close(self):
        self.file.close()


class repr_appender(object):
    """An existing `.mm-repr` opened in place (no header rewrite, no truncation): ranks of a multi-GPU
    run write their own rows into the file rank 0 created."""

    def __init__(self, filename):
        r = repr_reader(filename)
        self.count, self.rows, self.cols, self.dtype, self.value_type = r.count, r.rows, r.cols, r.dtype, r.value_type
        r.close()
        self.file = open(filename, "r+b")

    def close(self):
        self.file.close()


def triangle_index(n, i, j):
    """position of (i<j) in the packed strict upper triangle (generation_dists @0x10001bf7e)."""
    return n * i - i * (i + 3) // 2 + j - 1


class dist_writer(object):
    """Writes a `.mm-dist`: header then the strict upper triangle row by row."""

    def __init__(self, filename, dtype, size):
        self.value_type = element_type_of(dtype)
        self.dtype = ELEMENT_TYPES[self.value_type]
        self.size = int(size)
        self.file = open(filename, "wb")
        self.file.write(DIST_MAGIC + bytes([self.value_type]) + np.array([self.size], dtype="<u8").tobytes())

    def write_triangle(self, tri):
        tri = np.ascontiguousarray(tri, dtype=self.dtype)
        if tri.size != self.size * (self.size - 1) // 2:
            raise ValueError("triangle length does not match the header")
        self.file.flush()
        pwrite_from(self.file.fileno(), tri, DIST_HEADER)
        self.file.seek(DIST_HEADER + tri.nbytes)

    def write_next_row(self, row):
        np.ascontiguousarray(row, dtype=self.dtype).tofile(self.file)

    def close(self):
        self.file.close()


class dist_reader(object):
    @staticmethod
    def read_header(filename):
        with open(filename, "rb") as f:
            head = f.read(DIST_HEADER)
        if len(head) < DIST_HEADER or head[:7] != DIST_MAGIC or head[7] > 5:
            raise ValueError("not an .mm-dist file: %s" % filename)
        return ELEMENT_TYPES[head[7]], int(np.frombuffer(head, dtype="<u8", count=1, offset=8)[0])

    @staticmethod
    def read_triangle(filename):
        dt, n = dist_reader.read_header(filename)
        tri = np.empty(n * (n - 1) // 2, dtype=dt)
        with open(filename, "rb") as f:
            if pread_into(f.fileno(), tri, DIST_HEADER) != tri.nbytes:
                raise ValueError("truncated .mm-dist file")
        return tri, n

    @staticmethod
    def read_matrix(filename):
        """Full symmetric N x N ndarray (diagonal 0), as job_steps/classify.py:219 and mds.py:39 expect."""
        tri, n = dist_reader.read_triangle(filename)
        m = np.zeros((n, n), dtype=tri.dtype)
        iu = np.triu_indices(n, 1)
        m[iu] = tri
        m.T[iu] = tri
        return m
