"""CPU tests: file formats (byte layouts of SURVEY.md A.4/A.5), the C ABI surface, host logic."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from kameris_b200 import formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_repr_layout_and_roundtrip(tmp_path):
    p = str(tmp_path / "a.mm-repr")
    mats = (np.arange(3 * 16, dtype=np.uint16).reshape(3, 16) * 7)
    w = formats.repr_writer(p, np.zeros((1, 16), np.uint16), 3, create_file=True)
    for m in mats:
        w.write_matrix(m)
    w.file.close()
    raw = open(p, "rb").read()
    # A.4: "MMREPR\0", is_sparse, key_type=3, value_type=1 (u16), count, rows, cols (u64 LE), data at 0x22
    assert raw[:7] == b"MMREPR\0" and raw[7:10] == bytes([0, 3, 1])
    assert struct.unpack("<QQQ", raw[10:34]) == (3, 1, 16)
    assert len(raw) == 34 + 3 * 16 * 2 and raw[34:36] == b"\x00\x00" and raw[36:38] == struct.pack("<H", 7)
    r = formats.repr_reader(p)
    assert (r.count, r.rows, r.cols, r.value_type) == (3, 1, 16)[:3] + (1,)
    assert r.read_matrix(1).shape == (1, 16) and np.array_equal(r.read_matrix(2, flatten=True), mats[2])
    assert np.array_equal(r.read_all(), mats)
    r.file.close()


def test_repr_float64_rewrite_like_backend(tmp_path):
    """the frequencies rewrite of backend.py:45-56 through the re-provided API."""
    p = str(tmp_path / "f.mm-repr")
    w = formats.repr_writer(p, np.zeros((1, 4), np.uint16), 2)
    w.write_matrix(np.array([1, 2, 3, 4], np.uint16))
    w.write_matrix(np.array([0, 0, 5, 5], np.uint16))
    w.file.close()
    reader = formats.repr_reader(p)
    cgrs = []
    for i in range(reader.count):
        cgr = reader.read_matrix(i)
        cgrs.append(cgr / cgr.sum())
    reader.file.close()
    writer = formats.repr_writer(p, cgrs[0], len(cgrs), create_file=True)
    for cgr in cgrs:
        writer.write_matrix(cgr)
    writer.file.close()
    r = formats.repr_reader(p)
    assert r.value_type == 5 and (r.rows, r.cols) == (1, 4)
    assert np.array_equal(r.read_matrix(0, flatten=True), np.array([0.1, 0.2, 0.3, 0.4]))
    r.file.close()


def test_sparse_rejected(tmp_path):
    p = str(tmp_path / "s.mm-repr")
    open(p, "wb").write(b"MMREPR\0" + bytes([1, 3, 1]) + struct.pack("<QQQ", 1, 1, 4) + b"\0" * 8)
    with pytest.raises(ValueError, match="Sparse matrices are not supported"):
        formats.repr_reader(p).read_matrix(0)
    with pytest.raises(ValueError):
        open(p, "wb").write(b"garbage")
        formats.repr_reader(p)


def test_dist_layout_and_mirror(tmp_path):
    p = str(tmp_path / "d-manhat.mm-dist")
    n = 5
    tri = np.arange(10, dtype=np.uint32) + 1
    w = formats.dist_writer(p, np.uint32, n)
    w.write_triangle(tri)
    w.close()
    raw = open(p, "rb").read()
    assert raw[:7] == b"MMDIST\0" and raw[7] == 2 and struct.unpack("<Q", raw[8:16])[0] == n   # A.5
    assert len(raw) == 16 + 10 * 4
    m = formats.dist_reader.read_matrix(p)
    assert m.shape == (5, 5) and np.array_equal(m, m.T) and (np.diag(m) == 0).all()
    for i in range(n):
        for j in range(i + 1, n):
            assert m[i, j] == tri[formats.triangle_index(n, i, j)]
    assert formats.triangle_index(n, 0, 1) == 0 and formats.triangle_index(n, 3, 4) == 9
    w = formats.dist_writer(p, np.float32, 3)
    w.write_next_row(np.array([0.5, 0.25], np.float32))
    w.write_next_row(np.array([0.125], np.float32))
    w.write_next_row(np.array([], np.float32))
    w.close()
    t, nn = formats.dist_reader.read_triangle(p)
    assert nn == 3 and t.dtype == np.float32 and t.tolist() == [0.5, 0.25, 0.125]


def test_abi_exports_every_declared_symbol():
    """The library loads without a GPU and exports exactly what include/kameris_b200.h declares."""
    from kameris_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kameris_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kam_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.kam_version() == 100
    assert lib.kam_planes_words(64) == 4 and lib.kam_kmer_workspace_bytes(10, 9) > 4 ** 9 * 4
    assert lib.kam_kmer_workspace_bytes(10, 16) == 0            # k beyond the supported range


def test_fasta_record0_host_matches_oracle():
    from kameris_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    cases = [b">h\nACGT\nAC\n\n>h2\nGG\n", b"  ACG T \r\nNN\r\n", b"\n\n>x\n>y\nTT", b"ACGT", b">a\n\n", b"",
             b"\t\x0b\x0cAC\x0b\n>b", b">h\r\nAC\r\nGT\r\n\r\nTT\r\n"]
    for c in cases:
        out = ctypes.create_string_buffer(max(len(c), 1))
        n = lib.kam_fasta_record0(c, len(c), out)
        try:
            ref = O.fasta_record0(c)
            assert n == len(ref) and out.raw[:n] == ref
        except ValueError:
            assert n == -1


def test_requested_metrics_and_row_ranges():
    from kameris_b200.job_steps import backend
    assert backend.requested_metrics(["manhat"]) == ["manhat"]
    assert backend.requested_metrics(["info", "manhat", "pearson"]) == ["manhat", "info", "pearson"]
    with pytest.raises(ValueError):
        backend.requested_metrics(["chebyshev"])
    from kameris_b200 import dist as D
    for n, parts in [(10, 3), (1000, 8), (7, 8), (100000, 8)]:
        rr = D.balanced_row_ranges(n, parts)
        assert rr[0][0] == 0 and rr[-1][1] == n and all(a[1] == b[0] for a, b in zip(rr, rr[1:]))
        pairs = [sum(n - 1 - i for i in range(a, b)) if n <= 1000 else (b - a) * (2 * n - a - b - 1) // 2
                 for a, b in rr]
        assert sum(pairs) == n * (n - 1) // 2
        if n >= 1000:
            assert max(pairs) < 1.05 * (sum(pairs) / parts) + n


def test_list_input_files_sorted(tmp_path):
    from kameris_b200.job_steps import backend
    for n in ["b.fasta", "A1.fasta", "a.fasta", "B.fasta", "10.fasta", "9.fasta"]:
        (tmp_path / n).write_text(">x\nACGT\n")
    got = [os.path.basename(p) for p in backend.list_input_files(str(tmp_path))]
    assert got == ["10.fasta", "9.fasta", "A1.fasta", "B.fasta", "a.fasta", "b.fasta"]   # byte-wise


def test_abi_argument_validation_without_gpu():
    """Bad arguments are rejected with an error code and a message before any CUDA call."""
    import ctypes
    from kameris_b200 import _lib
    lib = _lib.load()
    p = ctypes.c_void_p(16)          # non-null dummy pointers: validation must stop before they are touched
    assert lib.kam_kmer_count(p, p, p, 1, 0, 33, 16, 0, p, 4, p, 1 << 20, None) == -1
    assert b"k is too large to fit in the index" in lib.kam_last_error()
    assert lib.kam_kmer_count(p, p, p, 1, 0, 16, 16, 0, p, 4 ** 16, p, 1 << 20, None) == -4      # k beyond this build
    assert lib.kam_kmer_count(p, p, p, 1, 0, 4, 8, 0, p, 256, p, 1 << 20, None) == -1
    assert b"count_bits" in lib.kam_last_error()
    assert lib.kam_kmer_count(p, p, p, 1, 0, 4, 16, 0, p, 10, p, 1 << 20, None) == -1             # out_stride < 4^k
    assert lib.kam_kmer_count(p, p, p, 1, 0, 9, 16, 0, p, 4 ** 9, p, 64, None) == -3              # workspace too small
    assert lib.kam_kmers_host(None, None, 1, 4, 16, 0, None) == -1
    assert lib.kam_gram_tri(p, 3, 10, 16, 16, 0, 10, 8, 1, None, p, None, p, 1 << 30, None) == -1  # uint64 input
    assert b"count vectors" in lib.kam_last_error()
    assert lib.kam_gram_tri(p, 1, 10, 16, 16, 0, 10, 8, 1, None, None, None, p, 1 << 30, None) == -1   # missing output
    assert lib.kam_manhattan_tri(p, 9, 10, 16, 16, 0, 10, p, p, 1 << 30, None) == -1
    assert lib.kam_dists_host(p, 1, 10, 16, 0, None, None, None, None, None) == -1                 # no metric asked
    assert lib.kam_gram_dpad(4 ** 9) == 4 ** 9 and lib.kam_gram_dpad(4) == 128
    # the fused exchange and the tensor path of manhat / info (round 2)
    seg = (ctypes.c_uint32 * 4)(0, 1, 0, 1)
    bad = (ctypes.c_uint32 * 4)(2, 1, 0, 1)                       # column tile begin > end
    assert lib.kam_gram_fused(p, p, p, 10, p, p, p, 10, 0, 16, 1, None, 1, None, 1, 8, 0, None, p, None, 10, 0, p, 1 << 20,
                              None) == -1                          # no segments
    assert lib.kam_gram_fused(p, p, p, 10, p, p, p, 10, 0, 16, 1, bad, 1, None, 1, 8, 0, None, p, None, 10, 0, p, 1 << 20,
                              None) == -1
    assert b"segment" in lib.kam_last_error()
    assert lib.kam_gram_fused(p, p, p, 10, p, p, p, 20, 0, 16, 1, seg, 1, None, 1, 8, 0, None, p, None, 10, 0, p, 1 << 20,
                              None) == -1                          # out_ld < n_b
    assert lib.kam_thermo_fused(p, p, p, 10, p, p, p, 10, 16, 128, 1, seg, 1, None, 1, 4, p, 10, p, 1 << 20, None) == -1
    assert b"manhat or info" in lib.kam_last_error()               # metric must be KAM_M_MANHAT or KAM_M_INFO
    assert lib.kam_thermo_fused(p, p, p, 10, p, p, p, 10, 16, 100, 1, seg, 1, None, 1, 1, p, 10, p, 1 << 20, None) == -1
    assert b"kpad" in lib.kam_last_error()                         # rows are multiples of 128 bytes
    assert lib.kam_dist_colmax(p, 4, 10, 16, 16, p, None) == -1   # float rows
    assert lib.kam_dist_encode(p, 1, 10, 16, 16, p, 100, p, p, p, p, 1 << 20, None) == -1
    assert lib.kam_dist_encode(p, 1, 10, 16, 16, p, 128, ctypes.c_void_p(1024), p, p, p, 16, None) == -3    # workspace too small
    assert lib.kam_dist_kpad(0) == 128 and lib.kam_dist_kpad(129) == 256
    assert lib.kam_peer_pull(None, 1, None, None, 0) == -1
    assert lib.kam_manhattan_workspace_bytes_levels(1000, 4096, 9) > lib.kam_manhattan_workspace_bytes(1000, 4096)
    assert lib.kam_manhattan_workspace_bytes_levels(1000, 4096, 2) == lib.kam_manhattan_workspace_bytes(1000, 4096)


def test_native_fasta_ingest_matches_record0_semantics(tmp_path):
    """kam_fasta_read_files (host thread pool, no GPU): record[0] of every file exactly as the per-file A.1
    restatement extracts it -- CRLF, blank lines, several records, header-less files, inner whitespace kept,
    leading / trailing whitespace trimmed, no trailing newline, lower case kept -- contiguous, in path order."""
    from kameris_b200.job_steps import backend
    from oracle import oracle as O
    from tests.helpers import synth_records
    bodies = [b">a desc\nACGT\nACGT\n", b">a\r\nACGTN\r\nNNAC\r\n\r\n>b\r\nTTTT\r\n", b"ACGT\nAC GT\n", b"\n\n>x\n  acgtRY \t\nGG\n>y\nAA",
              b">only\nA", b">h1\n>h2\nCCCC\n", b"  \t\n>z\n\nGATTACA\n\nTTTT\n", b"A" * 100000 + b"\n" + b"C" * 7]
    bodies += [b">r%d\n" % i + b"\n".join(r[j:j + 61] for j in range(0, len(r), 61)) + b"\n"
               for i, r in enumerate(synth_records(300, 700, seed=2, jitter=300, dirty=0.01))]
    d = tmp_path / "fa"
    d.mkdir()
    for i, b in enumerate(bodies):
        (d / ("s%04d.fasta" % i)).write_bytes(b)
    paths = backend.list_input_files(str(d))
    for threads in (0, 1, 3):
        chars, start = backend.read_records_packed(paths, threads)
        assert start[0] == 0 and len(start) == len(paths) + 1 and chars.size == start[-1]
        for i, p in enumerate(paths):
            with open(p, "rb") as f:
                want = O.fasta_record0(f.read())
            assert chars[int(start[i]):int(start[i + 1])].tobytes() == want, (threads, p)
    assert backend.read_records(paths[:5]) == [O.fasta_record0(b) for b in bodies[:5]]
    assert backend.read_records_packed([])[1].tolist() == [0]
    # a file without a record, a missing file and a directory entry fail loudly, naming the first offender
    (d / "s0003a.fasta").write_bytes(b">header only\n\n")
    with pytest.raises(ValueError, match="s0003a.fasta"):
        backend.read_records_packed(backend.list_input_files(str(d)))
    (d / "s0003a.fasta").unlink()
    with pytest.raises(OSError):
        backend.read_records_packed(paths + [str(d / "missing.fasta")])
    (d / "zdir").mkdir()
    with pytest.raises(OSError, match="zdir"):
        backend.read_records_packed(backend.list_input_files(str(d)))


def test_parallel_positional_io(tmp_path, monkeypatch):
    """formats.pread_into / pwrite_from (threaded positional I/O behind read_rows / write_all / write_triangle and
    the kmers step's chunk writes): byte-identical files and arrays whatever the slice boundaries, offsets that
    are not page-aligned, short reads at end of file, empty arrays."""
    monkeypatch.setattr(formats, "PARALLEL_IO_MIN_BYTES", 1 << 12)          # force the threaded path
    rng = np.random.Generator(np.random.PCG64(4))
    mats = rng.integers(0, 65536, size=(37, 4 ** 5), dtype=np.uint16)
    seq, par, chunked = (str(tmp_path / n) for n in ("seq.mm-repr", "par.mm-repr", "chunked.mm-repr"))
    w = formats.repr_writer(seq, mats[:1], len(mats), create_file=True)
    for m in mats:
        w.write_matrix(m)
    w.close()
    w = formats.repr_writer(par, mats[:1], len(mats), create_file=True)
    w.write_all(mats)
    w.close()
    # the kmers step's pattern: header, then chunks of rows written at their offsets, out of order
    w = formats.repr_writer(chunked, mats[:1], len(mats), create_file=True)
    row_bytes = mats.shape[1] * 2
    for i0, i1 in ((20, 37), (0, 7), (7, 20)):
        formats.pwrite_from(w.file.fileno(), np.ascontiguousarray(mats[i0:i1]), formats.REPR_HEADER + i0 * row_bytes,
                            threads=3)
    w.close()
    ref = open(seq, "rb").read()
    assert open(par, "rb").read() == ref and open(chunked, "rb").read() == ref
    with formats.repr_reader(seq) as r:
        assert np.array_equal(r.read_all(), mats)
        assert np.array_equal(r.read_rows(5, 23), mats[5:23])
        assert r.read_rows(9, 9).shape == (0, 4 ** 5)
        assert np.array_equal(r.read_matrix(36, flatten=True), mats[36])
    with open(seq, "r+b") as f:
        f.truncate(len(ref) - 100)
    with formats.repr_reader(seq) as r, pytest.raises(ValueError, match="truncated"):
        r.read_all()
    tri = rng.random(300 * 299 // 2).astype(np.float32)
    dw = formats.dist_writer(str(tmp_path / "t.mm-dist"), np.float32, 300)
    dw.write_triangle(tri)
    dw.close()
    got, n = formats.dist_reader.read_triangle(str(tmp_path / "t.mm-dist"))
    assert n == 300 and np.array_equal(got, tri)
    assert os.path.getsize(str(tmp_path / "t.mm-dist")) == formats.DIST_HEADER + tri.nbytes
    buf = np.zeros(1000, dtype=np.uint8)
    with open(str(tmp_path / "t.mm-dist"), "rb") as f:
        assert formats.pread_into(f.fileno(), buf, formats.DIST_HEADER + tri.nbytes - 10) == 10   # short at EOF


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: outside oracle/ itself, only tests/, __graft_entry__.py (build of the
    checker + smoke()) and bench.py (cpu_baseline / --impl reference) may import, load or execute it -- in
    particular nothing under kameris_b200/."""
    allowed = {"bench.py", "__graft_entry__.py"}
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libkameris_oracle|oracle/_ref|librefcode|librefdists", re.M)
    offenders = []
    for dirpath, dirs, files in os.walk(ROOT):
        rel = os.path.relpath(dirpath, ROOT)
        dirs[:] = [d for d in dirs if not d.startswith(".") and d not in ("gpurun_out", "__pycache__")]
        if rel.split(os.sep)[0] in ("oracle", "tests"):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")) or (rel.endswith("scripts") and "." not in fn):
                path = os.path.relpath(os.path.join(dirpath, fn), ROOT)
                if path in allowed:
                    continue
                with open(os.path.join(dirpath, fn), "rb") as f:
                    if pat.search(f.read().decode(errors="replace")):
                        offenders.append(path)
    assert offenders == []


def test_cli_usage_and_errors_without_gpu(tmp_path, capsys):
    """the argv-compatible shims (backend.py:36-39, :61-65): usage / bad input end with a non-zero code and a
    message on stdout (the reference aborts; _command.py:17-18 turns that into CalledProcessError) before any
    device work."""
    from kameris_b200 import cli
    assert cli.main([]) == 1
    assert cli.main(["generation_cgr", "cgr"]) == 1
    assert "Usage: generation_cgr" in capsys.readouterr().out
    assert cli.main(["generation_cgr", "twocgr", str(tmp_path), str(tmp_path / "o"), "4", "16"]) == 1
    assert cli.main(["generation_dists", "a", "b"]) == 1
    assert cli.main(["generation_dists", "a", "b", "chebyshev"]) == 1
    assert "Usage: generation_dists" in capsys.readouterr().out
    assert cli.main(["generation_dists", str(tmp_path / "missing.mm-repr"), str(tmp_path / "p"), "manhat,info"]) == 1
    assert cli.main(["generation_cgr", "cgr", str(tmp_path / "no such dir"), str(tmp_path / "o"), "4", "16"]) == 1
    empty = tmp_path / "fa"
    empty.mkdir()
    (empty / "x.fasta").write_text(">header only\n")
    assert cli.main(["generation_cgr", "cgr", str(empty), str(tmp_path / "o"), "4", "16"]) == 1
    assert "no FASTA record" in capsys.readouterr().out


def test_header_is_plain_c_and_the_example_links(tmp_path):
    """include/kameris_b200.h is a C header (C99, -pedantic -Werror) and examples/kmers_host.c -- the plain-C
    caller INTEGRATION.md shows -- compiles and links against the shared library (no compute call here)."""
    import shutil
    import subprocess
    from kameris_b200 import build
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    build.build()
    exe = str(tmp_path / "kmers_host")
    libdir = os.path.dirname(build.LIB)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "kmers_host.c"), "-L", libdir, "-lkameris_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"usage:" in r.stderr


# ---- A.1 front half, adversarial inputs with hand-written expectations ---------------------------------------
# getline / trim / '>' / record[0] (generation_cgr_mac_avx2 @0x10000d691-0x10000d83e) cannot be extracted from the
# reference binaries as a leaf function, so this part of the path is SPEC-PINNED only: the
# library and the oracle are both written from the same reading of the disassembly.  The expectations below are
# therefore literals derived by hand from that reading -- std::getline splits at '\n' only; boost::trim strips
# std::isspace characters (space, \t \n \v \f \r) from both ends of the line and nothing inside it; a trimmed line
# that is empty or starts with '>' ends the current record; NUL is an ordinary character -- so that a change in
# EITHER implementation shows up.
ADVERSARIAL = [
    (b"ACGT", b"ACGT"),                                     # no header, no trailing newline
    (b">h\r\nAC\r\nGT\r\n", b"ACGT"),                       # CRLF
    (b">h\nAC\n\nGT\n", b"AC"),                             # a blank line ends the record
    (b"\n\n\nAC\nGT", b"ACGT"),                             # leading blank lines open nothing
    (b"  AC \t\n\tGT  \n", b"ACGT"),                        # leading / trailing blanks and tabs
    (b"AC GT\n", b"AC GT"),                                 # inner whitespace is kept (and skipped by the CGR loop)
    (b"AC\x00GT\nTT\n", b"AC\x00GTTT"),                     # NUL bytes are ordinary characters
    (b">a\n>b\n>c\nGG\n>d\nTT", b"GG"),                     # consecutive headers, second record ignored
    (b" >indented header\nAC\n", b"AC"),                    # the '>' test sees the TRIMMED line
    (b"AC\n \t \nGT\n", b"AC"),                             # a whitespace-only line is a blank line
    (b"AC\rGT\n", b"AC\rGT"),                               # a lone CR inside a line is not a line break
    (b"\x0bAC\x0c\n", b"AC"),                               # VT / FF are whitespace
    (b"AC\n>", b"AC"),                                      # a final '>' without newline
    (b">h\nacgtnN\nRYKM\n", b"acgtnNRYKM"),                 # lower case and IUPAC stay in the record
    (b"A" * 70000 + b"\r\n" + b"C" * 3, b"A" * 70000 + b"CCC"),
]
NO_RECORD = [b"", b">only header", b">h\n", b"\n\r\n \t\n", b">a\n>b\n\n"]


def test_fasta_record0_adversarial_literals(tmp_path):
    import ctypes
    from kameris_b200 import _lib
    from kameris_b200.job_steps import backend
    from oracle import oracle as O
    lib = _lib.load()
    d = tmp_path / "fa"
    d.mkdir()
    for i, (body, want) in enumerate(ADVERSARIAL):
        out = ctypes.create_string_buffer(len(body) + 1)
        n = lib.kam_fasta_record0(body, len(body), out)
        assert n == len(want) and out.raw[:n] == want, (i, body[:40])
        assert O.fasta_record0(body) == want, (i, body[:40])
        (d / ("f%03d.fa" % i)).write_bytes(body)
    paths = backend.list_input_files(str(d))
    chars, start = backend.read_records_packed(paths)
    for i, (_, want) in enumerate(ADVERSARIAL):
        assert chars[int(start[i]):int(start[i + 1])].tobytes() == want, i
    for body in NO_RECORD:
        out = ctypes.create_string_buffer(len(body) + 1)
        assert lib.kam_fasta_record0(body, len(body), out) == -1, body
        with pytest.raises(ValueError):
            O.fasta_record0(body)
        p = tmp_path / "none.fa"
        p.write_bytes(body)
        with pytest.raises(ValueError, match="none.fa"):
            backend.read_records_packed([str(p)])


def test_directory_entries_that_are_not_regular_files_fail_fast(tmp_path):
    """generation_cgr takes EVERY directory entry (no regular-file filter, @0x1000017c9): a sub-directory, a FIFO
    or a dangling symlink in the FASTA directory must fail the step with the entry's name, not hang or be skipped."""
    from kameris_b200.job_steps import backend
    d = tmp_path / "fa"
    d.mkdir()
    (d / "a.fa").write_bytes(b">a\nACGT\n")
    for name, make in (("b_dir", lambda p: os.mkdir(p)), ("b_fifo", lambda p: os.mkfifo(p)),
                       ("b_dangling", lambda p: os.symlink(str(tmp_path / "nowhere"), p))):
        p = str(d / name)
        make(p)
        with pytest.raises(OSError, match=name):
            backend.read_records_packed(backend.list_input_files(str(d)))
        if os.path.isdir(p) and not os.path.islink(p):
            os.rmdir(p)
        else:
            os.unlink(p)
    chars, start = backend.read_records_packed(backend.list_input_files(str(d)))
    assert chars.tobytes() == b"ACGT"
