#!/usr/bin/env python3
"""Generate tests/golden/refcode_cgr.json.gz by running the REFERENCE'S OWN machine code.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/refcode/gen_golden.py

It compiles oracle/refcode/refcode.c into oracle/_ref/librefcode.so, which maps the CGR fill
function of /root/reference/kameris/scripts/generation_cgr_windows_avx2.exe (VA 0x140016040,
see refcode.c) and calls it in-process.  Every case stores the input record verbatim plus, per k,
sum / nnz / max / sha1 of the uint16 little-endian count vector (and the full vector when
4^k <= 64), so the fixtures are usable where /root/reference does not exist.
"""
import ctypes
import gzip
import hashlib
import json
import os
import subprocess
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PE = "/root/reference/kameris/scripts/generation_cgr_windows_avx2.exe"
ZIP = "/root/reference/demo/hiv1-genomes.zip"


def build():
    os.makedirs(os.path.join(ROOT, "oracle", "_ref"), exist_ok=True)
    so = os.path.join(ROOT, "oracle", "_ref", "librefcode.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "refcode.c")])
    lib = ctypes.CDLL(so)
    rc = lib.refcode_load(PE.encode())
    if rc != 0:
        raise RuntimeError("refcode_load failed: %d" % rc)
    lib.refcode_order.restype = ctypes.c_char_p
    assert lib.refcode_order() == b"ACGT"
    return lib


def ref_cgr(lib, rec, k):
    out = np.zeros(4 ** k, dtype="<u2")
    lib.refcode_cgr_u16(rec, ctypes.c_uint64(len(rec)), ctypes.c_uint32(k),
                        out.ctypes.data_as(ctypes.c_void_p))
    return out


def record0(text):
    """first FASTA record, A.1 semantics (python mirror used only to feed the machine code)."""
    cur, recs = [], []
    for line in text.split(b"\n"):
        line = line.strip(b" \t\n\v\f\r")
        if not line or line[:1] == b">":
            if cur:
                recs.append(b"".join(cur))
                cur = []
        else:
            cur.append(line)
    if cur:
        recs.append(b"".join(cur))
    return recs[0]


def main():
    lib = build()
    rng = np.random.Generator(np.random.PCG64(20261019))
    cases = []

    def add(name, rec, ks):
        entry = {"name": name, "record": rec.decode("latin-1"), "k": {}}
        for k in ks:
            v = ref_cgr(lib, rec, k)
            e = {"sum": int(v.astype(np.uint64).sum()), "nnz": int((v > 0).sum()),
                 "max": int(v.max()), "sha1": hashlib.sha1(v.tobytes()).hexdigest()}
            if 4 ** k <= 64:
                e["vector"] = [int(t) for t in v]
            entry["k"][str(k)] = e
        cases.append(entry)

    # quirk cases and edge cases
    add("toy-ACGT", b"ACGT", [1, 2, 3, 4])
    add("toy-NACGT", b"NACGT", [1, 2, 3, 4])
    add("toy-NNACGTAC", b"NNACGTAC", [1, 2, 3, 5])
    add("toy-lowercase", b"acgtACGT", [1, 2, 3])
    add("empty", b"", [1, 3])
    add("shorter-than-k", b"ACG", [3, 4, 6])
    add("all-invalid", b"NNNNNNNNNNRYKM-*nnacgt", [1, 2, 5])
    add("invalid-run-then-bases", b"N" * 40 + b"ACGTTGCAAC", [1, 3, 9, 12])
    add("single-base", b"T", [1, 2])
    add("polyA-wrap-u16", b"A" * 70000, [1, 2, 9])          # Q3: 70000 mod 65536
    add("polyT-65536", b"T" * 65536, [1])                    # wraps exactly to 0

    z = zipfile.ZipFile(ZIP)
    for n in sorted(z.namelist()):
        add("demo-" + n, record0(z.read(n)), list(range(1, 13)))

    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    iupac = np.frombuffer(b"NRYKMSWBDHVacgtn-", dtype=np.uint8)
    for i, (L, dirty) in enumerate([(1, 0.0), (7, 0.3), (31, 0.1), (32, 0.0), (33, 0.5), (150, 0.0),
                                    (150, 0.05), (1000, 0.002), (4097, 0.01), (9000, 0.0),
                                    (9000, 0.002), (20000, 0.2), (100003, 0.001)]):
        p = np.array([0.36, 0.18, 0.24, 0.22]) if i % 2 == 0 else np.full(4, 0.25)
        s = alpha[rng.choice(4, size=L, p=p)].copy()
        if dirty > 0:
            m = rng.random(L) < dirty
            s[m] = iupac[rng.integers(0, len(iupac), size=int(m.sum()))]
            if L > 3:
                s[rng.integers(0, min(L, 8))] = ord("N")      # Q1: invalid inside the first k-1
        add("rand-%d-L%d-dirty%g" % (i, L, dirty), s.tobytes(), [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])

    out = os.path.join(ROOT, "tests", "golden", "refcode_cgr.json.gz")
    doc = {"source": "machine code of kameris/scripts/generation_cgr_windows_avx2.exe VA 0x140016040-0x1400161a8 "
                     "(CGR fill, uint16), executed by oracle/refcode/refcode.c", "dtype": "<u2", "cases": cases}
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(doc, sort_keys=True).encode())
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases")

    gdir = os.path.join(ROOT, "tests", "golden", "hiv1-genomes")
    os.makedirs(gdir, exist_ok=True)
    for n in z.namelist():                                   # public sequence DATA for config 1
        with open(os.path.join(gdir, n), "wb") as f:
            f.write(z.read(n))


if __name__ == "__main__":
    main()
