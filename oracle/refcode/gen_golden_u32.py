#!/usr/bin/env python3
"""Generate tests/golden/refcode_cgr_u32.json.gz by running the REFERENCE'S OWN machine code for
bits_per_element = 32.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/refcode/gen_golden_u32.py

generation_cgr_windows_avx2.exe holds the CGR fill twice: the uint16 leaf at VA 0x140016040 (gen_golden.py)
and its uint32 twin at VA 0x14001b360 (`inc DWORD PTR [rcx+rax*4]` at VA 0x14001b47d), mapped and called here
through oracle/refcode/refcode.c.  Cases: the quirk records, the demo genomes, random dirty records, and records
whose counts pass 65535 (where the uint16 build wraps and this one must not).  Stored per k: sum / nnz / max /
sha1 of the uint32 little-endian vector (and the full vector when 4^k <= 64).
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
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(HERE, "refcode.c")])
    lib = ctypes.CDLL(so)
    rc = lib.refcode_load_u32(PE.encode())
    if rc != 0:
        raise RuntimeError("refcode_load_u32 failed: %d" % rc)
    rc = lib.refcode_load(PE.encode())
    if rc != 0:
        raise RuntimeError("refcode_load failed: %d" % rc)
    return lib


def ref_cgr32(lib, rec, k):
    out = np.zeros(4 ** k, dtype="<u4")
    assert lib.refcode_cgr_u32(rec, ctypes.c_uint64(len(rec)), ctypes.c_uint32(k), out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def ref_cgr16(lib, rec, k):
    out = np.zeros(4 ** k, dtype="<u2")
    assert lib.refcode_cgr_u16(rec, ctypes.c_uint64(len(rec)), ctypes.c_uint32(k), out.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def record0(text):
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
    rng = np.random.Generator(np.random.PCG64(20261020))
    cases = []

    def add(name, rec, ks, store_record=True, recipe=None):
        entry = {"name": name, "record": rec.decode("latin-1") if store_record else None, "recipe": recipe, "k": {}}
        for k in ks:
            v = ref_cgr32(lib, rec, k)
            # the two instantiations of the same source must agree modulo 2^16
            assert np.array_equal((v & 0xffff).astype("<u2"), ref_cgr16(lib, rec, k)), (name, k)
            e = {"sum": int(v.astype(np.uint64).sum()), "nnz": int((v > 0).sum()),
                 "max": int(v.max()), "sha1": hashlib.sha1(v.tobytes()).hexdigest()}
            if 4 ** k <= 64:
                e["vector"] = [int(t) for t in v]
            entry["k"][str(k)] = e
        cases.append(entry)

    add("toy-ACGT", b"ACGT", [1, 2, 3, 4])
    add("toy-NACGT", b"NACGT", [1, 2, 3, 4])
    add("toy-NNACGTAC", b"NNACGTAC", [1, 2, 3, 5])
    add("toy-lowercase", b"acgtACGT", [1, 2, 3])
    add("empty", b"", [1, 3])
    add("shorter-than-k", b"ACG", [3, 4, 6])
    add("all-invalid", b"NNNNNNNNNNRYKM-*nnacgt", [1, 2, 5])
    add("invalid-run-then-bases", b"N" * 40 + b"ACGTTGCAAC", [1, 3, 9, 12])
    # counts above 65535: stored as a recipe (the records are long and trivial)
    add("polyA-70000", b"A" * 70000, [1, 2, 9], store_record=False, recipe={"repeat": "A", "times": 70000})
    add("polyT-65536", b"T" * 65536, [1, 3], store_record=False, recipe={"repeat": "T", "times": 65536})
    add("AC-200000", b"AC" * 200000, [1, 2, 4, 9], store_record=False, recipe={"repeat": "AC", "times": 200000})
    add("NGATTACA-50000", b"N" + b"GATTACA" * 50000, [1, 3, 7, 10], store_record=False,
        recipe={"prefix": "N", "repeat": "GATTACA", "times": 50000})

    z = zipfile.ZipFile(ZIP)
    for n in sorted(z.namelist()):
        add("demo-" + n, record0(z.read(n)), [1, 3, 6, 9, 11])

    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    iupac = np.frombuffer(b"NRYKMSWBDHVacgtn-", dtype=np.uint8)
    for i, (L, dirty) in enumerate([(7, 0.3), (33, 0.5), (150, 0.05), (1000, 0.002), (4097, 0.01), (9000, 0.002),
                                    (20000, 0.2)]):
        s = alpha[rng.choice(4, size=L, p=[0.36, 0.18, 0.24, 0.22])].copy()
        m = rng.random(L) < dirty
        s[m] = iupac[rng.integers(0, len(iupac), size=int(m.sum()))]
        if L > 3:
            s[rng.integers(0, min(L, 8))] = ord("N")
        add("rand-%d-L%d-dirty%g" % (i, L, dirty), s.tobytes(), [1, 2, 4, 6, 8, 9, 10])

    out = os.path.join(ROOT, "tests", "golden", "refcode_cgr_u32.json.gz")
    doc = {"source": "machine code of kameris/scripts/generation_cgr_windows_avx2.exe VA 0x14001b360-0x14001b4c3 "
                     "(CGR fill, uint32 instantiation), executed by oracle/refcode/refcode.c", "dtype": "<u4",
           "cases": cases}
    with gzip.GzipFile(out, "wb", mtime=0) as f:
        f.write(json.dumps(doc, sort_keys=True).encode())
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases,", sum(len(c["k"]) for c in cases), "vectors")


if __name__ == "__main__":
    main()
