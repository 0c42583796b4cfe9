#!/usr/bin/env python3
"""Generate tests/golden/refcode_dists.json.gz by running the REFERENCE'S OWN machine code for the
`manhat` / `info` row loops (oracle/refcode/refdists.c: the six per-element-type row lambdas of
/root/reference/kameris/scripts/generation_dists_windows_avx2.exe mapped at their preferred image
base and called in-process).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/refcode/gen_golden_dists.py

Each case stores the input matrix (base64 of the little-endian bytes) and the packed triangles the
reference code produced, so the fixtures are usable where /root/reference does not exist.  Shapes
are chosen to walk every code path of the AVX2 loops: d < 16 (scalar), d not a multiple of 16
(tails), d >= 0x960 (the alignment-peeling path), all-zero rows (info: union = 0), counts that wrap
the uint32 accumulator, float inputs (the truncating accumulate, quirk Q5).
"""
import base64
import ctypes
import gzip
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
PE = "/root/reference/kameris/scripts/generation_dists_windows_avx2.exe"
DTYPES = ["<u1", "<u2", "<u4", "<u8", "<f4", "<f8"]


def build():
    os.makedirs(os.path.join(ROOT, "oracle", "_ref"), exist_ok=True)
    so = os.path.join(ROOT, "oracle", "_ref", "librefdists.so")
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "refdists.c")])
    lib = ctypes.CDLL(so)
    rc = lib.refdists_load(PE.encode())
    if rc != 0:
        raise RuntimeError("refdists_load failed: %d" % rc)
    return lib


def ref(lib, x):
    elem = [np.dtype(t) for t in DTYPES].index(x.dtype)
    n, d = x.shape
    m = np.zeros(n * (n - 1) // 2, dtype="<u4")
    f = np.zeros(n * (n - 1) // 2, dtype="<f4")
    rc = lib.refdists_run(elem, x.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n), ctypes.c_uint64(d),
                          m.ctypes.data_as(ctypes.c_void_p), f.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return elem, m, f


def main():
    lib = build()
    from oracle import oracle as O
    rng = np.random.Generator(np.random.PCG64(20261020))
    cases = []

    def add(name, x):
        x = np.ascontiguousarray(x)
        elem, m, f = ref(lib, x)
        cases.append({"name": name, "elem": elem, "shape": list(x.shape),
                      "x": base64.b64encode(x.tobytes()).decode(),
                      "manhat": [int(v) for v in m],
                      "info_bits": [int(v) for v in f.view("<u4")]})       # float32 bit patterns: exact

    for e, dt in enumerate(DTYPES[:4]):
        hi = {0: 256, 1: 65536, 2: 2 ** 32, 3: 2 ** 40}[e]
        for d in (1, 5, 16, 31, 100, 300, 2500):
            n = 5 if d > 1000 else 7
            x = rng.integers(0, hi, size=(n, d), dtype=np.uint64).astype(dt)
            x[rng.integers(0, n)] = 0
            x[:, rng.integers(0, d)] = 0
            add("int-%s-d%d" % (dt, d), x)
        small = rng.integers(0, 6, size=(9, 257), dtype=np.uint64).astype(dt)   # k-mer-count-like, many zeros
        add("counts-%s" % dt, small)
    # uint32 accumulator wrap: |a-b| sums beyond 2^32
    x = np.zeros((3, 40), dtype="<u4")
    x[0] = 0xFFFFFFF0
    x[2] = 7
    add("u32-wrap", x)
    for dt in DTYPES[4:]:
        for d in (3, 64, 500, 2500):
            x = (rng.random((5, d)) * 4.0).astype(dt)
            x[1] = 0
            add("float-%s-d%d" % (dt, d), x)
        xf = O.frequencies(rng.integers(0, 30, size=(1, 256)).astype(np.uint16)[0])
        x = np.stack([xf, np.roll(xf, 3), xf[::-1]]).astype(dt)
        add("freq-%s" % dt, x)                                              # Q5: frequencies -> manhat 0
        x = (rng.random((4, 50)) * 1e6).astype(dt)
        add("float-large-%s" % dt, x)
    # the bundled demo genomes at k=6 (inputs derive from the CGR-pinned oracle, only outputs stored)
    gdir = os.path.join(ROOT, "tests", "golden", "hiv1-genomes")
    recs = [O.fasta_record0(open(os.path.join(gdir, n), "rb").read()) for n in sorted(os.listdir(gdir))]
    for k in (3, 6, 8):
        x = np.stack([O.cgr_count(r, k, 16) for r in recs]).astype("<u2")
        elem, m, f = ref(lib, x)
        cases.append({"name": "demo-k%d" % k, "elem": 1, "shape": list(x.shape), "x": None, "demo_k": k,
                      "manhat": [int(v) for v in m], "info_bits": [int(v) for v in f.view("<u4")]})

    out = os.path.join(ROOT, "tests", "golden", "refcode_dists.json.gz")
    doc = {"source": "machine code of the six row lambdas of kameris/scripts/generation_dists_windows_avx2.exe "
                     "(VA 0x140024610, 0x1400253c0, 0x140025e30, 0x140026830, 0x1400273a0, 0x140027d50), executed by "
                     "oracle/refcode/refdists.c", "cases": cases}
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(json.dumps(doc, sort_keys=True).encode())
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
