import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def refcode_cases():
    """Outputs of the reference's own CGR machine code (oracle/refcode/gen_golden.py)."""
    with gzip.open(os.path.join(GOLDEN, "refcode_cgr.json.gz")) as f:
        return json.loads(f.read())["cases"]


def _record_of(case):
    """a golden case's input: stored verbatim, or as a recipe for the long trivial ones"""
    if case.get("record") is not None:
        return case["record"].encode("latin-1")
    r = case["recipe"]
    return r.get("prefix", "").encode() + r["repeat"].encode() * r["times"]


@pytest.fixture(scope="session")
def refcode_cases_u32():
    """Outputs of the reference's own CGR machine code, bits_per_element = 32 instantiation
    (oracle/refcode/gen_golden_u32.py): (name, record bytes, {k: summary})."""
    with gzip.open(os.path.join(GOLDEN, "refcode_cgr_u32.json.gz")) as f:
        cases = json.loads(f.read())["cases"]
    return [(c["name"], _record_of(c), c["k"]) for c in cases]


@pytest.fixture(scope="session")
def demo_records():
    """record[0] of the 4 bundled HIV-1 genomes, in the reference's sorted-path order."""
    from oracle import oracle as O
    d = os.path.join(GOLDEN, "hiv1-genomes")
    out = []
    for n in sorted(os.listdir(d)):
        with open(os.path.join(d, n), "rb") as f:
            out.append((n, O.fasta_record0(f.read())))
    return out


@pytest.fixture(scope="session")
def refcode_dists_cases():
    """Outputs of the reference's own manhat/info row-loop machine code (oracle/refcode/gen_golden_dists.py)."""
    import base64
    import numpy as np
    from oracle import oracle as O
    with gzip.open(os.path.join(GOLDEN, "refcode_dists.json.gz")) as f:
        cases = json.loads(f.read())["cases"]
    gdir = os.path.join(GOLDEN, "hiv1-genomes")
    recs = [O.fasta_record0(open(os.path.join(gdir, n), "rb").read()) for n in sorted(os.listdir(gdir))]
    out = []
    for c in cases:
        dt = np.dtype(O.ELEM_DTYPES[c["elem"]])
        if c["x"] is None:                     # demo genomes: inputs come from the CGR-pinned oracle
            x = np.stack([O.cgr_count(r, c["demo_k"], 16) for r in recs])
        else:
            x = np.frombuffer(base64.b64decode(c["x"]), dtype=dt).reshape(c["shape"]).copy()
        out.append((c["name"], x, np.array(c["manhat"], dtype=np.uint32),
                    np.array(c["info_bits"], dtype=np.uint32).view(np.float32)))
    return out
