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
