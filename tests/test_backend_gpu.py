"""GPU tests of the drop-in plugin layer: run_backend_kmers / run_backend_dists / CLI shims on the
bundled demo genomes (BASELINE config 1), checked against the oracle and Appendix B."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from kameris_b200 import formats
from oracle import oracle as O
from tests.helpers import HIV_P

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def backend():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200.job_steps import backend
    return backend


@pytest.fixture()
def genomes(tmp_path):
    d = tmp_path / "fasta"
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "hiv1-genomes"), str(d))
    return str(d)


@pytest.mark.parametrize("bits", [16, 32])
def test_kmers_counts_file(backend, genomes, demo_records, tmp_path, bits):
    out = str(tmp_path / "cgr-counts.mm-repr")
    backend.run_backend_kmers({"fasta_output_dir": genomes, "output_file": out, "k": 6, "bits_per_element": bits,
                               "mode": "counts", "disable_avx": True}, {})
    r = formats.repr_reader(out)
    assert (r.count, r.rows, r.cols, r.value_type) == (4, 1, 4096, 1 if bits == 16 else 2)
    for i, (_, rec) in enumerate(demo_records):               # sorted-path order: A1, A6, B, C
        assert np.array_equal(r.read_matrix(i, flatten=True), O.cgr_count(rec, 6, bits))
    r.file.close()


def test_kmers_frequencies_file_is_float64_like_numpy(backend, genomes, demo_records, tmp_path):
    out = str(tmp_path / "cgrs.mm-repr")
    backend.run_backend_kmers({"fasta_output_dir": genomes, "output_file": out, "k": 9, "bits_per_element": 16,
                               "mode": "frequencies", "disable_avx": False}, {})
    r = formats.repr_reader(out)
    assert (r.count, r.rows, r.cols, r.value_type) == (4, 1, 4 ** 9, 5)
    for i, (_, rec) in enumerate(demo_records):
        cgr = O.cgr_count(rec, 9, 16).reshape(1, -1)
        assert np.array_equal(r.read_matrix(i), cgr / cgr.sum())          # backend.py:49, bit for bit
    r.file.close()


def test_dists_files(backend, genomes, tmp_path):
    rep = str(tmp_path / "cgr-counts.mm-repr")
    backend.run_backend_kmers({"fasta_output_dir": genomes, "output_file": rep, "k": 6, "bits_per_element": 16,
                               "mode": "counts", "disable_avx": True}, {})
    prefix = str(tmp_path / "dists")
    backend.run_backend_dists({"input_file": rep, "output_prefix": prefix, "distances": ["manhat", "info"],
                               "disable_avx": True}, {})
    tri, n = formats.dist_reader.read_triangle(prefix + "-manhat.mm-dist")
    assert n == 4 and tri.dtype == np.uint32 and tri.tolist() == [3904, 4791, 4369, 4773, 4455, 4628]
    inf, _ = formats.dist_reader.read_triangle(prefix + "-info.mm-dist")
    np.testing.assert_allclose(inf, [0.22702530, 0.25923625, 0.25204274, 0.24173106, 0.25185874, 0.24215384],
                               rtol=2e-7)
    m = formats.dist_reader.read_matrix(prefix + "-manhat.mm-dist")
    assert m.shape == (4, 4) and m[1, 0] == 3904


def test_errors_raise_called_process_error(backend, tmp_path):
    empty = tmp_path / "empty"
    empty.mkdir()
    (empty / "x.fasta").write_text(">only a header\n\n")
    with pytest.raises(subprocess.CalledProcessError):
        backend.run_backend_kmers({"fasta_output_dir": str(empty), "output_file": str(tmp_path / "o"), "k": 4,
                                   "bits_per_element": 16, "mode": "counts", "disable_avx": True}, {})
    with pytest.raises(subprocess.CalledProcessError):
        backend.run_backend_dists({"input_file": str(tmp_path / "missing.mm-repr"), "output_prefix": "p",
                                   "distances": ["manhat"], "disable_avx": True}, {})


def test_cli_shims(genomes, demo_records, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = str(tmp_path / "cli.mm-repr")
    exe = os.path.join(ROOT, "kameris_b200", "scripts", "generation_cgr")
    cmd = '"{}" cgr "{}" "{}" {} {}'.format(exe, genomes, out, 5, 16)          # backend.py:36-39 format
    subprocess.check_call(cmd, shell=True)
    r = formats.repr_reader(out)
    assert np.array_equal(r.read_matrix(2, flatten=True), O.cgr_count(demo_records[2][1], 5, 16))
    r.file.close()
    exe2 = os.path.join(ROOT, "kameris_b200", "scripts", "generation_dists")
    subprocess.check_call('"{}" "{}" "{}" {}'.format(exe2, out, str(tmp_path / "c"), "manhat,info"), shell=True)
    assert os.path.exists(str(tmp_path / "c-manhat.mm-dist")) and os.path.exists(str(tmp_path / "c-info.mm-dist"))
    assert subprocess.call([sys.executable, "-m", "kameris_b200.cli", "generation_cgr", "cgr"], cwd=ROOT) != 0


def test_dists_gram_metrics_files(backend, genomes, tmp_path):
    """the metrics the north star adds, through the same step function and file format (float32)."""
    rep = str(tmp_path / "cgr-counts.mm-repr")
    backend.run_backend_kmers({"fasta_output_dir": genomes, "output_file": rep, "k": 6, "bits_per_element": 16,
                               "mode": "counts", "disable_avx": True}, {})
    prefix = str(tmp_path / "g")
    backend.run_backend_dists({"input_file": rep, "output_prefix": prefix,
                               "distances": ["cosine", "pearson", "euclid", "manhat"], "disable_avx": True}, {})
    with formats.repr_reader(rep) as r:
        x = r.read_all()
    for m in ("cosine", "pearson", "euclid"):
        tri, n = formats.dist_reader.read_triangle("%s-%s.mm-dist" % (prefix, m))
        assert n == 4 and tri.dtype == np.float32
        np.testing.assert_allclose(tri, O.cdist_triangle(x, m), rtol=1e-5)
    tri, _ = formats.dist_reader.read_triangle(prefix + "-manhat.mm-dist")
    assert tri.tolist() == [3904, 4791, 4369, 4773, 4455, 4628]
    # a uint64 file has no Gram metrics
    w = formats.repr_writer(str(tmp_path / "u64.mm-repr"), np.zeros((1, 16), dtype=np.uint64), 3, create_file=True)
    for i in range(3):
        w.write_matrix(np.arange(16, dtype=np.uint64).reshape(1, 16) + i)
    w.file.close()
    with pytest.raises(subprocess.CalledProcessError):
        backend.run_backend_dists({"input_file": str(tmp_path / "u64.mm-repr"), "output_prefix": prefix,
                                   "distances": ["cosine"], "disable_avx": True}, {})


@pytest.mark.parametrize("k", [3, 6, 8])
def test_dists_gram_metrics_on_frequencies_file(backend, genomes, tmp_path, k):
    """`mode: frequencies` (the reference's default, demo/hiv1-lanl.yml:30) writes float64 cgr / cgr.sum()
    (backend.py:45-56); the distances step takes that file too: euclid / cosine / pearson between the
    frequency vectors as stored, = scipy on the rows of the file.  k=6, 8: the counts behind the quotients
    are recovered and verified on the device and the tensor-core Gram runs; k=3: smallest count > 8,
    float64 CUDA-core Gram."""
    frq = str(tmp_path / "f.mm-repr")
    backend.run_backend_kmers({"fasta_output_dir": genomes, "output_file": frq, "k": k, "bits_per_element": 16,
                               "mode": "frequencies", "disable_avx": True}, {})
    prefix = str(tmp_path / "fq")
    backend.run_backend_dists({"input_file": frq, "output_prefix": prefix,
                               "distances": ["cosine", "pearson", "euclid"], "disable_avx": True}, {})
    with formats.repr_reader(frq) as r:
        x = r.read_all()
    assert x.dtype == np.float64
    for m in ("cosine", "pearson", "euclid"):
        tri, n = formats.dist_reader.read_triangle("%s-%s.mm-dist" % (prefix, m))
        assert n == 4 and tri.dtype == np.float32
        np.testing.assert_allclose(tri, O.cdist_triangle(x, m), rtol=1e-5, err_msg=m)


def test_kmers_then_distances_hand_over_on_the_device(tmp_path):
    """run-job runs kmers then distances (run_job.py:175-180): with the device-resident hand-over the distances
    step takes the count matrix the kmers step left on the GPU; the files must be byte-identical to the ones made
    by reading the .mm-repr back, for a counts file and for a frequencies file, and a file that changed after the
    kmers step must NOT be served from the device."""
    from kameris_b200 import formats
    from kameris_b200.job_steps import backend
    from tests.helpers import synth_records
    d = tmp_path / "fa"
    d.mkdir()
    recs = synth_records(150, 2500, seed=41, jitter=2000, dirty=0.003, p=HIV_P)
    for i, r in enumerate(recs):
        (d / ("%04d.fa" % i)).write_bytes(b">s\n" + r + b"\n")
    names = ["manhat", "info", "euclid", "cosine", "pearson"]
    limit = backend.RESIDENT_MAX_BYTES
    try:
        for mode in ("counts", "frequencies"):
            out = {}
            for tag, lim in (("file", 0), ("dev", 1 << 30)):
                backend.set_resident_limit(lim)
                hits0 = backend.resident_hits
                rep = str(tmp_path / ("%s-%s.mm-repr" % (mode, tag)))
                backend.run_backend_kmers({"fasta_output_dir": str(d), "output_file": rep, "k": 6, "bits_per_element": 16,
                                           "mode": mode}, {})
                backend.run_backend_dists({"input_file": rep, "output_prefix": str(tmp_path / ("%s-%s" % (mode, tag))),
                                           "distances": names}, {})
                assert (backend.resident_hits > hits0) == (tag == "dev")
                out[tag] = [open(rep, "rb").read()] + [open(str(tmp_path / ("%s-%s-%s.mm-dist" % (mode, tag, m))), "rb").read()
                                                       for m in names]
            for a, b, what in zip(out["file"], out["dev"], ["repr"] + names):
                assert len(a) > 34 and a == b, (mode, what)
        # a .mm-repr rewritten after the kmers step is read from the file again
        backend.set_resident_limit(1 << 30)
        rep = str(tmp_path / "stale.mm-repr")
        backend.run_backend_kmers({"fasta_output_dir": str(d), "output_file": rep, "k": 5, "bits_per_element": 16,
                                   "mode": "counts"}, {})
        with formats.repr_reader(rep) as r:
            x = r.read_all()
        x[0, :] = 7
        w = formats.repr_writer(rep, x[:1], len(x), create_file=True)
        w.write_all(x)
        w.close()
        hits0 = backend.resident_hits
        backend.run_backend_dists({"input_file": rep, "output_prefix": str(tmp_path / "stale"), "distances": ["manhat"]}, {})
        assert backend.resident_hits == hits0
        tri, n = formats.dist_reader.read_triangle(str(tmp_path / "stale-manhat.mm-dist"))
        assert tri[0] == int(np.abs(x[0].astype(np.int64) - x[1].astype(np.int64)).sum())
    finally:
        backend.set_resident_limit(limit)
        backend.release_resident()
