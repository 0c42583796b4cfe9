#!/usr/bin/env python3
"""Golden vectors for the `mds` job step, produced by the REFERENCE'S OWN function.

kameris/job_steps/mds.py cannot be imported here (its module imports kameris_formats, which is not
installed), but its `mds(delta, dim)` function (mds.py:10-35) only needs numpy and scipy: this script
cuts that function's source out of the file where it lies under /root/reference (nothing is copied into
the repository), executes it with the numpy / scipy of this image, and stores inputs and outputs as a
small fixture:

    python oracle/refpy/gen_golden_mds.py            # writes tests/golden/ref_mds.npz

Cases: the manhat triangle of the four demo genomes (k=6), euclidean distance matrices of points in a
low-dimensional space (exactly embeddable: MDS recovers them), uint32 manhat matrices and float32
cosine matrices of synthetic count vectors, a random non-metric matrix (negative eigenvalues among the
selected ones: NaN columns dropped), n from 4 to 300, dim 1..8.  eigsh starts from a random
vector, so signs of the output columns are arbitrary: consumers compare up to a sign per column.
"""
import ast
import gzip  # noqa: F401
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/kameris/job_steps/mds.py"


def load_reference_mds():
    src = open(REF).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "mds"][0]
    code = "from __future__ import division\nimport logging\nimport numpy as np\nimport scipy.sparse.linalg as linalg\n\n" + \
        ast.get_source_segment(src, fn)
    ns = {}
    exec(compile(code, REF, "exec"), ns)
    return ns["mds"]


def full_from_tri(tri, n):
    m = np.zeros((n, n), dtype=tri.dtype)
    iu = np.triu_indices(n, 1)
    m[iu] = tri
    m.T[iu] = tri
    return m


def main():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from tests.helpers import synth_records
    mds = load_reference_mds()
    rng = np.random.Generator(np.random.PCG64(20261020))
    cases = {}

    def add(name, tri, n, dim):
        delta = full_from_tri(tri, n)
        np.random.seed(7)                                   # eigsh's ARPACK start vector
        pts = mds(delta, dim)
        cases[name + "/tri"] = tri
        cases[name + "/n_dim"] = np.array([n, dim])
        cases[name + "/points"] = np.asarray(pts, dtype=np.float64)

    gdir = os.path.join(ROOT, "tests", "golden", "hiv1-genomes")
    recs = [O.fasta_record0(open(os.path.join(gdir, f), "rb").read()) for f in sorted(os.listdir(gdir))]
    x = np.stack([O.cgr_count(r, 6, 16) for r in recs])
    man, _ = O.dists_triangle(x, True, False)
    add("demo_manhat_k6_dim2", man, 4, 2)
    for n, p, dim in ((30, 3, 3), (120, 5, 4), (300, 8, 8), (64, 2, 1)):
        pts = rng.normal(size=(n, p)) * np.linspace(3.0, 1.0, p)
        d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
        add("euclid_points_n%d_p%d_dim%d" % (n, p, dim), d[np.triu_indices(n, 1)].astype(np.float32), n, dim)
    recs = synth_records(150, 2000, seed=5, jitter=500)
    x = np.stack([O.cgr_count(r, 4, 16) for r in recs])
    man, _ = O.dists_triangle(x, True, False)
    add("synth_manhat_n150_dim3", man, 150, 3)
    add("synth_cosine_n150_dim5", O.cdist_triangle(x, "cosine").astype(np.float32), 150, 5)
    # not a metric: B has large eigenvalues of both signs, the negative ones among the top-|lambda| are
    # dropped as NaN columns (mds.py:33-35)
    u = rng.uniform(0.0, 1.0, size=40 * 39 // 2).astype(np.float32)
    add("random_nonmetric_n40_dim6", u, 40, 6)
    out = os.path.join(ROOT, "tests", "golden", "ref_mds.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases) // 3, "cases")


if __name__ == "__main__":
    main()
