"""`mds` job step (kameris/job_steps/mds.py:10-35).  CPU: the oracle restatement against golden
vectors produced by the reference's own function (oracle/refpy/gen_golden_mds.py).  GPU (-m gpu): the
CUDA path (kam_mds_center + subspace iteration on kam_symm_block, through the C ABI) against the same
golden vectors and the oracle, tolerance 1e-6 relative per column up to the column's sign (eigenvector
signs are arbitrary in ARPACK too; north-star tolerance 1e-5)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mds.npz")


def _cases():
    z = np.load(GOLD)
    for name in sorted({k.split("/")[0] for k in z.files}):
        n, dim = (int(v) for v in z[name + "/n_dim"])
        yield name, z[name + "/tri"], n, dim, z[name + "/points"]


def test_oracle_restatement_matches_reference_function():
    names = []
    for name, tri, n, dim, ref in _cases():
        got = O.mds_reference(O.full_from_triangle(tri, n), dim)
        assert O.same_up_to_column_signs(got, ref, 1e-9), name
        names.append(name)
    assert len(names) >= 8 and any(ref.shape[1] < dim for _, _, _, dim, ref in _cases())   # a NaN-column case exists


def test_mds_points_reproduce_euclidean_distances():
    """property of the definition, on the oracle: distances between embedded points equal the input
    distances when the input is Euclidean of rank <= dim."""
    rng = np.random.Generator(np.random.PCG64(3))
    p = rng.normal(size=(40, 3))
    d = np.sqrt(((p[:, None] - p[None]) ** 2).sum(-1))
    q = O.mds_reference(d, 3)
    d2 = np.sqrt(((q[:, None] - q[None]) ** 2).sum(-1))
    np.testing.assert_allclose(d2, d, atol=1e-9)


# ---- GPU ------------------------------------------------------------------------------------------------
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200 import mds as M
    return M


def _dev_tri(tri):
    t = tri.view(np.int32) if tri.dtype == np.uint32 else tri.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(t)).cuda()


@pytest.mark.gpu
def test_center_kernel_vs_numpy(M):
    for name, tri, n, dim, _ in _cases():
        B = M.center(_dev_tri(tri), n).cpu().numpy()
        sq = O.full_from_triangle(tri, n).astype(float) ** 2
        tot = sq.sum(axis=0) / n
        ref = (sq - tot[None, :] - tot[:, None] + tot.sum() / n) * -0.5
        np.testing.assert_allclose(B, ref, rtol=0, atol=1e-12 * np.abs(ref).max(), err_msg=name)
        np.testing.assert_allclose(B, B.T, rtol=0, atol=1e-12 * np.abs(ref).max())   # the operation order of mds.py:19-24 is not symmetric in the last bit


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 5, 16, 33, 64, 66, 257, 1000, 2050, 4098])
def test_symm_block_vs_fp64_matmul(M, n):
    """both product kernels (TMA-fed tiles for even n >= 64, register-pipelined otherwise / on request)
    against a float64 matmul; same summation order, so the two kernels agree bit for bit."""
    from kameris_b200 import _lib
    g = torch.Generator(device="cuda")
    g.manual_seed(n)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g)
    B = (A + A.T).contiguous()
    Q = torch.randn((n, M.BLOCK), dtype=torch.float64, device="cuda", generator=g)
    ref = B @ Q
    Y = M.symm_block(B, Q)
    assert float((Y - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    _lib.load().kam_symm_set_tma(0)
    try:
        Y2 = M.symm_block(B, Q)
    finally:
        _lib.load().kam_symm_set_tma(1)
    assert torch.equal(Y, Y2)


@pytest.mark.gpu
def test_mds_vs_reference_golden_and_oracle(M):
    for name, tri, n, dim, ref in _cases():
        stats = {}
        got = M.mds_from_triangle(_dev_tri(tri), n, dim, stats=stats)
        assert O.same_up_to_column_signs(got, ref, 1e-6), (name, stats)
        assert O.same_up_to_column_signs(got, O.mds_reference(O.full_from_triangle(tri, n), dim), 1e-6), name


@pytest.mark.gpu
def test_mds_recovers_embedding_at_size(M):
    """n = 3000 points of a 6-dimensional Euclidean configuration: distances between the embedded points
    equal the input distances (float32 triangle: 1e-5 relative), eigenvalues descending."""
    rng = np.random.Generator(np.random.PCG64(11))
    n, p = 3000, 6
    pts = rng.normal(size=(n, p)) * np.linspace(4.0, 1.0, p)
    d = np.sqrt(np.maximum(((pts ** 2).sum(1)[:, None] + (pts ** 2).sum(1)[None, :] - 2 * pts @ pts.T), 0))
    tri = d[np.triu_indices(n, 1)].astype(np.float32)
    got = M.mds_from_triangle(torch.from_numpy(tri).cuda(), n, p)
    assert got.shape == (n, p)
    norms = (got ** 2).sum(0)
    assert (np.diff(norms) <= 1e-9 * norms[0]).all()
    ii, jj = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    dd = np.sqrt(((got[ii] - got[jj]) ** 2).sum(1))
    np.testing.assert_allclose(dd, d[ii, jj], rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_run_mds_step_files(M, tmp_path):
    from kameris_b200 import formats
    from kameris_b200.job_steps import step_runners
    name, tri, n, dim, ref = [c for c in _cases() if c[0].startswith("synth_manhat")][0]
    w = formats.dist_writer(str(tmp_path / "d-manhat.mm-dist"), np.uint32, n)
    w.write_triangle(tri)
    w.close()
    step_runners["mds"]({"type": "mds", "dists_file": str(tmp_path / "d-manhat.mm-dist"), "dimensions": dim,
                         "output_file": str(tmp_path / "mds.json")}, {})
    pts = np.array(json.load(open(str(tmp_path / "mds.json"))))
    assert O.same_up_to_column_signs(pts, ref, 1e-6)


@pytest.mark.gpu
def test_mds_errors(M):
    tri = torch.zeros(3, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        M.mds_from_triangle(tri, 3, 3)          # eigsh: k must be < n
    with pytest.raises(ValueError):
        M.center(tri, 4)
