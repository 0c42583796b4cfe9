"""`manhat` and `info` on the tensor cores (thermometer-coded rows, gram.cu kam_thermo_*; dist_tile.cu picks the
path): bit-exact against the reference's own row-lambda machine code (golden cases), against the oracle, and
against the CUDA-core tile kernels on the same inputs."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import HIV_P, synth_records

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture()
def D():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200 import _lib
    from kameris_b200 import dist as D
    yield D
    _lib.load().kam_dist_set_tensor(1)


def _mode(m):
    from kameris_b200 import _lib
    _lib.load().kam_dist_set_tensor(m)


def _dev(a):
    view = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32}
    tdt = {np.dtype(np.uint16): torch.uint16, np.dtype(np.uint32): torch.uint32}
    if a.dtype in view:
        return torch.from_numpy(a.view(view[a.dtype])).cuda().view(tdt[a.dtype])
    return torch.from_numpy(a).cuda()


def _counts(n, L, k, seed):
    recs = synth_records(n, L, seed=seed, p=HIV_P, jitter=L // 10)
    return np.stack([O.cgr_count(r, k, 16) for r in recs])


def test_golden_reference_machine_code_through_the_tensor_path(D, refcode_dists_cases):
    """every integer golden case of the reference's manhat / info machine code, forced through the tensor path
    (mode 2: any shape; cases whose largest count passes 16 fall back by themselves)"""
    _mode(2)
    ran = 0
    for name, x, man, inf in refcode_dists_cases:
        if x.dtype not in (np.uint8, np.uint16, np.uint32):
            continue
        xd = _dev(x)
        assert np.array_equal(D.manhattan_tri(xd).cpu().numpy().view(np.uint32), man), name
        assert np.array_equal(D.info_tri(xd).cpu().numpy().view(np.uint32), inf.view(np.uint32)), name
        ran += 1
    assert ran >= 20


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32])
@pytest.mark.parametrize("n,d,top", [(3, 5, 1), (130, 1000, 3), (260, 4096, 16), (300, 333, 17), (513, 256, 2)])
def test_tensor_path_vs_oracle(D, dtype, n, d, top):
    """ragged d (not a multiple of 16 or 128), counts right at and just past the 16-level limit, empty rows"""
    _mode(2)
    rng = np.random.Generator(np.random.PCG64(n * 7 + d))
    x = rng.poisson(0.5, size=(n, d)).clip(0, top).astype(dtype)
    x[rng.integers(0, n), rng.integers(0, d)] = top
    x[n // 2] = 0
    ref_m, ref_i = O.dists_triangle(x, True, True)
    xd = _dev(x)
    assert np.array_equal(D.manhattan_tri(xd).cpu().numpy().view(np.uint32), ref_m)
    assert np.array_equal(D.info_tri(xd).cpu().numpy().view(np.uint32), ref_i.view(np.uint32))


def test_tensor_path_equals_cuda_core_kernels_many_tiles(D):
    """N = 1300 x D = 4096 short-read-like counts: triangle (several waves of 256 x 256 tiles), row ranges,
    a row-strided input, the rectangle against 300 rows -- default mode takes the tensor path at this size"""
    x = _counts(1300, 300, 6, seed=41)
    assert x.max() <= 16
    wide = torch.zeros((1300, 4096 + 64), dtype=torch.uint16, device="cuda")
    wide[:, :4096] = _dev(x)
    xs = wide[:, :4096]                                  # ld > d
    y = _dev(_counts(300, 300, 6, seed=42))
    res = {}
    for mode in (0, 1):
        _mode(mode)
        tri = D.manhattan_tri(xs)
        parts = torch.zeros_like(tri)
        for r in D.balanced_row_ranges(1300, 3):
            D.manhattan_tri(xs, rows=r, out=parts)
        res[mode] = (tri, parts, D.info_tri(xs), D.manhattan_rect(xs, y))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    assert torch.equal(res[1][0], res[1][1])
    rng = np.random.Generator(np.random.PCG64(3))
    tri = res[1][0].cpu().numpy().view(np.uint32)
    for _ in range(100):
        i, j = sorted(rng.choice(1300, 2, replace=False))
        assert tri[O.triangle_index(1300, i, j)] == O.manhat(x[i], x[j])


def test_larger_counts_need_the_levels_workspace(D):
    """counts up to 9 at a shape the default mode accepts: kam_manhattan_workspace_bytes covers 4 levels, so the
    default call falls back to the byte kernel; with kam_manhattan_workspace_bytes_levels it runs on the tensor
    cores -- same bits either way"""
    from kameris_b200 import _lib
    lib = _lib.load()
    rng = np.random.Generator(np.random.PCG64(9))
    x = rng.poisson(1.5, size=(700, 1024)).clip(0, 9).astype(np.uint16)
    x[5, 7] = 9
    xd = _dev(x)
    ref_m, _ = O.dists_triangle(x, True, False)
    assert np.array_equal(D.manhattan_tri(xd).cpu().numpy().view(np.uint32), ref_m)
    out = torch.zeros(D.tri_len(700), dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.kam_manhattan_workspace_bytes_levels(700, 1024, 9), dtype=torch.uint8, device="cuda")
    assert ws.numel() > lib.kam_manhattan_workspace_bytes(700, 1024)
    _lib.check(lib.kam_manhattan_tri(xd.data_ptr(), _lib.KAM_U16, 700, 1024, 1024, 0, 700, out.data_ptr(), ws.data_ptr(),
                                     ws.numel(), torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref_m)
