"""The multi-GPU `distances` exchange (parallel.gram_ring and parallel.thermo_ring: one fused kam_gram_fused /
kam_thermo_fused launch per rank, operand shards pulled over CUDA IPC by the copy engines, per-chunk arrival
flags; euclid / cosine / pearson AND the reference's manhat / info) exercised on a box with ONE GPU:
2, 3 and 4 ranks (gloo for the small collectives) share cuda:0.  Every block of every rank must equal the
single-GPU triangle exactly and the ranks must cover every pair exactly once (tests/ring_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ring(rows, k, env_extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(len(rows)),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "ring_worker.py"), "gloo", str(k)] + [str(r) for r in rows]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), KAM_THERMO_MAX_AVG="64",
               **env_extra)
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0 and "OK" in out, out[:5000] + "\n...\n" + out[-1500:]
    return out


@pytest.mark.parametrize("rows", [[600, 523], [700, 300, 515], [530, 600, 257, 700]])
def test_fused_ring_ranks_share_one_gpu(rows):
    """shards of different sizes, none a multiple of the 256-row tile; one arrival flag per column tile, so
    every rank's launch waits on many chunks (KAM_RING_CHUNK_TILES=1)"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    out = _ring(rows, 6, {"KAM_RING_CHUNK_TILES": "1"})
    assert "fused" in out


def test_fused_ring_small_and_empty_shards():
    """a shard below one tile, and a rank without rows"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    _ring([300, 40, 0, 129], 5, {"KAM_RING_CHUNK_TILES": "1"})


def test_stepwise_ring_still_agrees():
    """the per-step launches (KAM_RING_FUSED=0) stay selectable"""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    out = _ring([400, 300, 350], 6, {"KAM_RING_FUSED": "0"})
    assert "fused" not in out
