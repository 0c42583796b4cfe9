"""Two-rank runs of the drop-in job steps on a box with >= 2 GPUs (skipped on one GPU): the CLI shims
under torch.distributed.run, one process per GPU over NCCL, must write byte-identical `.mm-repr` /
`.mm-dist` files to the single-process steps."""
import os
import socket
import subprocess
import sys

import pytest

from tests.helpers import synth_records

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, args, cwd):
    cmd = [sys.executable]
    if world > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                "--master-port", str(_free_port())]
    cmd += ["-m", "kameris_b200.cli"] + args
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0, r.stdout.decode(errors="replace")[-3000:]


@pytest.mark.parametrize("world", [2])
def test_cli_steps_two_ranks_equal_single_process(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    root = str(tmp_path)
    fdir = os.path.join(root, "fasta")
    os.mkdir(fdir)
    recs = synth_records(61, 3000, seed=8, jitter=2500, dirty=0.002)
    for i, r in enumerate(recs):
        with open(os.path.join(fdir, "%05d.fasta" % i), "wb") as f:
            f.write(b">s%d\n" % i + r + b"\n")
    names = "manhat,info,euclid,cosine,pearson"
    for tag, w in (("one", 1), ("two", world)):
        _run(w, ["generation_cgr", "cgr", fdir, os.path.join(root, tag + ".mm-repr"), "6", "16"], root)
        _run(w, ["generation_dists", os.path.join(root, tag + ".mm-repr"), os.path.join(root, tag), names], root)
    files = ["%s.mm-repr"] + ["%s-" + m + ".mm-dist" for m in names.split(",")]
    for f in files:
        a = open(os.path.join(root, f % "one"), "rb").read()
        b = open(os.path.join(root, f % "two"), "rb").read()
        assert len(a) > 34 and a == b, f
