"""Byte-identity self-check of the sharded job steps, run by `bench.py --gpus N` (N > 1) inside its own
process group: the `kmers` and `distances` steps on N ranks (sequence shards; ring-scheduled Gram blocks and
all-gathered row blocks) must write the same `.mm-repr` / `.mm-dist` bytes as the same steps in one process.
The driver's GPU test box has a single GPU, so tests/test_multi_gpu.py is skipped there; this puts the same
comparison into every multi-GPU bench run.  Product code only (no oracle): it compares the library with itself.
"""
import hashlib
import os
import shutil
import tempfile

import numpy as np

NAMES = ("manhat", "info", "euclid", "cosine", "pearson")


def _records(n, length, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    other = np.frombuffer(b"NRYacgt-", dtype=np.uint8)
    recs = []
    for _ in range(n):
        m = int(rng.integers(length // 3, length * 2))
        s = alpha[rng.choice(4, size=m, p=[0.36, 0.18, 0.24, 0.22])].copy()
        bad = rng.random(m) < 0.003
        s[bad] = other[rng.integers(0, len(other), size=int(bad.sum()))]
        s[int(rng.integers(0, 6))] = ord("N")           # a dirty head (quirk Q1)
        recs.append(s.tobytes())
    return recs


def _sha(path):
    h = hashlib.sha1()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def run(dist_pg, rank, world, n=257, length=3000, k=6):
    """returns {"ok": bool, "files": ..., "detail": ...} on rank 0 (other ranks: {"ok": bool})"""
    import torch
    from .job_steps import backend
    dev = torch.device("cuda", torch.cuda.current_device())
    root = None
    box = [None]
    if rank == 0:
        root = tempfile.mkdtemp(prefix="kam_selfcheck_")
        os.mkdir(os.path.join(root, "fasta"))
        for i, r in enumerate(_records(n, length, 77)):
            with open(os.path.join(root, "fasta", "%05d.fasta" % i), "wb") as f:
                f.write(b">s%d\n" % i + r + b"\n")
        box[0] = root
    dist_pg.broadcast_object_list(box, src=0)
    root = box[0]
    fdir = os.path.join(root, "fasta")
    detail = {}
    try:
        for mode, tag in (("counts", "c"), ("frequencies", "f")):
            opt = {"fasta_output_dir": fdir, "output_file": os.path.join(root, "many-%s.mm-repr" % tag), "k": k,
                   "bits_per_element": 16, "mode": mode}
            backend.run_backend_kmers(opt, {})                       # sharded: the group is initialised
            backend.run_backend_dists({"input_file": opt["output_file"], "output_prefix": os.path.join(root, "many-" + tag),
                                       "distances": list(NAMES)}, {})
        ok = True
        if rank == 0:
            for mode, tag in (("counts", "c"), ("frequencies", "f")):
                opt = {"fasta_output_dir": fdir, "output_file": os.path.join(root, "one-%s.mm-repr" % tag), "k": k,
                       "bits_per_element": 16, "mode": mode}
                backend._kmers_single(opt)
                backend._dists_single({"input_file": opt["output_file"], "output_prefix": os.path.join(root, "one-" + tag)},
                                      backend.requested_metrics(list(NAMES)))
                files = ["%s-" + tag + ".mm-repr"] + ["%s-" + tag + "-" + m + ".mm-dist" for m in NAMES]
                for f in files:
                    a, b = _sha(os.path.join(root, f % "one")), _sha(os.path.join(root, f % "many"))
                    detail[f % "many"] = "identical" if a == b else "DIFFERENT"
                    ok &= a == b
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist_pg.broadcast(flag, src=0)
        ok = bool(int(flag))
    finally:
        dist_pg.barrier()
        if rank == 0:
            shutil.rmtree(root, ignore_errors=True)
    return {"ok": ok, "ranks": world, "vectors": n, "k": k,
            "what": "kmers (counts, frequencies) + distances (%s) job steps on %d ranks vs one process: sha1 of every "
                    "output file" % (",".join(NAMES), world), "files": detail}
