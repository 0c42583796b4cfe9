#!/usr/bin/env python3
"""tests/suite.py -- the five BASELINE.json configs end to end on B200s: results checked against the
oracle (bit-exact counts / manhat, 1e-5 relative for frequencies and Gram metrics), throughput
reported per config.  One JSON line per config (rank 0).  A verification harness (it uses the oracle as
the checker), hence under tests/.

    python tests/suite.py --config 1 2 4 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 tests/suite.py --config 3 5

Large synthetic sets are generated ON DEVICE (torch.Generator seeded per rank and batch); the
sampled sequences / pairs that are verified are copied back and recomputed by the CPU oracle.
--scale shrinks configs 3-5 (e.g. --scale 0.1) for quick runs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kameris_b200 import dist as D  # noqa: E402
from kameris_b200 import parallel  # noqa: E402
from kameris_b200 import repr as R  # noqa: E402
from oracle import oracle as O  # noqa: E402

ALPHA = torch.tensor([65, 67, 71, 84], dtype=torch.uint8)     # A C G T
HIV_P = torch.tensor([0.36, 0.18, 0.24, 0.22])


def ctx():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev


def gen_chars(dev, n, L, seed, p=None):
    """n*L ASCII bases on device (+16 B pad for the pack kernel's vector loads)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.zeros(n * L + 16, dtype=torch.uint8, device=dev)
    alpha = ALPHA.to(dev)
    step = 1 << 28
    cdf = torch.cumsum(p.to(dev), 0) if p is not None else None
    for o in range(0, n * L, step):
        m = min(step, n * L - o)
        if cdf is None:
            idx = torch.randint(0, 4, (m,), generator=g, device=dev)
        else:
            idx = torch.bucketize(torch.rand(m, generator=g, device=dev), cdf[:3])
        out[o:o + m] = alpha[idx]
    return out


def rows_of(x, idx, dev):
    """host copy of selected rows of a uint16 count matrix (torch cannot index uint16 on CUDA)."""
    t = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(dev)
    return x.view(torch.int16)[t].cpu().numpy().view(np.uint16)


def pack(dev, chars, n, L):
    start = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    return R.pack_chars(chars, start, n, n * L)


def tmax(x, world, dev):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def allok(ok, world, dev):
    t = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t))


def timed(fn, dev, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return r, e0.elapsed_time(e1) * 1e-3


# ---- config 1 ---------------------------------------------------------------------------------------
def config1(rank, world, dev, scale):
    gdir = os.path.join(ROOT, "tests", "golden", "hiv1-genomes")
    recs = [O.fasta_record0(open(os.path.join(gdir, n), "rb").read()) for n in sorted(os.listdir(gdir))]
    b = R.pack_records(recs, dev)
    (c6, t) = timed(lambda: R.kmer_vectors(b, 6, 16, "counts"), dev, 1)
    x = c6.cpu().numpy()
    ok = all(np.array_equal(x[i], O.cgr_count(r, 6)) for i, r in enumerate(recs))
    g = D.gram_tri(c6, ("euclid", "cosine", "pearson"), euclid_on_freq=True)
    f = x / x.sum(axis=1, keepdims=True)
    ok &= np.allclose(g["euclid"].cpu().numpy(), O.cdist_triangle(f, "euclid"), rtol=1e-5, atol=0)
    ok &= np.allclose(g["euclid"].cpu().numpy(), [0.0107171533, 0.0122965513, 0.0118525572, 0.0119359576,
                                                  0.0116858869, 0.0116338804], rtol=1e-5, atol=0)   # Appendix B
    man = D.manhattan_tri(c6).cpu().numpy().view(np.uint32)
    ok &= man.tolist() == [3904, 4791, 4369, 4773, 4455, 4628]
    return {"config": 1, "what": "demo hiv1 genomes: k=6 counts + euclid(freq)/cosine/pearson/manhat", "ok": bool(ok),
            "n": 4, "kmers_ms": t * 1e3}


# ---- config 2 ---------------------------------------------------------------------------------------
def config2(rank, world, dev, scale):
    n, L = 10000, 9000
    chars = gen_chars(dev, n, L, 20261002 + rank, HIV_P)
    b = pack(dev, chars, n, L)
    outs = {k: torch.empty((n, 4 ** k), dtype=torch.float32, device=dev) for k in range(1, 10)}
    wss = {k: R.kmer_workspace(n, k, dev) for k in range(1, 10)}

    def sweep():      # orders 9, 8: the tiled kernel; orders 7..1: one fused scan at k=7 pooled down in the CTA
        R.kmer_sweep(b, 1, 9, 16, "freq32", out_tensors=outs, workspace=wss[9])
    sweep()
    _, t = timed(sweep, dev, world)
    rng = np.random.Generator(np.random.PCG64(1))
    samp = rng.choice(n, 64, replace=False)
    host = chars.cpu().numpy()
    ok = True
    for i in samp:
        rec = host[i * L:(i + 1) * L].tobytes()
        for k in range(1, 10):
            ref = O.frequencies(O.cgr_count(rec, k, 16)).astype(np.float32)
            ok &= np.allclose(outs[k][i].cpu().numpy(), ref, rtol=1e-6, atol=0)
    tm = tmax(t, world, dev)
    byt = ((L + 3) // 4 + 4 * sum(4 ** k for k in range(1, 10))) * n
    return {"config": 2, "what": "10k x 9 kb per GPU, k=1..9 float32 CGR vectors", "ok": allok(ok, world, dev),
            "n_gpus": world, "seq_per_s": world * n / tm, "ms": tm * 1e3, "GBps_per_gpu": byt / tm / 1e9,
            "verified": "64 sequences x 9 k vs oracle (rtol 1e-6)"}


# ---- config 3 ---------------------------------------------------------------------------------------
def config3(rank, world, dev, scale):
    n_total = int(100000 * scale)
    n_loc = n_total // world
    n_total = n_loc * world
    L, k = 9000, 9
    chars = gen_chars(dev, n_loc, L, 20261003 + rank, HIV_P)
    b = pack(dev, chars, n_loc, L)
    x_loc, t_repr = timed(lambda: R.kmer_vectors(b, k, 16, "counts"), dev, world)
    del chars, b
    torch.cuda.empty_cache()
    # block-cyclic pair ownership: ring exchange of the uint8 operand rows overlapped with the blocks
    warm = parallel.gram_ring(x_loc, ("cosine", "pearson"))       # first call sets up the NCCL P2P connections
    del warm
    torch.cuda.empty_cache()
    (blocks, counts), t_d = timed(lambda: parallel.gram_ring(x_loc, ("cosine", "pearson")), dev, world)
    # verify random pairs of this rank's blocks against scipy; the column rows are fetched from their owner
    rng = np.random.Generator(np.random.PCG64(7 + rank))
    ok = True
    npairs = 0
    if world > 1:
        # only a sample is checked: gather 64 fixed rows of every shard instead of the whole matrix
        samp_rows = np.sort(np.random.Generator(np.random.PCG64(99)).choice(n_loc, 64, replace=False))
        mine = torch.from_numpy(rows_of(x_loc, samp_rows, dev).view(np.uint8)).to(dev)      # NCCL has no 16-bit ints
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        samp = [p.cpu().numpy().view(np.uint16) for p in parts]
    else:
        samp_rows = np.sort(np.random.Generator(np.random.PCG64(99)).choice(n_loc, 64, replace=False))
        samp = [rows_of(x_loc, samp_rows, dev)]
    for blk in blocks:
        a, c = blk["row_shard"], blk["col_shard"]
        r0, r1 = blk["rows"]
        ii = [i for i in rng.integers(r0, max(r1, r0 + 1), size=6) if i < r1]
        if not ii:
            continue
        rows_a = rows_of(x_loc, ii, dev).astype(np.float64) if a == rank else None
        if rows_a is None:          # 'hi' half: the A rows came from the other rank; use sampled A rows instead
            sel = [t for t in range(64) if r0 <= samp_rows[t] < r1][:6]
            if not sel:
                continue
            ii = [int(samp_rows[t]) for t in sel]
            rows_a = samp[a][sel].astype(np.float64)
        cols_b = samp[c].astype(np.float64)
        for m in ("cosine", "pearson"):
            ref = O.cdist_rect(rows_a, cols_b, m)
            got = blk[m][torch.as_tensor(np.array(ii) - r0).to(dev)][:, torch.as_tensor(samp_rows).to(dev)].cpu().numpy()
            mask = np.ones_like(ref, dtype=bool)
            if blk["tri"]:
                mask = samp_rows[None, :] > np.array(ii)[:, None]
            ok &= np.allclose(got[mask], ref[mask], rtol=1e-5, atol=0)
            npairs += int(mask.sum())
    t_dm = tmax(t_d, world, dev)
    pairs = n_total * (n_total - 1) // 2
    return {"config": 3, "what": "N x 9 kb, k=9 counts -> cosine+pearson over all pairs, block-cyclic pair ownership "
                                 "(ring exchange of uint8 operand rows overlapped with the Gram blocks)",
            "ok": allok(ok, world, dev), "n_gpus": world, "n": n_total, "pairs": pairs,
            "pairs_per_s": pairs / t_dm, "dist_ms": t_dm * 1e3, "blocks_per_rank": len(blocks),
            "tflops_equiv_per_gpu": 2.0 * 4 ** k * pairs / t_dm / world / 1e12,
            "repr_vec_per_s": world * n_loc / tmax(t_repr, world, dev),
            "verified": "%d sampled pairs/rank vs scipy (1e-5)" % npairs}


# ---- config 4 ---------------------------------------------------------------------------------------
def config4(rank, world, dev, scale):
    m_total, L, k, c = int(1_000_000 * scale), 150, 6, 1000
    m = m_total // world
    chars = gen_chars(dev, m, L, 20261004 + rank)
    b = pack(dev, chars, m, L)
    x = torch.empty((m, 4 ** k), dtype=torch.uint16, device=dev)
    ws = R.kmer_workspace(m, k, dev)
    R.kmer_vectors(b, k, 16, "counts", out_tensor=x, workspace=ws)          # warm-up (module load, attributes)
    _, t_repr = timed(lambda: R.kmer_vectors(b, k, 16, "counts", out_tensor=x, workspace=ws), dev, world)
    # centroids: mean of 1000 groups of 1000 reads drawn with a fixed seed (same on every rank)
    cchars = gen_chars(dev, c * 100, L, 4242)
    cb = pack(dev, cchars, c * 100, L)
    cx = R.kmer_vectors(cb, k, 16, "counts").to(torch.float32).view(c, 100, 4 ** k).mean(dim=1).contiguous()
    out, t_d = timed(lambda: D.manhattan_rect(x, cx), dev, world)
    rng = np.random.Generator(np.random.PCG64(3 + rank))
    ii = rng.integers(0, m, size=100)
    xs = rows_of(x, ii, dev).astype(np.float64)
    ref = O.cdist_rect(xs, cx.cpu().numpy().astype(np.float64), "cityblock")
    got = out[torch.from_numpy(ii).to(dev)].cpu().numpy()
    max_rel = float(np.abs(got / ref - 1).max())
    ok = max_rel < 1e-5
    host = chars.cpu().numpy()
    for i in ii[:32]:
        ok &= np.array_equal(rows_of(x, [int(i)], dev)[0], O.cgr_count(host[i * L:(i + 1) * L].tobytes(), k, 16))
    t_dm = tmax(t_d, world, dev)
    return {"config": 4, "what": "short reads x 150 bp, k=6 counts + float32 L1 to 1k centroids", "ok": allok(ok, world, dev),
            "n_gpus": world, "reads": m * world, "repr_vec_per_s": world * m / tmax(t_repr, world, dev),
            "pairs_per_s": world * m * c / t_dm, "dist_ms": t_dm * 1e3, "l1_max_rel_err": max_rel,
            "verified": "100 reads x 1000 centroids/rank vs scipy float64 (1e-5), 32 count vectors bit-exact"}


# ---- config 5 ---------------------------------------------------------------------------------------
def config5(rank, world, dev, scale):
    n_total, L, k = int(5000 * scale), 5_000_000, 9
    n = max(1, n_total // world)
    batch = 148        # one batch = 148 x 4 tile work items of the long-sequence kernel: four full waves
    ok = True
    t_sum = 0.0
    done = 0
    ws = R.kmer_workspace(batch, k, dev)
    out = torch.empty((batch, 4 ** k), dtype=torch.float32, device=dev)
    for b0 in range(0, n, batch):
        nb = min(batch, n - b0)
        chars = gen_chars(dev, nb, L, 20261005 + rank * 100003 + b0)
        pb = pack(dev, chars, nb, L)
        _, t = timed(lambda: R.kmer_vectors(pb, k, 16, "freq32", out_tensor=out[:nb], workspace=ws), dev, 1)
        t_sum += t
        done += nb
        if b0 == 0:
            rec = chars[:L].cpu().numpy().tobytes()
            ref = O.frequencies(O.cgr_count(rec, k, 16)).astype(np.float32)
            ok &= np.allclose(out[0].cpu().numpy(), ref, rtol=1e-6, atol=0)
        del chars, pb
    tm = tmax(t_sum, world, dev)
    return {"config": 5, "what": "bacterial-scale sequences x 5 Mb, k=9 float32 CGR vectors, sharded by sequence",
            "ok": allok(ok, world, dev), "n_gpus": world, "n": done * world, "seq_per_s": world * done / tm,
            "kmers_per_s": world * done * L / tm, "ms": tm * 1e3, "verified": "first sequence of each rank vs oracle"}


# ---- config 6: the drop-in plugin path with FILES (what `kameris run-job` executes) --------------------
def config6(rank, world, dev, scale):
    """tests/fixtures/hiv1-lanl-small.yml-like job on a synthetic directory: kmers (frequencies, k=9),
    kmers (counts, k=6), distances [manhat] -- through run_backend_kmers / run_backend_dists with real
    files, next to the reference pipeline's port (threaded C counts + the numpy loop of backend.py:45-56
    + the same file formats), both reading the same directory."""
    import shutil
    import tempfile
    from kameris_b200 import formats
    from kameris_b200.job_steps import backend
    if rank != 0:
        return {}
    n, L = int(2000 * scale) or 50, 9000
    root = tempfile.mkdtemp(prefix="kam_suite_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fdir = os.path.join(root, "fasta")
        os.mkdir(fdir)
        rng = np.random.Generator(np.random.PCG64(6))
        alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
        for i in range(n):
            seq = alpha[rng.choice(4, size=L, p=[0.36, 0.18, 0.24, 0.22])].tobytes()
            with open(os.path.join(fdir, "%010d.fasta" % i), "wb") as f:      # selection.py:141-149 naming
                f.write(b">%d\n" % i + seq + b"\n")
        t = {}
        backend.run_backend_kmers({"fasta_output_dir": fdir, "output_file": os.path.join(root, "warm.mm-repr"), "k": 3,
                                   "bits_per_element": 16, "mode": "counts", "disable_avx": False}, {})   # CUDA context, library
        t0 = time.perf_counter()
        backend.run_backend_kmers({"fasta_output_dir": fdir, "output_file": os.path.join(root, "cgrs.mm-repr"), "k": 9,
                                   "bits_per_element": 16, "mode": "frequencies", "disable_avx": False}, {})
        t["ours_kmers_freq_k9_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        backend.run_backend_kmers({"fasta_output_dir": fdir, "output_file": os.path.join(root, "counts.mm-repr"), "k": 6,
                                   "bits_per_element": 16, "mode": "counts", "disable_avx": False}, {})
        t["ours_kmers_counts_k6_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        backend.run_backend_dists({"input_file": os.path.join(root, "counts.mm-repr"), "output_prefix": os.path.join(root, "d"),
                                   "distances": ["manhat"], "disable_avx": False}, {})
        t["ours_dists_manhat_s"] = time.perf_counter() - t0

        # the reference pipeline (port): read files, threaded counts, numpy frequency loop, write files
        threads = O.host_threads()
        paths = backend.list_input_files(fdir)
        t0 = time.perf_counter()
        recs = [O.fasta_record0(open(p, "rb").read()) for p in paths]
        chars = np.frombuffer(b"".join(recs), dtype=np.uint8)
        start = np.concatenate([[0], np.cumsum([len(r) for r in recs])]).astype(np.uint64)
        c9 = O.cgr_batch(chars, start, 9, 16, threads)
        w = formats.repr_writer(os.path.join(root, "ref9.mm-repr"), np.zeros((1, 4 ** 9), np.uint16), n)
        w.write_all(c9)
        w.file.close()
        reader = formats.repr_reader(os.path.join(root, "ref9.mm-repr"))           # backend.py:45-56 verbatim
        cgrs = []
        for i in range(reader.count):
            cgr = reader.read_matrix(i)
            cgrs.append(cgr / cgr.sum())
        reader.file.close()
        writer = formats.repr_writer(os.path.join(root, "ref9.mm-repr"), cgrs[0], len(cgrs), create_file=True)
        for cgr in cgrs:
            writer.write_matrix(cgr)
        writer.file.close()
        t["port_kmers_freq_k9_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        c6 = O.cgr_batch(chars, start, 6, 16, threads)
        man, _ = O.dists_triangle(c6, True, False, threads)
        t["port_counts_k6_plus_manhat_s"] = time.perf_counter() - t0

        ok = open(os.path.join(root, "cgrs.mm-repr"), "rb").read() == open(os.path.join(root, "ref9.mm-repr"), "rb").read()
        tri, nn = formats.dist_reader.read_triangle(os.path.join(root, "d-manhat.mm-dist"))
        ok = ok and nn == n and np.array_equal(tri, man)
        return {"config": 6, "what": "plugin path with files: %d FASTA files x %d bp -> .mm-repr (k=9 float64 frequencies), "
                                     ".mm-repr (k=6 counts), -manhat.mm-dist" % (n, L),
                "ok": bool(ok), "files_identical_to_port": bool(ok), "host_cores": threads,
                **{k: round(v, 3) for k, v in t.items()}}
    finally:
        shutil.rmtree(root, ignore_errors=True)


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5, 6: config6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="+", default=[1, 2, 4, 5])
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    rank, world, dev = ctx()
    for c in args.config:
        t0 = time.time()
        r = CONFIGS[c](rank, world, dev, args.scale)
        torch.cuda.empty_cache()
        if rank == 0 and r:
            r["wall_s"] = round(time.time() - t0, 1)
            print(json.dumps(r), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
