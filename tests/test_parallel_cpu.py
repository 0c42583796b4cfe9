"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sequence sharding, all-gather of the
feature rows, balanced triangle row blocks, slice offsets, gathering the triangle on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kameris_b200 import parallel
from kameris_b200.dist import balanced_row_ranges


def test_shard_by_bases():
    rng = np.random.Generator(np.random.PCG64(0))
    lens = rng.integers(100, 5_000_000, size=1000)
    for world in (1, 2, 4, 8):
        sh = parallel.shard_by_bases(lens, world)
        assert sh[0][0] == 0 and sh[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        tot = [int(lens[a:b].sum()) for a, b in sh]
        assert max(tot) - min(tot) <= 2 * lens.max()
    assert parallel.shard_by_bases([5], 4)[-1] == (1, 1) or parallel.shard_by_bases([5], 4)[0] == (0, 1)


def test_tri_row_offset():
    n = 9
    k = 0
    for r in range(n):
        assert parallel.tri_row_offset(n, r) == k
        k += n - 1 - r
    assert parallel.tri_row_offset(n, n) == n * (n - 1) // 2


def _host_compute(x_full, metrics, rows, tri_offset, slice_len):
    """numpy stand-in for the CUDA kernels (plumbing test only)."""
    from oracle import oracle as O
    x = x_full.view(torch.int16).numpy().view(np.uint16)
    man, _ = O.dists_triangle(x, True, False, threads=1)
    return {"manhat": torch.from_numpy(man[tri_offset:tri_offset + slice_len].view(np.int32).copy())}


def _worker(rank, world, port, x_all, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shards = parallel.shard_by_bases([1] * x_all.shape[0], world)     # ragged on purpose below
        if world == 2:
            shards = [(0, 5), (5, x_all.shape[0])]
        b, e = shards[rank]
        x_local = x_all[b:e].clone()
        x_full, counts = parallel.all_gather_rows(x_local)
        assert counts == [s[1] - s[0] for s in shards] and torch.equal(x_full, x_all)
        slices, (r0, r1), n = parallel.distances_sharded(x_local, ["manhat"], compute=_host_compute)
        assert (r0, r1) == balanced_row_ranges(n, world)[rank]
        full = parallel.gather_triangle(slices["manhat"], n)
        if rank == 0:
            ret.put(full.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_distances_sharded_gloo_world2():
    from oracle import oracle as O
    rng = np.random.Generator(np.random.PCG64(4))
    x = rng.integers(0, 300, size=(23, 64)).astype(np.uint16)
    x_t = torch.from_numpy(x.view(np.int16)).view(torch.uint16)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x_t, ret)) for r in range(2)]
    for p in procs:
        p.start()
    full = ret.get()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, _ = O.dists_triangle(x, True, False)
    assert np.array_equal(full.view(np.uint32), ref)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
def test_ring_schedule_covers_every_shard_pair_once(world):
    plan = parallel.ring_schedule(world)
    cover = {}
    for r, steps in enumerate(plan):
        assert steps[0] == (0, r, r, "tri")
        for s, a, c, part in steps:
            # the rank holds both shards at that step: its own, and the one received at step s
            assert r in (a, c) and {a, c} <= {r, (r + s) % world}
            key = (min(a, c), max(a, c))
            cover.setdefault(key, []).append(part)
    want = {(a, c) for a in range(world) for c in range(a, world)}
    assert set(cover) == want
    for key, parts in cover.items():
        assert sorted(parts) in (["tri"], ["full"], ["hi", "lo"]), (key, parts)
    loads = [sum({"tri": 0.5, "full": 1.0, "lo": 0.5, "hi": 0.5}[p] for _, _, _, p in steps) for steps in plan]
    assert max(loads) - min(loads) < 1e-9          # perfectly balanced in units of blocks


# ---- the two job steps under a process group: every rank writes its part of ONE output file ------------
def _host_compute_all(x_full, metrics, rows, tri_offset, slice_len):
    from oracle import oracle as O
    x = x_full.numpy()
    man, info = O.dists_triangle(x, True, True, threads=1)
    out = {}
    if "manhat" in metrics:
        out["manhat"] = torch.from_numpy(man[tri_offset:tri_offset + slice_len].view(np.int32).copy())
    if "info" in metrics:
        out["info"] = torch.from_numpy(info[tri_offset:tri_offset + slice_len].copy())
    return out


def _steps_worker(rank, world, port, root):
    from kameris_b200 import formats
    from kameris_b200.job_steps import backend
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        paths = backend.list_input_files(os.path.join(root, "fasta"))
        out = os.path.join(root, "counts.mm-repr")

        def produce(records, row_offset, total_rows):      # host stand-in for the CUDA streaming writer
            w = formats.repr_appender(out)
            assert w.count == total_rows
            w.file.seek(formats.REPR_HEADER + row_offset * 4 ** 3 * 2)
            for r in records:
                O.cgr_count(r, 3, 16).astype("<u2").tofile(w.file)
            w.close()
        backend.kmers_to_file_sharded(paths, 3, 16, "counts", out, dist, produce=produce)
        backend.dists_to_files_sharded(out, os.path.join(root, "d"), ["manhat", "info"], dist,
                                       compute=_host_compute_all)
    finally:
        dist.destroy_process_group()


def test_job_steps_sharded_files_gloo_world2(tmp_path):
    """kmers + distances steps on 2 ranks (gloo, host stand-ins for the kernels): the `.mm-repr` and
    `.mm-dist` files assembled from per-rank writes equal the single-process files byte for byte."""
    from kameris_b200 import formats
    from oracle import oracle as O
    from tests.helpers import synth_records
    root = str(tmp_path)
    os.mkdir(os.path.join(root, "fasta"))
    recs = synth_records(11, 300, seed=3, jitter=250, dirty=0.01)
    for i, r in enumerate(recs):
        with open(os.path.join(root, "fasta", "%04d.fasta" % i), "wb") as f:
            f.write(b">x\n" + r + b"\n")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_steps_worker, args=(r, 2, port, root)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = np.stack([O.cgr_count(r, 3, 16) for r in recs])
    with formats.repr_reader(os.path.join(root, "counts.mm-repr")) as rd:
        assert rd.count == 11 and rd.cols == 64 and np.array_equal(rd.read_all(), x)
        assert np.array_equal(rd.read_rows(3, 7), x[3:7])
    man, info = O.dists_triangle(x, True, True)
    tri, n = formats.dist_reader.read_triangle(os.path.join(root, "d-manhat.mm-dist"))
    assert n == 11 and np.array_equal(tri, man)
    tri, n = formats.dist_reader.read_triangle(os.path.join(root, "d-info.mm-dist"))
    assert n == 11 and np.array_equal(tri, info)


def test_kmers_to_file_streaming_writer_with_host_stand_in(tmp_path, monkeypatch):
    """The streaming `.mm-repr` writer of the kmers step (chunks produced into alternating buffers, written at
    their offsets by positional multi-threaded writes) with a host stand-in for kam_kmers_host: several chunks,
    a row offset into a larger file, packed and list inputs -> byte-identical to the file written row by row."""
    import ctypes
    from kameris_b200 import _lib, formats
    from kameris_b200.job_steps import backend
    from oracle import oracle as O
    from tests.helpers import synth_records
    k, bits = 4, 16
    d = 4 ** k
    recs = synth_records(23, 300, seed=12, jitter=100, dirty=0.01)
    real = _lib.load()

    class Fake(object):
        def __getattr__(self, name):
            return getattr(real, name)

        @staticmethod
        def kam_kmers_host_keep(base, starts, n, kk, b, kind, out_ptr, d_counts, stride):
            assert d_counts is None             # no device here: nothing can stay resident
            return Fake.kam_kmers_host(base, starts, n, kk, b, kind, out_ptr)

        @staticmethod
        def kam_kmers_host(base, starts, n, kk, b, kind, out_ptr):
            st = np.ctypeslib.as_array(ctypes.cast(starts, ctypes.POINTER(ctypes.c_uint64)), shape=(n + 1,))
            raw = ctypes.string_at(base, int(st[n])) if base else b""
            out = np.ctypeslib.as_array(ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_uint16)), shape=(n, d))
            for i in range(n):
                out[i] = O.cgr_count(raw[int(st[i]):int(st[i + 1])], kk, b)
            return 0
    monkeypatch.setattr(_lib, "load", lambda: Fake())
    monkeypatch.setattr(backend, "CHUNK_BYTES", 5 * d * 2)                  # 5 rows per chunk -> 5 chunks
    monkeypatch.setattr(formats, "PARALLEL_IO_MIN_BYTES", 256)              # threaded positional writes
    want = str(tmp_path / "want.mm-repr")
    w = formats.repr_writer(want, np.zeros((1, d), dtype=np.uint16), len(recs), create_file=True)
    for r in recs:
        w.write_matrix(O.cgr_count(r, k, bits))
    w.close()
    got = str(tmp_path / "got.mm-repr")
    backend.kmers_to_file(recs, k, bits, "counts", got)
    assert open(got, "rb").read() == open(want, "rb").read()
    got2 = str(tmp_path / "got2.mm-repr")
    backend.kmers_to_file(backend._pack_records(recs), k, bits, "counts", got2)
    assert open(got2, "rb").read() == open(want, "rb").read()
    # a shard written in place into a file another rank created (rows 10..22 of 23)
    got3 = str(tmp_path / "got3.mm-repr")
    w = formats.repr_writer(got3, np.zeros((1, d), dtype=np.uint16), len(recs), create_file=True)
    w.file.truncate(formats.REPR_HEADER + len(recs) * d * 2)
    w.close()
    backend.kmers_to_file(recs[10:], k, bits, "counts", got3, row_offset=10, total_rows=len(recs), create=False)
    backend.kmers_to_file(recs[:10], k, bits, "counts", got3, row_offset=0, total_rows=len(recs), create=False)
    assert open(got3, "rb").read() == open(want, "rb").read()


# ---- a failure on ONE rank must fail EVERY rank (no rank left waiting in a collective), and remove the output ----
def _failing_worker(rank, world, port, root, which, ret):
    from kameris_b200 import formats
    from kameris_b200.job_steps import backend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        paths = backend.list_input_files(os.path.join(root, "fasta"))
        out = os.path.join(root, "bad.mm-repr")
        msg = "no error"
        try:
            if which == "kmers":
                def produce(records, row_offset, total_rows):
                    if rank == 1:
                        raise ValueError("no FASTA record in a file of this shard")
                backend.kmers_to_file_sharded(paths, 3, 16, "counts", out, dist, produce=produce)
            else:
                def compute(x_full, metrics, rows, tri_offset, slice_len):
                    if rank == 1:
                        raise RuntimeError("device out of memory (injected)")
                    return _host_compute_all(x_full, metrics, rows, tri_offset, slice_len)
                backend.dists_to_files_sharded(os.path.join(root, "good.mm-repr"), os.path.join(root, "bad"), ["manhat"],
                                               dist, compute=compute)
        except Exception as e:   # noqa: BLE001
            msg = "%s: %s" % (type(e).__name__, e)
        ret.put((rank, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("which", ["kmers", "dists"])
def test_sharded_steps_agree_on_failure(tmp_path, which):
    from kameris_b200 import formats
    from oracle import oracle as O
    from tests.helpers import synth_records
    root = str(tmp_path)
    os.mkdir(os.path.join(root, "fasta"))
    recs = synth_records(6, 200, seed=4)
    for i, r in enumerate(recs):
        with open(os.path.join(root, "fasta", "%04d.fasta" % i), "wb") as f:
            f.write(b">x\n" + r + b"\n")
    w = formats.repr_writer(os.path.join(root, "good.mm-repr"), np.zeros((1, 64), dtype=np.uint16), 6)
    for r in recs:
        w.write_matrix(O.cgr_count(r, 3, 16))
    w.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_failing_worker, args=(r, 2, port, root, which, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert "failed on another rank" in got[0], got          # the healthy rank is told, it does not hang
    assert ("no FASTA record" in got[1]) if which == "kmers" else ("injected" in got[1]), got
    assert not os.path.exists(os.path.join(root, "bad.mm-repr"))
    assert not os.path.exists(os.path.join(root, "bad-manhat.mm-dist"))


def test_ring_groups_cover_every_step_once():
    """launch groups of the ring: every step s = 1..world//2 exactly once and in order, the first group a single
    step (work starts after ONE shard), merged groups only of full blocks, the shared half block alone and last"""
    for world in range(1, 17):
        for merge in (False, True):
            assert [s for g in parallel.ring_groups(world, merge) for s in g] == list(range(1, world // 2 + 1))
        groups = parallel.ring_groups(world, merge=True)
        flat = [s for g in groups for s in g]
        assert flat == list(range(1, world // 2 + 1)), world
        plan = {s: part for s, _, _, part in parallel.ring_schedule(world)[0]}
        for g in groups:
            assert 1 <= len(g) <= 2
            if len(g) > 1:
                assert all(plan[s] == "full" for s in g), (world, g)
        if groups:
            assert len(groups[0]) == 1
        if world % 2 == 0 and world > 1:
            assert groups[-1] == [world // 2] and plan[world // 2] in ("lo", "hi")


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 6, 7, 8])
def test_fused_plan_covers_every_pair_once_and_segments_hold_every_tile(world):
    """the fused ring's plan (parallel.fused_plan / fused_segments, host logic of kam_gram_fused): over all ranks
    every unordered pair of rows is computed exactly once -- ragged shards, one below a tile, one empty -- and a
    rank's tile segments contain every (row tile, column tile) its entries need, each column tile in one segment"""
    from kameris_b200 import parallel
    counts = [(517, 300, 0, 1025, 40, 256, 777, 513)[r] for r in range(world)]
    offs = np.concatenate([[0], np.cumsum(counts)])
    n = int(offs[-1])
    cover = np.zeros((n, n), dtype=np.int32)
    for rank in range(world):
        plan = parallel.fused_plan(rank, counts)
        segs = parallel.fused_segments(plan)
        assert plan[0]["part"] == "tri" and plan[0]["b_rows"] == (0, counts[rank])
        prev_end = 0
        for c0, c1, r0, r1 in segs:
            assert c0 == prev_end and c1 > c0 and r0 == 0
            prev_end = c1
        assert prev_end == (plan[-1]["b_rows"][1] + 255) // 256
        for e in plan:
            b0, b1 = e["b_rows"]
            s0, s1 = e["src_rows"]
            a0, a1 = e["a_rows"]
            assert b1 - b0 == s1 - s0
            gi = offs[rank] + np.arange(a0, a1)
            gj = offs[e["src"]] + np.arange(s0, s1)
            if e["part"] == "tri":
                for i in range(a0, a1):
                    cover[offs[rank] + i, offs[rank] + i + 1:offs[rank] + a1] += 1
            elif gi.size and gj.size:
                lo, hi = np.minimum.outer(gi, gj), np.maximum.outer(gi, gj)
                np.add.at(cover, (lo.ravel(), hi.ravel()), 1)
            # every tile of this entry's rectangle lies in a segment with enough rows
            for ct in range(b0 // 256, (b1 + 255) // 256 if b1 > b0 else b0 // 256):
                seg = [sg for sg in segs if sg[0] <= ct < sg[1]]
                assert len(seg) == 1
                assert seg[0][3] * 256 >= a1 or seg[0][3] == (counts[rank] + 255) // 256
    iu = np.triu(np.ones((n, n), dtype=bool), 1)
    assert (cover[iu] == 1).all() and (cover[~iu] == 0).all()
