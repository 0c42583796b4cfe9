"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the two paths shard.

  kmers      sequences are independent: contiguous shards of the sorted file list, balanced by
             bases; no data-path collective.
  distances  one exchange step: all-gather of the feature rows every rank produced (NCCL over
             NVLink on GPUs, gloo in the CPU tests), then rank r computes the row block [r0, r1) of
             the triangle that holds 1/world of the pairs.  Rows [r0, r1) are CONTIGUOUS in the
             packed `.mm-dist` triangle, so each rank allocates only its slice; nothing is reduced.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .dist import balanced_row_ranges


def shard_by_bases(lengths, world):
    """Contiguous shards [b, e) of the (sorted) sequence list with nearly equal total bases."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    cuts = [0]
    for p in range(1, world):
        target = total * p / world
        i = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def tri_row_offset(n, r):
    """number of packed-triangle entries in rows < r  (= index of pair (r, r+1))."""
    return r * n - r * (r + 1) // 2


def all_gather_rows(x_local, group=None):
    """All-gather row blocks of possibly different heights -> (n_total, d) on every rank.
    Rows travel as bytes (NCCL has no uint16)."""
    world = dist.get_world_size(group)
    dev = x_local.device
    cnt = torch.tensor([x_local.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    d = x_local.shape[1]
    row_bytes = d * x_local.element_size()
    cmax = max(counts)
    if min(counts) == cmax and cmax > 0:          # even shards: gather straight into the result
        full = torch.empty((world * cmax, row_bytes), dtype=torch.uint8, device=dev)
        mine = x_local.contiguous().view(torch.uint8).view(cmax, row_bytes)
        dist.all_gather_into_tensor(full, mine, group=group)
        return full.view(x_local.dtype).view(world * cmax, d), counts
    # equal-sized contributions (gloo and NCCL both take them): pad the short shards
    mine = torch.zeros((cmax, row_bytes), dtype=torch.uint8, device=dev)
    if x_local.shape[0]:
        mine[:x_local.shape[0]] = x_local.contiguous().view(torch.uint8).view(x_local.shape[0], row_bytes)
    parts = [torch.empty((cmax, row_bytes), dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    full = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return full.view(x_local.dtype).view(sum(counts), d), counts


def distances_sharded(x_local, metrics, group=None, compute=None, euclid_on_freq=True):
    """x_local: this rank's rows (count vectors).  Returns (slices, (r0, r1), n) where slices maps
    each metric to this rank's contiguous piece of the packed triangle (rows r0..r1-1).

    `compute(x_full, metric_names, (r0, r1), tri_offset, slice_len)` defaults to the CUDA kernels;
    the CPU tests inject a host function to exercise the plumbing under gloo.  euclid_on_freq: euclid
    between the frequency vectors x/sum(x) (default) or between the rows as stored (the file-based job
    step, like kam_dists_host)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x_full, _ = all_gather_rows(x_local, group)
    n = x_full.shape[0]
    r0, r1 = balanced_row_ranges(n, world)[rank]
    off0, off1 = tri_row_offset(n, r0), tri_row_offset(n, r1)
    if compute is None:
        def compute(*a):
            return _compute_cuda(*a, euclid_on_freq=euclid_on_freq)
    return compute(x_full, list(metrics), (r0, r1), off0, off1 - off0), (r0, r1), n


def _compute_cuda(x_full, metrics, rows, tri_offset, slice_len, euclid_on_freq=True):
    from . import _lib
    from . import dist as D
    lib = _lib.load()
    dev = x_full.device
    n, d = x_full.shape
    elem = D._ELEM[x_full.dtype]
    stream = torch.cuda.current_stream(dev).cuda_stream
    out = {}

    def sliced(dtype):
        t = torch.zeros(max(slice_len, 1), dtype=dtype, device=dev)
        return t, t.data_ptr() - tri_offset * 4        # kernels index the full triangle: shift the base

    with torch.cuda.device(dev):
        if "manhat" in metrics:
            t, p = sliced(torch.int32)
            ws = torch.empty(lib.kam_manhattan_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
            _lib.check(lib.kam_manhattan_tri(x_full.data_ptr(), elem, n, d, x_full.stride(0), rows[0], rows[1], p,
                                             ws.data_ptr(), ws.numel(), stream))
            out["manhat"] = t[:slice_len]
        if "info" in metrics:
            t, p = sliced(torch.float32)
            ws = torch.empty(lib.kam_info_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
            _lib.check(lib.kam_info_tri(x_full.data_ptr(), elem, n, d, x_full.stride(0), rows[0], rows[1], p,
                                        ws.data_ptr(), ws.numel(), stream))
            out["info"] = t[:slice_len]
        gram = [m for m in metrics if m in ("euclid", "cosine", "pearson")]
        if gram:
            bits = {"euclid": _lib.KAM_M_EUCLID, "cosine": _lib.KAM_M_COSINE, "pearson": _lib.KAM_M_PEARSON}
            ptrs, mask = {}, 0
            for m in gram:
                t, p = sliced(torch.float32)
                out[m], ptrs[m] = t[:slice_len], p
                mask |= bits[m]
            for size_fn in (lib.kam_gram_workspace_bytes, lib.kam_gram_workspace_bytes_int):
                ws = torch.empty(max(size_fn(n, d), 256), dtype=torch.uint8, device=dev)
                rc = lib.kam_gram_tri(x_full.data_ptr(), elem, n, d, x_full.stride(0), rows[0], rows[1], mask,
                                      int(euclid_on_freq),
                                      ptrs.get("euclid"), ptrs.get("cosine"), ptrs.get("pearson"), ws.data_ptr(),
                                      ws.numel(), stream)
                if rc != _lib.KAM_E_WORKSPACE:
                    break
                del ws
            _lib.check(rc)
    return out


def gather_triangle(slice_t, n, group=None, dst=0):
    """Collect the per-rank triangle slices on rank `dst` (for writing one .mm-dist file)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ranges = balanced_row_ranges(n, world)
    sizes = [tri_row_offset(n, b) - tri_row_offset(n, a) for a, b in ranges]
    if rank == dst:
        full = torch.empty(n * (n - 1) // 2, dtype=slice_t.dtype, device=slice_t.device)
        parts = list(full.split(sizes))
    else:
        full, parts = None, None
    # gather() wants equal sizes on some backends: use point-to-point for the ragged slices
    if rank == dst:
        parts[dst].copy_(slice_t)
        for r in range(world):
            if r != dst and sizes[r]:
                dist.recv(parts[r], src=r, group=group)
    elif sizes[rank]:
        dist.send(slice_t.contiguous(), dst=dst, group=group)
    return full


# ---- Gram metrics with block-cyclic pair ownership (ring exchange overlapped with compute) ------------
def ring_schedule(world):
    """Which shard-pair blocks each rank computes.  Rank r owns shard r and, at step s = 0..world//2,
    works on the block (shard r) x (shard (r+s) % world): s = 0 is its own triangular diagonal block;
    for even `world` the last step's block is shared by the two ranks that hold both shards, each
    taking half of the rows of the lower-numbered shard.  Returns, per rank, a list of
    (step, row_shard, col_shard, part) with part in {'tri', 'full', 'lo', 'hi'}; the global pair set
    {(a, c): a <= c} is covered exactly once."""
    plan = []
    for r in range(world):
        steps = [(0, r, r, "tri")]
        for s in range(1, world // 2 + 1):
            c = (r + s) % world
            if world % 2 == 0 and s == world // 2:
                # pair {r, c}: the rank with the smaller index takes the low half of ITS rows, the other
                # rank takes the high half of those same rows (it holds that shard in its receive buffer)
                if r < c:
                    steps.append((s, r, c, "lo"))
                else:
                    steps.append((s, c, r, "hi"))
            else:
                steps.append((s, r, c, "full"))
        plan.append(steps)
    return plan


def ring_groups(world, merge=None):
    """Steps s >= 1 of ring_schedule grouped into kernel launches.  Default: one launch per step.  With merge
    (KAM_RING_MERGE=1): the first full block alone, the remaining full blocks two at a time on a shared receive
    buffer -- fewer, larger launches lose less to wave quantisation (8 ranks x 5000 rows: 400 tiles on 74 CTA
    pairs is 5.4 -> 6 waves, 800 tiles 10.8 -> 11; measured 7.2 -> 6.6 ms for the two blocks) but must wait for
    BOTH shards: at the ~550 GB/s a rank pulls while all eight pull at once that cost 1.3 ms more than it saved
    (17.3 vs 16.3 ms per call), so it is off unless the shards are small against the blocks."""
    if merge is None:
        merge = os.environ.get("KAM_RING_MERGE", "0") == "1"
    last = world // 2
    shared = last if (world % 2 == 0 and world > 1) else None
    full = [s for s in range(1, last + 1) if s != shared]
    if merge:
        groups = [[s] for s in full[:1]] + [full[i:i + 2] for i in range(1, len(full), 2)]
    else:
        groups = [[s] for s in full]
    if shared is not None:
        groups.append([shared])
    return groups


_PEER = {}      # device index -> this rank's exported operand buffer and the peers' buffers it has opened


def _peer_state(dev):
    return _PEER.setdefault(dev.index, {"ptr": 0, "bytes": 0, "handle": b"", "opened": {}})


def _peer_buffer(lib, dev, nbytes):
    """this rank's exportable device buffer (kam_peer_alloc), grown when needed; (pointer, 64-byte handle)"""
    import ctypes
    from . import _lib
    st = _peer_state(dev)
    if st["bytes"] < nbytes:
        if st["ptr"]:
            # a peer may still hold this buffer mapped from an earlier call (it drops the mapping when it opens the
            # next handle): freeing exported memory under an open mapping is undefined, so the buffer is only
            # retired here and freed behind the closing barrier of the call (_peer_free_retired)
            st.setdefault("retired", []).append(st["ptr"])
            st["ptr"], st["bytes"] = 0, 0
        want = nbytes + nbytes // 8 + 4096
        p = ctypes.c_void_p()
        _lib.check(lib.kam_peer_alloc(want, ctypes.byref(p)))
        h = ctypes.create_string_buffer(_lib.KAM_PEER_HANDLE_BYTES)
        _lib.check(lib.kam_peer_export(p, h))
        st["ptr"], st["bytes"], st["handle"] = p.value, want, h.raw
    return st["ptr"], st["handle"]


def _peer_free_retired(lib, dev):
    """behind a call's closing barrier: every rank that reads this rank's rows has opened the new buffer (and with
    that closed its mapping of the old one), so the buffers retired by _peer_buffer can go"""
    from . import _lib
    st = _peer_state(dev)
    for p in st.pop("retired", []):
        torch.cuda.synchronize(dev)
        _lib.check(lib.kam_peer_free(p))


def _peer_open(lib, dev, src_rank, handle):
    """pointer through which this rank reads the exported buffer of `src_rank` (cached per handle)"""
    import ctypes
    from . import _lib
    st = _peer_state(dev)
    old = st["opened"].get(src_rank)
    if old is not None and old[0] == handle:
        return old[1]
    if old is not None:
        lib.kam_peer_close(old[1])            # the owner re-allocated: drop the stale mapping
    p = ctypes.c_void_p()
    _lib.check(lib.kam_peer_open(handle, ctypes.byref(p)))
    st["opened"][src_rank] = (handle, p.value)
    return p.value


def _align(x, a=256):
    return (x + a - 1) // a * a


def shared_split(n_rows):
    """Rows of the lower-numbered shard that its owner takes of a block two ranks share (fused ring): the first
    h rows -- half of them, rounded up to whole 256-row tiles so that the split is a property of the tile list
    -- the other rank takes rows [h, n_rows)."""
    return min(n_rows, ((n_rows + 1) // 2 + 255) // 256 * 256)


def fused_plan(rank, counts):
    """What rank `rank` computes in the ONE launch of the fused ring.  Its B matrix is the concatenation of its own
    rows and the rows it pulls, in the order of ring_schedule's steps.  Returns a list of
        {"src", "src_rows": (s0, s1), "b_rows": (b0, b1), "part", "a_rows": (r0, r1)}
    -- B rows [b0, b1) hold rows [s0, s1) of shard `src`; the rank computes its own rows [r0, r1) against them.
    part: 'tri' (own rows), 'full', 'lo' (rows [0, h) of mine against all of src) or 'hi' (all my rows against
    rows [h, n) of src: the other half of the block the two ranks share, held transposed)."""
    world = len(counts)
    n_loc = counts[rank]
    plan = [{"src": rank, "src_rows": (0, n_loc), "b_rows": (0, n_loc), "part": "tri", "a_rows": (0, n_loc)}]
    off = n_loc
    for s in range(1, world // 2 + 1):
        c = (rank + s) % world
        if world % 2 == 0 and s == world // 2:
            if rank < c:
                h = shared_split(n_loc)
                e = {"src": c, "src_rows": (0, counts[c]), "part": "lo", "a_rows": (0, h)}
            else:
                h = shared_split(counts[c])
                e = {"src": c, "src_rows": (h, counts[c]), "part": "hi", "a_rows": (0, n_loc)}
        else:
            e = {"src": c, "src_rows": (0, counts[c]), "part": "full", "a_rows": (0, n_loc)}
        n = e["src_rows"][1] - e["src_rows"][0]
        e["b_rows"] = (off, off + n)
        off += n
        plan.append(e)
    return plan


def fused_segments(plan):
    """Tile-list segments (column tile begin, end, row tile begin, end; tiles of 256 rows) of a fused_plan: the
    column tiles of B in arrival order, each tile in the first segment that reaches it; a tile that straddles two
    shards goes with the earlier one (it waits for the chunk that completes it).  Only the 'lo' share limits
    the rows."""
    segs, done = [], 0
    n_a = plan[0]["a_rows"][1]
    all_rows = (n_a + 255) // 256
    for e in plan:
        end = (e["b_rows"][1] + 255) // 256
        if end > done:
            rows = all_rows if e["part"] != "lo" else min(all_rows, (e["a_rows"][1] + 255) // 256)
            segs.append((done, end, 0, rows))
            done = end
    return segs


def _gather_small(t_cpu, dev, group):
    """all-gather of a few bytes per rank -> CPU tensor [world, len]; over the device for NCCL, on the host for gloo"""
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "gloo":
        parts = [torch.empty_like(t_cpu) for _ in range(world)]
        dist.all_gather(parts, t_cpu, group=group)
        return torch.stack(parts)
    out = torch.empty((world, t_cpu.numel()), dtype=t_cpu.dtype, device=dev)
    dist.all_gather_into_tensor(out, t_cpu.to(dev), group=group)
    return out.cpu()


def _gram_ring_fused(x_local, metrics, euclid_on_freq, group, timing, thermo=None):
    """gram_ring for ranks of one NVLink box, the exchange fused into the kernel: ONE kam_gram_fused launch per
    rank computes all its blocks while the copy engines pull the other shards' operand rows chunk by chunk from
    the owners' exported buffers (kam_peer_pull); every chunk is followed by a flag the kernel's TMA producers
    wait for, tile by tile.  No launch boundary between blocks (the per-step launches lost 2 of 24 waves of tiles
    to quantisation at 8 ranks x 5000 rows) and no block waits for a whole shard."""
    import ctypes
    from . import _lib
    from . import dist as D
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = x_local.device
    n_loc, d = x_local.shape
    elem = D._ELEM[x_local.dtype]
    cur = torch.cuda.current_stream(dev)
    stream = cur.cuda_stream
    dpad = int(lib.kam_gram_dpad(d))
    bits = {"euclid": _lib.KAM_M_EUCLID, "cosine": _lib.KAM_M_COSINE, "pearson": _lib.KAM_M_PEARSON}
    mask = 0
    if thermo is not None:
        metrics = (thermo,)
    for m in metrics:
        mask |= bits.get(m, 0)
    marks = []

    def mark(label):
        if timing is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(cur)
            marks.append((label, e))

    mark("start")
    # 1. who holds how many rows -> what this rank pulls, its B matrix and its tile segments
    counts = [int(v) for v in _gather_small(torch.tensor([n_loc], dtype=torch.int64), dev, group)[:, 0].tolist()]
    plan = fused_plan(rank, counts)
    n_b = plan[-1]["b_rows"][1]

    def header(n):                   # exported buffer: [S | G | operand rows ...]; offsets of G and of the rows
        hb = _align(max(n, 1) * 8, 1024)
        return hb, 2 * hb

    # 2. own operand rows + statistics into the exported buffer; the operand kind is a global decision, and the
    #    gather that takes it also carries the buffer handles and proves every owner's rows complete
    kinds = ((_lib.KAM_GRAM_U8, 1), (_lib.KAM_GRAM_F16, 2))
    if thermo is not None:
        # manhat / info: thermometer-coded rows (kam_dist_*); the layout of a manhat row follows the per-column
        # maxima over the rows of ALL ranks (one all-reduce of 4^k words), so every rank encodes alike
        if elem not in (_lib.KAM_U8, _lib.KAM_U16, _lib.KAM_U32):
            raise NotImplementedError("the tensor path of manhat / info takes uint8/16/32 counts")
        g_off, rows_off = header(n_loc)
        off_ptr, kpad = None, dpad
        if thermo == "manhat":
            colmax = torch.zeros(d, dtype=torch.int32, device=dev)
            if n_loc:
                with torch.cuda.device(dev):
                    _lib.check(lib.kam_dist_colmax(x_local.data_ptr(), elem, n_loc, d, x_local.stride(0), colmax.data_ptr(),
                                                   stream))
            if dist.get_backend(group) == "gloo":
                c = colmax.cpu()
                dist.all_reduce(c, op=dist.ReduceOp.MAX, group=group)
                colmax.copy_(c)
            else:
                dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=group)
            off = torch.empty(d + 2, dtype=torch.int32, device=dev)
            stats = torch.empty(2, dtype=torch.int64, device=dev)
            hs = (ctypes.c_uint64 * 2)()
            with torch.cuda.device(dev):
                _lib.check(lib.kam_dist_layout(colmax.data_ptr(), d, off.data_ptr(), stats.data_ptr(), hs, stream))
            top, ktot = int(hs[0]), int(hs[1])
            limit = int(os.environ.get("KAM_THERMO_MAX_AVG", "12"))
            if top > 255 or ktot >= 2 ** 31 or ktot > limit * d:
                raise NotImplementedError("manhat: counts too large for the tensor path (largest %d, column maxima sum "
                                          "to %.1f d): use distances_sharded()" % (top, ktot / d))
            off_ptr, kpad = off.data_ptr(), int(lib.kam_dist_kpad(ktot))
        base, handle = _peer_buffer(lib, dev, rows_off + max(n_b, 1) * kpad)
        if n_loc:
            enc_ws = torch.empty(n_loc * 48, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.kam_dist_encode(x_local.data_ptr(), elem, n_loc, d, x_local.stride(0), off_ptr, kpad,
                                               base + rows_off, base, base + g_off, enc_ws.data_ptr(), enc_ws.numel(),
                                               stream))
        cur.synchronize()                                             # my rows are complete
        mine = torch.zeros(24 + _lib.KAM_PEER_HANDLE_BYTES, dtype=torch.uint8)
        mine[24:] = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
        infos = _gather_small(mine, dev, group)
        kind, esz, dpad, kinds = "thermo-" + thermo, 1, kpad, ()      # (rows of kpad bytes from here on)
    for kind, esz in kinds:
        g_off, rows_off = header(n_loc)
        base, handle = _peer_buffer(lib, dev, rows_off + max(n_b, 1) * dpad * esz)
        flags = torch.zeros(8, dtype=torch.int64, device=dev)
        if n_loc:
            with torch.cuda.device(dev):
                _lib.check(lib.kam_gram_prepare(x_local.data_ptr(), elem, n_loc, d, x_local.stride(0), kind,
                                                base + rows_off, base, base + g_off, flags.data_ptr(), stream))
        mine = torch.zeros(24 + _lib.KAM_PEER_HANDLE_BYTES, dtype=torch.uint8)
        mine[:24] = flags[:3].cpu().view(torch.uint8)                 # (host sync: my rows are complete)
        mine[24:] = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
        infos = _gather_small(mine, dev, group)
        fl = infos[:, :24].contiguous().view(torch.int64).view(world, 3)
        mx, gmax, bad = int(fl[:, 0].max()), int(fl[:, 1].max()), int(fl[:, 2].max())
        if bad:
            raise ValueError("%d floating-point row(s) are not frequency vectors (count / sum): "
                             "use distances_sharded()" % bad)
        ok = (mx <= 255 and gmax < 2 ** 31) if kind == _lib.KAM_GRAM_U8 else (mx <= 2048 and gmax < 2 ** 24)
        if ok:
            break
        dist.barrier(group)          # nobody re-converts into a buffer a peer might (wrongly) still read
    else:
        if thermo is None:
            raise NotImplementedError("counts outside the tensor-exact envelope: use distances_sharded()")
    handles = [bytes(infos[r, 24:].tolist()) for r in range(world)]
    mark("prepared")

    # 3. outputs (their fill runs while the host prepares the transfers), then the transfer lists: statistics of
    #    all B rows (needed complete at launch: a few KB per shard, pulled first) and the operand chunks with
    #    their arrival flags.  Lists, flags and the statistics buffer are kept per geometry between calls.
    outs, ptrs = {}, {}
    for m in metrics:
        t = torch.empty((max(n_loc, 1), max(n_b, 1)), dtype=torch.int32 if m == "manhat" else torch.float32, device=dev)
        t[:, :n_loc].zero_()         # the triangular head keeps zeros on and below its diagonal
        outs[m], ptrs[m] = t, t.data_ptr()
    ws = torch.empty(lib.kam_gram_block_workspace_bytes(n_loc, n_b), dtype=torch.uint8, device=dev)
    row_b = dpad * esz
    tile_b = 256 * row_b
    ready_tiles = int(os.environ.get("KAM_RING_CHUNK_TILES", "0")) or max(1, (32 << 20) // tile_b)
    st = _peer_state(dev)
    if st.get("stream") is None:
        st["stream"], st["stream2"] = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    side, side2 = st["stream"], st["stream2"]
    key = (tuple(counts), tuple(handles), kind, d, dpad, base, ready_tiles, rank)
    fc = st.get("fused")
    if fc is None or fc["key"] != key:
        chunk_rows = 256 * ready_tiles
        n_chunks = (max(n_b, 1) + chunk_rows - 1) // chunk_rows
        local_chunks = n_loc // chunk_rows if n_b > n_loc else n_chunks   # chunks made of own rows only
        stats = torch.empty((2, max(n_b, 1)), dtype=torch.int64, device=dev)
        ready0 = torch.zeros(n_chunks, dtype=torch.int32, device=dev)
        ready0[:local_chunks] = 1
        pieces_stats, pieces = [], []
        S_b, G_b = stats[0].data_ptr(), stats[1].data_ptr()
        for e in plan:
            s0, s1 = e["src_rows"]
            b0 = e["b_rows"][0]
            peer = base if e["src"] == rank else _peer_open(lib, dev, e["src"], handles[e["src"]])
            g_src, rows_src = header(counts[e["src"]])
            if s1 > s0 and n_loc:    # (a rank without rows computes nothing: it only serves its peers)
                pieces_stats.append((S_b + b0 * 8, peer + s0 * 8, (s1 - s0) * 8, _lib.KAM_PEER_NO_FLAG))
                pieces_stats.append((G_b + b0 * 8, peer + g_src + s0 * 8, (s1 - s0) * 8, _lib.KAM_PEER_NO_FLAG))
            e["_peer_rows"] = peer + rows_src
        for c in range(local_chunks, n_chunks if n_loc else 0):
            lo, hi = c * chunk_rows, min((c + 1) * chunk_rows, n_b)
            for e in plan[1:]:
                b0, b1 = e["b_rows"]
                x0, x1 = max(lo, b0), min(hi, b1)
                if x1 > x0:
                    pieces.append((base + rows_off + x0 * row_b,
                                   e["_peer_rows"] + (e["src_rows"][0] + x0 - b0) * row_b, (x1 - x0) * row_b, c))
            if not pieces or pieces[-1][3] != c:
                pieces.append((0, 0, 0, c))                          # nothing to pull: the flag alone

        def as_array(plist):
            arr = (_lib.PeerPiece * max(len(plist), 1))()
            for i, (dst, src, nbytes, chunk) in enumerate(plist):
                arr[i].d_dst, arr[i].d_src, arr[i].bytes, arr[i].chunk = dst, src, nbytes, chunk
            return arr, len(plist)

        segs = fused_segments(plan)
        fc = st["fused"] = {"key": key, "stats": stats, "ready0": ready0, "ready": torch.empty_like(ready0),
                            "pieces_stats": as_array(pieces_stats), "pieces": as_array(pieces), "segs": segs,
                            "seg_arr": (ctypes.c_uint32 * (4 * len(segs)))(*[v for sg in segs for v in sg]),
                            "chunks": n_chunks - local_chunks, "chunk_bytes": chunk_rows * row_b}
    ready, segs, seg_arr = fc["ready"], fc["segs"], fc["seg_arr"]
    S_b, G_b = fc["stats"][0].data_ptr(), fc["stats"][1].data_ptr()
    ready.copy_(fc["ready0"])        # (the launch of the call before has finished: every call ends synchronised)
    for sd in (side, side2):
        sd.wait_stream(cur)          # the flags are reset, the statistics buffer exists

    def pull(pieces_n, streams):
        arr, n_p = pieces_n
        if not n_p:
            return
        sp = (ctypes.c_void_p * len(streams))(*[sd.cuda_stream for sd in streams])
        with torch.cuda.device(dev):
            _lib.check(lib.kam_peer_pull(arr, n_p, ready.data_ptr(), sp, len(streams)))

    pull(fc["pieces_stats"], [side])
    got_stats = torch.cuda.Event()
    got_stats.record(side)
    side2.wait_event(got_stats)      # (keeps the two copy streams' first chunks behind the small transfers)

    # 4. the one launch; then the chunk transfers it waits for
    mark("wait_stats")
    cur.wait_event(got_stats)
    mark("block")
    probe = None
    if timing is not None and os.environ.get("KAM_RING_PROBE", "0") == "1":
        probe = torch.zeros(4, dtype=torch.int64, device=dev)
        lib.kam_gram_set_wait_probe(probe.data_ptr())
    if n_loc and n_b and thermo is not None:
        with torch.cuda.device(dev):
            _lib.check(lib.kam_thermo_fused(base + rows_off, base, base + g_off, n_loc, base + rows_off, S_b, G_b, n_b, d,
                                            dpad, 1, seg_arr, len(segs), ready.data_ptr(), ready_tiles,
                                            _lib.KAM_M_MANHAT if thermo == "manhat" else _lib.KAM_M_INFO, ptrs[thermo],
                                            max(n_b, 1), ws.data_ptr(), ws.numel(), stream))
    elif n_loc and n_b:
        with torch.cuda.device(dev):
            _lib.check(lib.kam_gram_fused(base + rows_off, base, base + g_off, n_loc, base + rows_off, S_b, G_b, n_b, kind,
                                          d, 1, seg_arr, len(segs), ready.data_ptr(), ready_tiles, mask,
                                          int(euclid_on_freq), ptrs.get("euclid"), ptrs.get("cosine"),
                                          ptrs.get("pearson"), max(n_b, 1), 0, ws.data_ptr(), ws.numel(), stream))
    pull(fc["pieces"], [side, side2])
    mark("end")
    pulled = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] if timing is not None else None
    if pulled:
        pulled[0].record(side)
        pulled[1].record(side2)
    cur.synchronize()
    cur.wait_stream(side)
    cur.wait_stream(side2)
    if probe is not None:
        lib.kam_gram_set_wait_probe(None)
    dist.barrier(group)              # nobody overwrites or frees its exported rows while a peer still reads them
    _peer_free_retired(lib, dev)

    out_blocks = []
    for e in plan:
        b0, b1 = e["b_rows"]
        r0, r1 = e["a_rows"]
        if e["part"] == "hi":        # held transposed: pairs (row s of shard src, my row j)
            blk = {"row_shard": e["src"], "col_shard": rank, "rows": e["src_rows"], "tri": False}
            for m in metrics:
                blk[m] = outs[m][:n_loc, b0:b1].t()
        else:
            blk = {"row_shard": rank, "col_shard": e["src"], "rows": (r0, r1), "tri": e["part"] == "tri"}
            for m in metrics:
                blk[m] = outs[m][r0:r1, b0:b1]
        if blk["rows"][1] > blk["rows"][0] or e["part"] == "tri":
            out_blocks.append(blk)
    if timing is not None:
        t = {"prepare_ms": 0.0, "block_ms": 0.0, "exposed_comm_ms": 0.0, "other_ms": 0.0, "blocks_ms": []}
        for (la, ea), (_, eb) in zip(marks[:-1], marks[1:]):
            dt = ea.elapsed_time(eb)
            key = "prepare_ms" if la == "start" else "block_ms" if la == "block" else \
                "exposed_comm_ms" if la == "wait_stats" else "other_ms"
            t[key] += dt
            if la == "block":
                t["blocks_ms"].append(round(dt, 3))
        blk_ev = [e for la, e in marks if la == "block"][0]
        t["pull_done_ms_after_launch"] = [round(blk_ev.elapsed_time(pe), 3) for pe in pulled]
        if probe is not None:
            pv = probe.tolist()
            t["producer_wait"] = {"sum_ms_over_all_ctas": pv[0] / 1.965e6, "longest_ms": pv[1] / 1.965e6, "waits": pv[2]}
        t.update(sm_reserve=0, transport="peer (fused: one launch, chunk flags)",
                 operand=kind if thermo else ("uint8" if kind == _lib.KAM_GRAM_U8 else "fp16"), operand_row_bytes=int(row_b),
                 recv_bytes=int((n_b - n_loc) * row_b), chunks=int(fc["chunks"]), chunk_bytes=int(fc["chunk_bytes"]),
                 tiles_segments=[list(sg) for sg in segs])
        timing.update(t)
    return out_blocks, counts


def thermo_ring(x_local, metrics=("manhat",), group=None, timing=None):
    """`manhat` / `info` over the rows of all ranks of one NVLink box, on the tensor cores and on the schedule of
    gram_ring: every rank encodes its count rows as thermometer bytes (manhat: one layout for all ranks, from the
    all-reduced per-column maxima; info: presence bytes), the rows travel chunk by chunk from the owners' exported
    buffers while ONE fused launch per rank and metric (kam_thermo_fused) computes all its blocks.  Returns
    (blocks, counts) like gram_ring; `manhat` entries are the reference's uint32 sums (int32 storage), `info`
    float32.  Raises NotImplementedError where the tensor path does not apply (counts above 255, column maxima
    summing to more than 12 d, 64-bit or floating-point rows, no process group): use distances_sharded() then."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        raise NotImplementedError("thermo_ring needs a process group of at least two ranks")
    merged = None
    for m in metrics:
        if m not in ("manhat", "info"):
            raise ValueError("thermo_ring computes manhat and info")
        t = {} if timing is not None else None
        blocks, counts = _gram_ring_fused(x_local, (), True, group, t, thermo=m)
        if timing is not None:
            timing[m] = t
        if merged is None:
            merged = blocks
        else:
            for dst, src in zip(merged, blocks):
                assert (dst["row_shard"], dst["col_shard"], dst["rows"]) == (src["row_shard"], src["col_shard"], src["rows"])
                dst[m] = src[m]
    return merged, counts


def gram_ring(x_local, metrics=("cosine", "pearson"), euclid_on_freq=True, group=None, sm_reserve=None, timing=None,
              transport=None):
    """euclid / cosine / pearson over the rows of all ranks, block-cyclic pair ownership.

    Every rank converts its own rows once (uint8 operands when every count of every rank fits a
    byte, else fp16), then at step s fetches shard (rank+s) % world from its owner while computing
    the block of step s-1; only world//2 shards reach each rank (half an all-gather) and the first
    block needs no communication at all.  Returns a list of blocks
        {"row_shard": a, "col_shard": c, "rows": (r0, r1), "tri": bool, metric: tensor[(r1-r0), n_c]}
    where entry [i - r0, j] is the distance between row i of shard a and row j of shard c (for a
    triangular block only j > i is meaningful).

    transport: how a shard reaches the rank that needs it.
      "peer" (default; ranks of one NVLink box): the owner keeps its operand rows and row statistics in a
          buffer exported over CUDA IPC (kam_peer_*), the reader PULLS them with copy-engine transfers on
          side streams -- no SM is involved, so the persistent cluster kernel keeps all 148 SMs.  By default the
          exchange is FUSED into the kernel (_gram_ring_fused: one kam_gram_fused launch per rank, the shards
          arrive chunk by chunk while it runs, its TMA producers wait on per-chunk flags; the block two ranks
          share is split along whole 256-row tiles and the second rank's half comes back as a transposed view);
          KAM_RING_FUSED=0 keeps one launch per step, each waiting for its whole shard;
      "nccl": isend/irecv pairs; their kernels need SMs, `sm_reserve` of which (default 16) the Gram kernel
          leaves free while transfers are in flight.
    KAM_RING_TRANSPORT / KAM_RING_SM_RESERVE in the environment override the defaults.

    timing: a dict that receives, in milliseconds on this rank's compute stream, `prepare_ms` (operand
    conversion + statistics), `block_ms` (the Gram kernels) and `exposed_comm_ms` (time the compute stream
    sat waiting for a shard that had not arrived yet -- the non-overlapped part of the exchange).
    """
    from . import _lib
    from . import dist as D
    lib = _lib.load()
    multi = dist.is_available() and dist.is_initialized()
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if multi else (0, 1)
    dev = x_local.device
    n_loc, d = x_local.shape
    elem = D._ELEM[x_local.dtype]
    if x_local.dtype in (torch.float32, torch.float64):
        euclid_on_freq = True       # frequency rows: euclid between the rows as stored (kam_gram_tri does the same)
    if transport is None:
        transport = os.environ.get("KAM_RING_TRANSPORT", "peer")
    if world == 1:
        transport = "local"
    if transport == "peer" and os.environ.get("KAM_RING_FUSED", "1") == "1":
        return _gram_ring_fused(x_local, metrics, euclid_on_freq, group, timing)
    cur = torch.cuda.current_stream(dev)
    stream = cur.cuda_stream
    dpad = int(lib.kam_gram_dpad(d))
    bits = {"euclid": _lib.KAM_M_EUCLID, "cosine": _lib.KAM_M_COSINE, "pearson": _lib.KAM_M_PEARSON}
    mask = 0
    for m in metrics:
        mask |= bits[m]
    marks = []                      # (label, event) on the compute stream

    def mark(label):
        if timing is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(cur)
            marks.append((label, e))

    # 1. local operand rows + statistics, laid out [rows | S | G] in one buffer; the operand kind is a global
    #    decision (all-reduced envelope flags)
    mark("start")
    keep = []
    for kind, esz in ((_lib.KAM_GRAM_U8, 1), (_lib.KAM_GRAM_F16, 2)):
        rows_b = _align(max(n_loc, 1) * dpad * esz)
        total_b = rows_b + 2 * max(n_loc, 1) * 8
        if transport == "peer":
            base, handle = _peer_buffer(lib, dev, total_b)
        else:
            own = torch.empty(total_b, dtype=torch.uint8, device=dev)
            keep.append(own)
            base, handle = own.data_ptr(), b""
        flags = torch.zeros(8, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.kam_gram_prepare(x_local.data_ptr(), elem, n_loc, d, x_local.stride(0), kind, base,
                                            base + rows_b, base + rows_b + max(n_loc, 1) * 8, flags.data_ptr(), stream))
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
        mx, gmax, bad = (int(v) for v in flags[:3].tolist())     # (host sync: every rank's rows are complete now)
        if bad:
            raise ValueError("%d floating-point row(s) are not frequency vectors (count / sum): "
                             "use distances_sharded()" % bad)
        ok = (mx <= 255 and gmax < 2 ** 31) if kind == _lib.KAM_GRAM_U8 else (mx <= 2048 and gmax < 2 ** 24)
        if ok:
            break
    else:
        raise NotImplementedError("counts outside the tensor-exact envelope: use distances_sharded()")
    mark("prepared")

    # 2. who holds how many rows (and, for the peer transport, where)
    if world > 1:
        info = torch.zeros(_lib.KAM_PEER_HANDLE_BYTES + 8, dtype=torch.uint8)
        if handle:
            info[:_lib.KAM_PEER_HANDLE_BYTES] = torch.frombuffer(bytearray(handle), dtype=torch.uint8)
        info[_lib.KAM_PEER_HANDLE_BYTES:] = torch.tensor([n_loc], dtype=torch.int64).view(torch.uint8)
        infos = _gather_small(info, dev, group)
        counts = [int(infos[r, _lib.KAM_PEER_HANDLE_BYTES:].view(torch.int64)[0]) for r in range(world)]
        handles = [bytes(infos[r, :_lib.KAM_PEER_HANDLE_BYTES].tolist()) for r in range(world)]
    else:
        counts, handles = [n_loc], [handle]

    def layout(n):                  # offsets of S and G behind n operand rows, and the buffer's size
        rb = _align(max(n, 1) * dpad * esz)
        return rb, rb + max(n, 1) * 8, rb + 2 * max(n, 1) * 8

    # 3. start fetching the shards of the later steps, in the order they are needed.  Steps that run as ONE
    #    launch (ring_groups) share a receive buffer laid out [rows of all its shards | S | G].
    plan = {st: (a, c, part) for st, a, c, part in ring_schedule(world)[rank]}
    groups = ring_groups(world) if transport == "peer" else [[st] for st in range(1, world // 2 + 1)]
    split_pull = os.environ.get("KAM_RING_SPLIT_PULL", "1") == "1"
    bufs, ready = {}, {}
    if transport == "peer":
        side = _peer_state(dev).get("stream")
        if side is None:
            side = _peer_state(dev)["stream"] = torch.cuda.Stream(dev)
            _peer_state(dev)["stream2"] = torch.cuda.Stream(dev)
        side2 = _peer_state(dev)["stream2"]     # a second copy engine takes the upper half of every shard's rows
        for gi, g in enumerate(groups):
            srcs = [(rank + st) % world for st in g]
            n_g = sum(counts[r] for r in srcs)
            sg_, gg_, nb = layout(n_g)
            bufs[gi] = torch.empty(nb, dtype=torch.uint8, device=dev)
            off = 0
            with torch.cuda.device(dev):
                for src in srcs:
                    n_s = counts[src]
                    ss_, gs_, _ = layout(n_s)
                    peer = _peer_open(lib, dev, src, handles[src])
                    rb = n_s * dpad * esz
                    h1 = (n_s // 2) * dpad * esz if split_pull else rb
                    for dst_off, src_off, nbytes, stm in ((off * dpad * esz, 0, h1, side),
                                                          (off * dpad * esz + h1, h1, rb - h1, side2),
                                                          (sg_ + off * 8, ss_, n_s * 8, side),
                                                          (gg_ + off * 8, gs_, n_s * 8, side)):
                        if nbytes:
                            _lib.check(lib.kam_peer_copy(bufs[gi].data_ptr() + dst_off, peer + src_off, nbytes,
                                                         stm.cuda_stream))
                    off += n_s
            ev2 = torch.cuda.Event()
            ev2.record(side2)
            side.wait_event(ev2)                # the group is ready when both halves have landed
            ready[gi] = torch.cuda.Event()
            ready[gi].record(side)
    elif transport == "nccl":
        # each pair of ops is its own NCCL group, so they complete in order while the compute stream works
        for gi, g in enumerate(groups):
            src, dst = (rank + g[0]) % world, (rank - g[0]) % world
            bufs[gi] = torch.empty(layout(counts[src])[2], dtype=torch.uint8, device=dev)
            ready[gi] = dist.batch_isend_irecv([dist.P2POp(dist.isend, keep[-1][:layout(n_loc)[2]], dst, group),
                                                dist.P2POp(dist.irecv, bufs[gi], src, group)])
    elif world > 1:
        raise ValueError("unknown ring transport %r" % transport)

    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    if sm_reserve is None:
        sm_reserve = int(os.environ.get("KAM_RING_SM_RESERVE", "16"))
    if transport != "nccl":
        sm_reserve = 0              # copy-engine transfers need no SM
    out_blocks = []
    for gi, g in enumerate([[0]] + groups):
        gi -= 1                     # index into bufs / ready; -1 = the local triangular block
        a, c, part = plan[g[0]]
        # NCCL transport: SMs are left to the send/recv kernels only while transfers are still in flight
        sm_limit = max(sms - sm_reserve, 2) if (sm_reserve and gi < len(groups) - 1) else 0
        if gi >= 0:
            mark("wait%d" % g[0])
            if transport == "peer":
                cur.wait_event(ready[gi])
            else:
                for w in ready[gi]:
                    w.wait()        # the compute stream waits for the transfer, the host does not
        mark("block%d" % g[0])
        cols = [(rank + st) % world for st in g] if gi >= 0 else [rank]
        n_g = sum(counts[r] for r in cols)
        mine, theirs = (base, n_loc), ((bufs[gi].data_ptr(), n_g) if gi >= 0 else (base, n_loc))
        if part == "hi":            # A = the received shard a (rows of the shared block), B = my shard
            (pa, n_a), (pb, n_b) = theirs, mine
        else:                        # A = my shard, B = my shard (tri) or the received one(s)
            (pa, n_a), (pb, n_b) = mine, theirs
        sa, ga, _ = layout(n_a)
        sb, gb, _ = layout(n_b)
        h = n_a // 2
        r0, r1 = {"tri": (0, n_a), "full": (0, n_a), "lo": (0, h), "hi": (h, n_a)}[part]
        ptrs, outs = {}, {}
        for m in metrics:
            # every entry of a rectangular block is written by the kernel; a triangular block keeps zeros
            # on and below its diagonal
            alloc = torch.zeros if part == "tri" else torch.empty
            t = alloc((max(r1 - r0, 1), max(n_b, 1)), dtype=torch.float32, device=dev)
            outs[m] = t
            ptrs[m] = t.data_ptr() - r0 * n_b * 4          # the kernel indexes rows from 0
        ws = torch.empty(lib.kam_gram_block_workspace_bytes(n_a, n_b), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.kam_gram_block(pa, pa + sa, pa + ga, n_a, pb, pb + sb, pb + gb, n_b, kind, d,
                                          int(part == "tri"), r0, r1, mask, int(euclid_on_freq),
                                          ptrs.get("euclid"), ptrs.get("cosine"), ptrs.get("pearson"), n_b, sm_limit,
                                          ws.data_ptr(), ws.numel(), stream))
        # one block per (row shard, column shard): a merged launch is cut along its column shards
        off = 0
        for st in g:
            a_s, c_s, _ = plan[st]
            n_c = n_b if len(g) == 1 else counts[c_s]
            blk = {"row_shard": a_s, "col_shard": c_s, "rows": (r0, r1), "tri": part == "tri", "_keep": ws}
            for m in metrics:
                blk[m] = outs[m][:r1 - r0, off:off + n_c]
            off += n_c
            out_blocks.append(blk)
    mark("end")
    cur.synchronize()
    if transport == "peer":
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.current_stream(dev).wait_stream(side2)
        dist.barrier(group)         # nobody overwrites or frees its exported rows while a peer still reads them
        _peer_free_retired(lib, dev)
    for b in out_blocks:
        b.pop("_keep", None)
    if timing is not None:
        t = {"prepare_ms": 0.0, "block_ms": 0.0, "exposed_comm_ms": 0.0, "other_ms": 0.0, "blocks_ms": []}
        for (la, ea), (_, eb) in zip(marks[:-1], marks[1:]):
            dt = ea.elapsed_time(eb)
            key = "prepare_ms" if la == "start" else "block_ms" if la.startswith("block") else \
                "exposed_comm_ms" if la.startswith("wait") else "other_ms"
            t[key] += dt
            if la.startswith("block"):
                t["blocks_ms"].append(round(dt, 3))
        t["sm_reserve"] = sm_reserve
        t["transport"] = transport
        t["operand"] = "uint8" if kind == _lib.KAM_GRAM_U8 else "fp16"
        t["recv_bytes"] = int(sum(b.numel() for b in bufs.values()))
        timing.update(t)
    return out_blocks, counts
