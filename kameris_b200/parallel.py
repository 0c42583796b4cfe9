"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the two paths shard.

  kmers      sequences are independent: contiguous shards of the sorted file list, balanced by
             bases; no data-path collective.
  distances  one exchange step: all-gather of the feature rows every rank produced (NCCL over
             NVLink on GPUs, gloo in the CPU tests), then rank r computes the row block [r0, r1) of
             the triangle that holds 1/world of the pairs.  Rows [r0, r1) are CONTIGUOUS in the
             packed `.mm-dist` triangle, so each rank allocates only its slice; nothing is reduced.
"""
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


def distances_sharded(x_local, metrics, group=None, compute=None):
    """x_local: this rank's rows (count vectors).  Returns (slices, (r0, r1), n) where slices maps
    each metric to this rank's contiguous piece of the packed triangle (rows r0..r1-1).

    `compute(x_full, metric_names, (r0, r1), tri_offset, slice_len)` defaults to the CUDA kernels;
    the CPU tests inject a host function to exercise the plumbing under gloo."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    x_full, _ = all_gather_rows(x_local, group)
    n = x_full.shape[0]
    r0, r1 = balanced_row_ranges(n, world)[rank]
    off0, off1 = tri_row_offset(n, r0), tri_row_offset(n, r1)
    fn = compute or _compute_cuda
    return fn(x_full, list(metrics), (r0, r1), off0, off1 - off0), (r0, r1), n


def _compute_cuda(x_full, metrics, rows, tri_offset, slice_len):
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
                rc = lib.kam_gram_tri(x_full.data_ptr(), elem, n, d, x_full.stride(0), rows[0], rows[1], mask, 1,
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
