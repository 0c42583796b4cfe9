"""Worker of tests/test_ring_one_gpu.py (and of 2-GPU runs): every rank of a process group takes its shard of
one deterministic count matrix, runs parallel.gram_ring, and compares every entry of every block it got with
the single-GPU triangle (kam_gram_tri) of the whole matrix -- exactly (both are exact integer Grams through the
same float64 formulas).  Together the ranks must cover every pair exactly once.

    python -m torch.distributed.run --nproc-per-node W tests/ring_worker.py <backend> <k> <rows of rank 0> <rows of rank 1> ...

backend gloo: all ranks share cuda:0 (the copy-engine pulls then cross process boundaries over CUDA IPC on one
device), so the fused exchange of a 3- or 4-rank ring is exercised on a box with ONE GPU; backend nccl: one GPU
per rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import dist as D  # noqa: E402
from kameris_b200 import parallel  # noqa: E402


def main():
    backend, k = sys.argv[1], int(sys.argv[2])
    rows = [int(v) for v in sys.argv[3:]]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert len(rows) == world
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    n, d = sum(rows), 4 ** k
    rng = np.random.default_rng(20261020)
    # sparse-ish counts like k-mer vectors; a few larger values, every row non-empty
    x = rng.poisson(0.4, size=(n, d)).astype(np.uint16)
    x[:, 0] += 1
    x[rng.integers(0, n, 50), rng.integers(0, d, 50)] = 200
    x_all = torch.from_numpy(x.astype(np.int32)).to(dev).to(torch.uint16)
    offs = np.concatenate([[0], np.cumsum(rows)])
    metrics = ("cosine", "pearson", "euclid")
    timing = {}
    blocks, counts = parallel.gram_ring(x_all[offs[rank]:offs[rank + 1]].contiguous(), metrics, timing=timing)
    assert counts == rows, (counts, rows)
    ref = D.gram_tri(x_all, metrics)
    covered = torch.zeros((n, n), dtype=torch.int32, device=dev)
    ok = True
    # the reference's own metrics on the same schedule (thermometer-coded rows, kam_thermo_fused)
    if world > 1 and os.environ.get("KAM_RING_FUSED", "1") == "1":
        ref.update({"manhat": D.manhattan_tri(x_all), "info": D.info_tri(x_all)})
        tblocks, tcounts = parallel.thermo_ring(x_all[offs[rank]:offs[rank + 1]].contiguous(), ("manhat", "info"))
        assert tcounts == rows and len(tblocks) == len(blocks)
        for blk, tb in zip(blocks, tblocks):
            assert (blk["row_shard"], blk["col_shard"], blk["rows"]) == (tb["row_shard"], tb["col_shard"], tb["rows"])
            blk["manhat"], blk["info"] = tb["manhat"], tb["info"]
        metrics = metrics + ("manhat", "info")
    for blk in blocks:
        a, c = blk["row_shard"], blk["col_shard"]
        r0, r1 = blk["rows"]
        gi = torch.arange(offs[a] + r0, offs[a] + r1, device=dev).view(-1, 1)
        gj = torch.arange(offs[c], offs[c] + counts[c], device=dev).view(1, -1)
        lo, hi = torch.minimum(gi, gj), torch.maximum(gi, gj)
        idx = n * lo - lo * (lo + 3) // 2 + hi - 1
        valid = (gj > gi) if blk["tri"] else torch.ones_like(idx, dtype=torch.bool)
        for m in metrics:
            got = blk[m]
            assert tuple(got.shape) == (r1 - r0, counts[c]), (m, got.shape, blk["rows"], counts[c])
            want = ref[m][idx.clamp(0, max(ref[m].numel() - 1, 0))]
            same = torch.equal(got[valid], want[valid])
            if not same:
                bad = int(((got != want) & valid).sum())
                print("rank %d: block (%d,%d) rows %s %s: %d entries differ" % (rank, a, c, blk["rows"], m, bad), flush=True)
            ok = ok and same
        li, hj = lo.expand_as(idx)[valid], hi.expand_as(idx)[valid]
        covered.index_put_((li, hj), torch.ones_like(li, dtype=torch.int32), accumulate=True)
    if backend == "nccl":
        dist.all_reduce(covered)
    else:
        cc = covered.cpu()
        dist.all_reduce(cc)
        covered = cc.to(dev)
    iu = torch.triu(torch.ones((n, n), dtype=torch.bool, device=dev), diagonal=1)
    once = bool((covered[iu] == 1).all()) and bool((covered[~iu] == 0).all())
    if not once:
        print("rank %d: pairs covered %d..%d times" % (rank, int(covered[iu].min()), int(covered[iu].max())), flush=True)
    flag = torch.tensor([1 if (ok and once) else 0])
    if backend == "nccl":
        f = flag.to(dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        flag = f.cpu()
    else:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ring_worker world %d rows %s k %d transport %s chunks %s: %s" % (
            world, rows, k, timing.get("transport"), timing.get("chunks"), "OK" if int(flag) == 1 else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
