"""torchrun check + timing of parallel.gram_ring: every block vs the single-GPU triangle."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import parallel, dist as D
from kameris_b200.dist_bench import synth_counts

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "check"
if mode == "check":
    n, k = 1500 * world + 7, 7
    x_all = synth_counts(torch, dev, n, 3000, k, 123)
    shards = parallel.shard_by_bases([3000] * n, world)
    b, e = shards[rank]
    blocks, counts = parallel.gram_ring(x_all[b:e].contiguous(), ("cosine", "pearson", "euclid"))
    ref = D.gram_tri(x_all)
    offs = np.concatenate([[0], np.cumsum(counts)])
    ok = True
    iu = None
    for blk in blocks:
        a, c = blk["row_shard"], blk["col_shard"]
        r0, r1 = blk["rows"]
        for m in ("cosine", "pearson", "euclid"):
            got = blk[m]
            gi = torch.arange(offs[a] + r0, offs[a] + r1, device=dev).view(-1, 1)
            gj = torch.arange(offs[c], offs[c] + counts[c], device=dev).view(1, -1)
            lo, hi = torch.minimum(gi, gj), torch.maximum(gi, gj)
            idx = n * lo - lo * (lo + 3) // 2 + hi - 1
            valid = (gj > gi) if blk["tri"] else torch.ones_like(idx, dtype=torch.bool)
            want = ref[m][idx.clamp(0, ref[m].numel() - 1)]
            ok = ok and bool(torch.equal(got[valid], want[valid]))
    t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("gram_ring check:", "OK" if int(t) == 1 else "MISMATCH", "world", world, "blocks/rank", len(blocks))
else:
    N, k = int(sys.argv[2]) if len(sys.argv) > 2 else 40000, 9
    n_loc = N // world
    x_loc = synth_counts(torch, dev, n_loc, 9000, k, 1000 + rank)
    res = {}
    for reserve in (0, 16):
        def step():
            return parallel.gram_ring(x_loc, ("cosine", "pearson"), sm_reserve=reserve)
        r = step(); del r; torch.cuda.synchronize()
        times = []
        for _ in range(3):
            if world > 1: dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = step(); e1.record(); torch.cuda.synchronize(); del r
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t))
        res[reserve] = min(times)
    if rank == 0:
        n = n_loc * world
        print(json.dumps({"n_gpus": world, "N": n, "D": 4 ** k, "ring_ms_sm_reserve0": res[0], "ring_ms_sm_reserve16": res[16],
                          "pairs_per_s": n * (n - 1) / 2 / (min(res.values()) * 1e-3)}))
if world > 1:
    dist.destroy_process_group()
