"""ring-scheduled Gram under torchrun: whole-call time and its parts for several SM reservations
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scratch/ring_exp.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from kameris_b200 import parallel, ring_bench

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
x = ring_bench.gen_shard(dev, rank, world, ring_bench.N_TOTAL, 9, 2026100300)
parallel.gram_ring(x, ("cosine", "pearson"))
for tr, res, split in ([("local", 0, "1")] if world == 1 else [("peer", 0, "1"), ("peer", 0, "0"), ("peer", 0, "1")]):
    os.environ["KAM_RING_SPLIT_PULL"] = split
    best = None
    for rep in range(3):
        t = {}
        ring_bench._barrier(dist, dev, world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blocks, _ = parallel.gram_ring(x, ("cosine", "pearson"), sm_reserve=res, timing=t, transport=tr)
        e1.record()
        ring_bench._barrier(dist, dev, world)
        del blocks
        ms = ring_bench._max_over_ranks(dist, dev, world, [e0.elapsed_time(e1)])[0]
        if best is None or ms < best[0]:
            best = (ms, t)
    if rank == 0:
        print(json.dumps({"world": world, "split_pull": split, "ms": best[0], **best[1]}), flush=True)
if world > 1:
    dist.destroy_process_group()
