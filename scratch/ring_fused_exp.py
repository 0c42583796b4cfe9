"""torchrun: parallel.gram_ring on the bench's fixed 40 000 x 4^9 set: per-step launches vs the fused launch with
different chunk sizes and flag mechanisms; where the fused launch waits (KAM_RING_PROBE)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from kameris_b200 import parallel, ring_bench

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
x = ring_bench.gen_shard(dev, rank, world, n_total, 9, 2026100300)
os.environ["KAM_RING_PROBE"] = "1"
from kameris_b200 import _lib
lib = _lib.load()
variants = [("0", "0", "write", 12), ("1", "0", "write", 12), ("1", "0", "copy", 12), ("1", "0", "write", 40),
            ("1", "2", "write", 12), ("1", "0", "write", 24)]
for fused, chunk, flag, strip in variants:
    os.environ["KAM_RING_FUSED"], os.environ["KAM_RING_CHUNK_TILES"], os.environ["KAM_PEER_FLAG"] = fused, chunk, flag
    lib.kam_gram_set_strip(strip)
    parallel.gram_ring(x, ("cosine", "pearson"))
    best = None
    for _ in range(3):
        t = {}
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b, _ = parallel.gram_ring(x, ("cosine", "pearson"), timing=t); e1.record(); torch.cuda.synchronize(); del b
        v = torch.tensor([e0.elapsed_time(e1), t["block_ms"], t["exposed_comm_ms"], t["prepare_ms"], t["other_ms"]], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        if best is None or float(v[0]) < best[0]:
            best = [round(float(q), 3) for q in v] + [t]
    if rank in (0, world - 1):
        t = best[5]
        print(json.dumps({"rank": rank, "n_gpus": world, "fused": fused, "chunk_tiles": chunk, "flag": flag, "strip": strip, "ms_max": best[0],
                          "block_ms_max": best[1], "exposed_ms": best[2], "prepare_ms": best[3], "other_ms": best[4],
                          "blocks_ms": t.get("blocks_ms"), "chunks": t.get("chunks"),
                          "pull_done_ms_after_launch": t.get("pull_done_ms_after_launch"),
                          "producer_wait": t.get("producer_wait")}), flush=True)
dist.destroy_process_group()
