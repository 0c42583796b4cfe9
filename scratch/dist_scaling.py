"""Strong scaling of the distances step at fixed N (torchrun, 1/2/4/8 GPUs): all-gather + cosine/pearson."""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import parallel, dist as D
from kameris_b200.dist_bench import synth_counts

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, k = 40000, 9
n_loc = N // world
x_loc = synth_counts(torch, dev, n_loc, 9000, k, 1000 + rank)
def step():
    if world > 1:
        x_full, _ = parallel.all_gather_rows(x_loc)
    else:
        x_full = x_loc
    r0, r1 = D.balanced_row_ranges(N, world)[rank]
    o0, o1 = parallel.tri_row_offset(N, r0), parallel.tri_row_offset(N, r1)
    return parallel._compute_cuda(x_full, ["cosine", "pearson"], (r0, r1), o0, o1 - o0)
step(); torch.cuda.synchronize()
times = []
for _ in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = step(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t)); del r
if rank == 0:
    t = min(times) * 1e-3
    print(json.dumps({"n_gpus": world, "N": N, "D": 4 ** k, "step_ms": t * 1e3, "pairs_per_s": N * (N - 1) / 2 / t,
                      "what": "all-gather of uint16 rows + cosine+pearson on this rank's triangle row block"}))
if world > 1: dist.destroy_process_group()
