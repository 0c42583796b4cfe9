"""2+-GPU check of distances_sharded (run under torchrun): every rank's slice vs single-GPU result."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import parallel, dist as D
from kameris_b200.dist_bench import synth_counts

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n, k = 3000, 7
x_all = synth_counts(torch, dev, n, 3000, k, 123)            # same seed on every rank
shards = parallel.shard_by_bases([3000] * n, world)
b, e = shards[rank]
slices, (r0, r1), nn = parallel.distances_sharded(x_all[b:e].contiguous(), ["manhat", "info", "cosine", "pearson", "euclid"])
ref_m = D.manhattan_tri(x_all); ref_i = D.info_tri(x_all); ref_g = D.gram_tri(x_all)
o0, o1 = parallel.tri_row_offset(n, r0), parallel.tri_row_offset(n, r1)
ok = torch.equal(slices["manhat"], ref_m[o0:o1]) and torch.equal(slices["info"], ref_i[o0:o1])
for m in ("cosine", "pearson", "euclid"):
    ok = ok and torch.equal(slices[m], ref_g[m][o0:o1])
full = parallel.gather_triangle(slices["manhat"], n)
if rank == 0:
    ok = ok and torch.equal(full, ref_m)
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU distances_sharded:", "OK" if int(t) == 1 else "MISMATCH", "world", world, "rows", (r0, r1))
dist.destroy_process_group()
