"""kam_pack_ascii throughput (input characters per second), clean and dirty inputs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kameris_b200 import repr as R
from tests.suite import gen_chars
dev = torch.device("cuda:0")
res = {}
for name, n, L, dirty in (("genomes_clean", 40000, 9000, 0.0), ("genomes_dirty0.2%", 40000, 9000, 0.002), ("reads", 2_000_000, 150, 0.0)):
    chars = gen_chars(dev, n, L, 7)
    if dirty:
        g = torch.Generator(device=dev); g.manual_seed(1)
        idx = torch.randint(0, n * L, (int(n * L * dirty),), generator=g, device=dev)
        chars[idx] = 78   # 'N'
    start = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    for _ in range(2):
        b = R.pack_chars(chars, start, n, n * L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b = R.pack_chars(chars, start, n, n * L)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = {"ms": round(ms, 3), "GB_per_s_in": n * L / ms / 1e6}
print(json.dumps(res, indent=1))
