"""Performance of the repr path on the other BASELINE shapes (resident planes, CUDA events)."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kameris_b200 import repr as R

dev = torch.device("cuda", 0)
PEAK = 6545.0

def planes(n, L, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    total = n * L
    d_chars = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for o in range(0, total, step):
        m = min(step, total - o)
        d_chars[o:o + m] = torch.from_numpy(alpha[rng.integers(0, 4, size=m)]).to(dev)
    start = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    b = R.pack_chars(d_chars, start, n, total)          # warm-up (allocations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b = R.pack_chars(d_chars, start, n, total); e1.record(); torch.cuda.synchronize()
    tp = e0.elapsed_time(e1) * 1e-3
    return b, tp, total

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

res = {}
# config 4: 1M x 150 bp, k=6, u16 counts
n, L, k = 1_000_000, 150, 6
b, tp, total = planes(n, L, 4)
out = torch.empty((n, 4 ** k), dtype=torch.uint16, device=dev); ws = R.kmer_workspace(n, k, dev)
t = timeit(lambda: R.kmer_vectors(b, k, 16, "counts", out_tensor=out, workspace=ws))
byt = ((L + 3) // 4 + 4 ** k * 2) * n
res["cfg4_1Mx150_k6_u16"] = {"vec_per_s": n / t, "ms": t * 1e3, "GBps": byt / t / 1e9, "frac": byt / t / 1e9 / PEAK, "pack_ms": tp * 1e3, "pack_GBps_in": total / tp / 1e9}
del out, b
# config 2 sweep: 10k x 9 kb, k=1..9 f32
n, L = 10000, 9000
b, tp, total = planes(n, L, 2)
outs = {k: torch.empty((n, 4 ** k), dtype=torch.float32, device=dev) for k in range(1, 10)}
wss = {k: R.kmer_workspace(n, k, dev) for k in range(1, 10)}
def sweep():
    for k in range(1, 10): R.kmer_vectors(b, k, 16, "freq32", out_tensor=outs[k], workspace=wss[k])
t = timeit(sweep)
byt = ((L + 3) // 4 + 4 * sum(4 ** k for k in range(1, 10))) * n
res["cfg2_sweep_k1to9_f32_per_k_launches"] = {"seq_per_s": n / t, "ms": t * 1e3, "GBps": byt / t / 1e9, "frac": byt / t / 1e9 / PEAK}
t = timeit(lambda: R.kmer_sweep(b, 1, 9, 16, "freq32", out_tensors=outs, workspace=wss[9]))
res["cfg2_sweep_k1to9_f32_fused"] = {"seq_per_s": n / t, "ms": t * 1e3, "GBps": byt / t / 1e9, "frac": byt / t / 1e9 / PEAK}
per_k = {}
for k in range(1, 10):
    tk = timeit(lambda: R.kmer_vectors(b, k, 16, "freq32", out_tensor=outs[k], workspace=wss[k]))
    per_k[k] = round(tk * 1e3, 3)
res["cfg2_per_k_ms"] = per_k
# u16 counts at k=9
out16 = torch.empty((n, 4 ** 9), dtype=torch.uint16, device=dev)
t = timeit(lambda: R.kmer_vectors(b, 9, 16, "counts", out_tensor=out16, workspace=wss[9]))
byt = ((L + 3) // 4 + 4 ** 9 * 2) * n
res["k9_u16_counts"] = {"vec_per_s": n / t, "ms": t * 1e3, "GBps": byt / t / 1e9, "frac": byt / t / 1e9 / PEAK}
out64 = torch.empty((n // 2, 4 ** 9), dtype=torch.float64, device=dev)
del outs, out16
# config 5 (scaled): 64 x 5 Mb, k=9 f32
n, L, k = 64, 5_000_000, 9
b, tp, total = planes(n, L, 5)
out = torch.empty((n, 4 ** k), dtype=torch.float32, device=dev); ws = R.kmer_workspace(n, k, dev)
t = timeit(lambda: R.kmer_vectors(b, k, 16, "freq32", out_tensor=out, workspace=ws), reps=3)
byt = ((L + 3) // 4 + 4 ** k * 4) * n
res["cfg5_64x5Mb_k9_f32"] = {"seq_per_s": n / t, "ms": t * 1e3, "GBps": byt / t / 1e9, "frac": byt / t / 1e9 / PEAK,
                             "kmers_per_s": n * L / t, "pack_ms": tp * 1e3, "pack_GBps": total / tp / 1e9}
print(json.dumps(res, indent=1))
