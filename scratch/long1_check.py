"""scratch: one-pass path L kernel vs the tiled one -- identical vectors, timings (148 x 5 Mb, k = 7..9)."""
import sys, torch
sys.path.insert(0, ".")
from kameris_b200 import _lib, repr as R
from kameris_b200.repr_bench import _gen, _time

dev = torch.device("cuda:0")
lib = _lib.load()
sms = torch.cuda.get_device_properties(dev).multi_processor_count


def skewed(n, L, seed, probs):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    alpha = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    p = torch.tensor(probs, device=dev)
    idx = torch.multinomial(p, n * L, replacement=True, generator=g)
    out = torch.zeros(n * L + 16, dtype=torch.uint8, device=dev)
    out[:n * L] = alpha[idx]
    return out


DATASETS = [("uniform", sms, 5_000_000, lambda: _gen(dev, sms, 5_000_000, 20261005)),
                       ("gc65", sms, 5_000_000, lambda: skewed(sms, 5_000_000, 5, [0.175, 0.325, 0.325, 0.175])),
                       ("at80", 32, 5_000_000, lambda: skewed(32, 5_000_000, 6, [0.4, 0.1, 0.1, 0.4]))]
for name, n, L, mk in DATASETS[:int(sys.argv[1]) if len(sys.argv) > 1 else 3]:
    chars = mk()
    b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
    for k in (9, 8, 7):
        for kind, bits in (("freq32", 16), ("counts", 32)):
            ws = R.kmer_workspace(n, k, dev)
            res = {}
            for one in (0, 1):
                lib.kam_kmer_set_long_one_pass(one)
                o = torch.zeros((n, 4 ** k), dtype=torch.float32 if kind == "freq32" else torch.uint32, device=dev)
                t = _time(lambda: R.kmer_vectors(b, k, bits, kind, out_tensor=o, workspace=ws), 3)
                res[one] = (t, o)
            lib.kam_kmer_set_long_one_pass(1)
            same = torch.equal(res[0][1], res[1][1])
            mx = int(res[1][1].to(torch.int64).max()) if kind == "counts" else -1
            print("%-8s k=%d %-6s tiled %.3f ms  one-pass %.3f ms  (%.2fx)  %.3g k-mers/s  identical=%s max=%d" % (
                name, k, kind, res[0][0] * 1e3, res[1][0] * 1e3, res[0][0] / res[1][0], n * L / res[1][0], same, mx), flush=True)
            del o, res
    del chars, b
    torch.cuda.empty_cache()
