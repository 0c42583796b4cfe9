"""Secondary benchmark lines of the `kmers` path for bench.py: the other two sequence shapes of
BASELINE.json, each on its own kernel.

  short_reads    : config 4 shape -- 1 000 000 reads x 150 bp, k=6, uint16 counts: warp-per-read kernel
                   (cgr_warp_kernel).  HBM-bound; algorithmic bytes/vector = ceil(L/4) + 4^6 * 2 = 8230.
  long_sequences : config 5 shape -- 148 sequences x 5 Mb, k=9, float32 frequencies: the one-pass path L
                   kernel (cgr_long1_kernel: the whole 4^9 table of a sequence as 6-bit fields in the shared
                   memory of one CTA, one increment per k-mer, one scan of the planes).  Bound by the
                   shared-memory atomic increment rate (~12 per clock per SM measured with
                   scratch/atoms_bench.cu), not by HBM: reported as k-mers/s against that rate, with the HBM
                   figure beside it; `tiled_ms` is the same call on the tiled uint16 kernel it replaced
                   (cgr_long_kernel, four scans per sequence).
"""
import torch


def _gen(dev, n, L, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    alpha = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=dev)
    out = torch.zeros(n * L + 16, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for o in range(0, n * L, step):
        m = min(step, n * L - o)
        out[o:o + m] = alpha[torch.randint(0, 4, (m,), generator=g, device=dev)]
    return out


def _time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def run(dev, hbm_peak):
    from . import repr as R
    out = {}
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    n, L, k = 1_000_000, 150, 6
    chars = _gen(dev, n, L, 20261004)
    b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
    o = torch.empty((n, 4 ** k), dtype=torch.uint16, device=dev)
    ws = R.kmer_workspace(n, k, dev)
    t = _time(lambda: R.kmer_vectors(b, k, 16, "counts", out_tensor=o, workspace=ws), 10)
    byt = n * ((L + 3) // 4 + 4 ** k * 2)
    out["short_reads"] = {"metric": "cgr_vectors_per_sec_k6", "value": n / t, "unit": "vectors/s", "ms": t * 1e3,
                          "config": {"workload": "1 000 000 reads x 150 bp, k=6, uint16 counts (configs[3] shape)"},
                          "roofline": {"bound": "hbm", "achieved": byt / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "frac": byt / t / 1e9 / hbm_peak, "algorithmic_bytes_per_vector": byt // n,
                                       "kernel": "cgr_warp_kernel<uint16>"}}
    del chars, b, o, ws
    torch.cuda.empty_cache()

    n, L, k = sms, 5_000_000, 9
    chars = _gen(dev, n, L, 20261005)
    b = R.pack_chars(chars, torch.arange(n + 1, dtype=torch.int64, device=dev) * L, n, n * L)
    o = torch.empty((n, 4 ** k), dtype=torch.float32, device=dev)
    ws = R.kmer_workspace(n, k, dev)
    t = _time(lambda: R.kmer_vectors(b, k, 16, "freq32", out_tensor=o, workspace=ws), 5)
    from . import _lib
    _lib.load().kam_kmer_set_long_one_pass(0)
    try:
        t_tiled = _time(lambda: R.kmer_vectors(b, k, 16, "freq32", out_tensor=o, workspace=ws), 3)
    finally:
        _lib.load().kam_kmer_set_long_one_pass(1)
    byt = n * ((L + 3) // 4 + 4 ** k * 4)
    atom_peak = sms * 12.0 * 1.965e9            # measured shared-memory increment rate (scratch/atoms_bench.cu)
    out["long_sequences"] = {"metric": "kmers_per_sec_k9", "value": n * L / t, "unit": "k-mers/s", "ms": t * 1e3,
                             "sequences_per_s": n / t, "tiled_ms": t_tiled * 1e3,
                             "config": {"workload": "%d sequences x 5 Mb, k=9, float32 frequencies (configs[4] shape)" % n,
                                        "includes": "path A pass that routes every sequence to path L, the zeroing of "
                                                    "the side tables, one read-back of the overflow lists"},
                             "roofline": {"bound": "shared-memory atomics", "achieved": n * L / t / 1e9,
                                          "peak": atom_peak / 1e9, "unit": "G increments/s",
                                          "frac": n * L / t / atom_peak, "kernel": "cgr_long1_kernel<float, 6-bit fields>",
                                          "hbm_GBps": byt / t / 1e9, "hbm_frac": byt / t / 1e9 / hbm_peak}}
    return out
