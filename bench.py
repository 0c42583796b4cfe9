#!/usr/bin/env python3
"""bench.py -- headline benchmark of kameris_b200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic HIV-1-like set, 10 000 sequences x 9 kb, k=9 CGR
frequency vectors (4^9 bins) -- the configuration "CGR vectors/sec (k=9)" is quoted on.
One step = one pass of the `kmers` hot path over the whole batch.

  value : vectors/s of the device path, inputs resident in HBM: the ASCII records -> 2-bit planes
          (kam_pack_ascii) -> float32 frequency vectors (kam_kmer_count), both inside the timed step
  roofline : HBM, for the dominant kernel (cgr_tile_kernel, timed with its own CUDA events inside the
          steps); algorithmic bytes/vector = ceil(L/4) + 4^k*4 against MEASURED_PEAKS.json.
          roofline.also carries the other kernels' figures (Gram uint8 / fp16 / D=4096 / limb-split, manhattan
          and info on the CUDA cores AND on the tensor cores (thermometer-coded rows), float L1, short reads,
          megabase sequences, pack, and the multi-GPU rings)
  e2e   : vectors/s of the whole `kmers` step in the reference's default mode (frequencies, as in
          demo/hiv1-lanl.yml) through kam_kmers_host with HOST buffers: ASCII records in pinned host memory
          -> H2D -> pack -> count -> float64 frequency vectors (the reference's .mm-repr value type, bit-exact
          cgr/cgr.sum()) in host memory.  e2e.also: `mode: counts` (uint16), float32, the dense-copy
          variants, the distances step (kam_dists_host), the fused kmers -> distances pipeline
  cpu_baseline : the reference `kmers` step on all host cores, on a bounded sample: generation_cgr's
          per-file CGR fill -- the reference's OWN machine code (`kind: reference`: the leaf function of
          generation_cgr_windows_avx2.exe, extracted into oracle/_ref by the build where /root/reference
          is mounted and run in-process, oracle/refcode/refcode.c) or, without it, the oracle's threaded
          restatement (`kind: port`) -- followed by the reference's own single-threaded numpy pass
          (backend.py:45-49); `counts_only` = the step in `mode: counts` (no numpy pass).

With --gpus N (torchrun, one rank per GPU):
  * the `kmers` legs are WEAK scaling: every rank its own 10 000-sequence shard (sequences are independent:
    no data-path collective); value = all ranks' vectors / max-over-ranks time;
  * the `distances` legs are STRONG scaling on a fixed set (kameris_b200/ring_bench.py): 40 000 x 4^9 count
    vectors, cosine+pearson by the ring-scheduled tcgen05 Gram with the exchange fused into ONE launch per rank
    (operand chunks pulled from peer memory by the copy engines, arrival flags awaited by the TMA producers),
    the reference's own `manhat` of the same set on the same schedule (thermometer-coded rows), and manhat of
    32 768 x 4^6 vectors by all-gather + balanced triangle row blocks -- reported under roofline.also.gram_ring /
    .manhat_ring / .manhat_sharded at EVERY N (1 included), so the scaling record covers the path that has an
    exchange step;
  * for N > 1 the sharded job steps are checked byte for byte against one process (config.multi_gpu_self_check).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEQS, SEQ_LEN, K, SEED = 10000, 9000, 9, 20261002
HIV_P = np.array([0.36, 0.18, 0.24, 0.22])
BINS = 4 ** K
ALG_BYTES_PER_VEC = (SEQ_LEN + 3) // 4 + BINS * 4          # 2250 + 1 048 576
WORKLOAD = "synthetic HIV-1-like 10k x 9 kb, k=9 CGR frequency vectors (configs[1])"


def synth_chars(n, length, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    chars = alpha[rng.choice(4, size=n * length, p=HIV_P)]
    start = (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return chars, start


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- the reference's CPU path (cpu_baseline leg and --impl reference) -----------------------------------------
def reference_fill(O):
    """The reference's own CGR-fill machine code when oracle/_ref holds it (built in the container where
    /root/reference is mounted; see oracle/refcode/refcode.c), checked here against the port on a few
    records; None -> the port is timed instead."""
    try:
        if O.refcode() is None:
            return None
        chars, start = synth_chars(8, SEQ_LEN, SEED + 1)
        got = np.zeros((8, BINS), dtype=np.uint16)
        O.refcode_cgr_batch(chars, start, K, got, threads=2)
        if not np.array_equal(got, O.cgr_batch(chars, start, K, 16)):
            print("bench: the reference's machine code and its port disagree; timing the port", file=sys.stderr)
            return None
        return O.refcode_cgr_batch
    except (OSError, RuntimeError) as e:      # a baseline problem must not cost the bench line
        print("bench: reference machine code unavailable (%s); timing the port" % e, file=sys.stderr)
        return None


def reference_counts(O, chars, start, n, counts, threads, fill=None):
    if fill is not None:
        fill(chars, start, K, counts, threads=threads)
    else:
        O.lib().orc_cgr_batch(chars.ctypes.data, start.ctypes.data, n, K, 16, counts.ctypes.data, threads)


def reference_kmers_step(O, chars, start, n, counts, threads, fill=None):
    """The reference `kmers` step, mode=frequencies, minus file I/O: generation_cgr's per-file CGR fill on
    all host threads (the reference's own machine code when available, else the threaded port; uint16
    counts) then the numpy loop of kameris/job_steps/backend.py:45-49 verbatim."""
    reference_counts(O, chars, start, n, counts, threads, fill)
    cgrs = []
    for i in range(n):
        cgr = counts[i].reshape(1, -1)          # reader.read_matrix(i)
        cgrs.append(cgr / cgr.sum())
    return cgrs


def _baseline_kind(fill, threads):
    if fill is not None:
        return "reference", ("the reference's own CGR-fill machine code (generation_cgr_windows_avx2.exe VA "
                             "0x140016040, mapped from oracle/_ref) on %d threads" % threads)
    return "port", "threaded uint16 counts (port of generation_cgr, %d threads)" % threads


def cpu_baseline(sample_n=1500, reps=2):
    from oracle import oracle as O
    threads = O.host_threads()
    fill = reference_fill(O)
    kind, what = _baseline_kind(fill, threads)
    n = sample_n
    chars, start = synth_chars(n, SEQ_LEN, SEED)
    counts = np.zeros((n, BINS), dtype=np.uint16)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        reference_kmers_step(O, chars, start, n, counts, threads, fill)
        best = min(best, time.perf_counter() - t0)
    t_fill = float("inf")                             # the step in `mode: counts`: no numpy pass, warm buffers
    for _ in range(3):
        t0 = time.perf_counter()
        reference_counts(O, chars, start, n, counts, threads, fill)
        t_fill = min(t_fill, time.perf_counter() - t0)
    n2 = max(64, 250 * threads)
    chars2, start2 = synth_chars(n2, SEQ_LEN, SEED)
    out = np.zeros((n2, BINS), dtype=np.float32)     # pre-touched
    best2 = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        O.cgr_freq_batch(chars2, start2, K, 16, f64=False, threads=threads, out=out)
        best2 = min(best2, time.perf_counter() - t0)
    return {"value": n / best, "unit": "vectors/s", "cores": threads, "kind": kind,
            "sample": "%d of the %d sequences x %d bp, k=%d: %s + the reference's single-threaded numpy float64 "
                      "pass (backend.py:45-49), best of %d" % (n, N_SEQS, SEQ_LEN, K, what, reps),
            "counts_only": {"value": n / t_fill, "unit": "vectors/s",
                            "what": "the step in `mode: counts` (%s): the uint16 counting alone, no frequency pass, "
                                    "best of 3" % kind},
            "threaded_c_f32": {"value": n2 / best2, "unit": "vectors/s",
                               "what": "counts + float32 normalisation fused in C on all %d threads, output "
                                       "pre-touched (NOT what the reference does; upper bound for a CPU)" % threads}}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores -- its own CGR-fill machine code from
    oracle/_ref when the build container extracted it, else the oracle port -- plus its numpy pass.  A step is
    a bounded sample of the 10 000-sequence workload: as many sequences as keep the whole run (warm-up + steps)
    near BUDGET_S seconds at the rate measured on a first 250-sequence probe, at most all 10 000."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.host_threads()
    fill = reference_fill(O)
    kind, what = _baseline_kind(fill, threads)
    BUDGET_S = 150.0
    probe = 250
    chars, start = synth_chars(probe, SEQ_LEN, SEED)
    counts = np.zeros((probe, BINS), dtype=np.uint16)
    reference_kmers_step(O, chars, start, probe, counts, threads, fill)          # touches the buffers
    t0 = time.perf_counter()
    reference_kmers_step(O, chars, start, probe, counts, threads, fill)
    rate = probe / (time.perf_counter() - t0)
    n = int(min(N_SEQS, max(250, rate * BUDGET_S / max(args.steps + args.warmup, 1))))
    n -= n % 50
    chars, start = synth_chars(n, SEQ_LEN, SEED)
    counts = np.zeros((n, BINS), dtype=np.uint16)
    for _ in range(args.warmup):
        reference_kmers_step(O, chars, start, n, counts, threads, fill)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_kmers_step(O, chars, start, n, counts, threads, fill)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = ("%d sequences x %d bp per step -- a bounded sample of the %d-sequence workload sized so that %d warm-up + "
              "%d timed steps take about %d s (the rate does not depend on the sample size: every sequence is "
              "independent work): %s + single-threaded numpy float64 pass (backend.py:45-49)"
              % (n, SEQ_LEN, N_SEQS, args.warmup, args.steps, BUDGET_S, what))
    print(json.dumps({
        "impl": "reference", "metric": "cgr_vectors_per_sec_k9", "value": v, "unit": "vectors/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_seqs_per_step": n, "seq_len": SEQ_LEN, "k": K, "out": "float64"},
        "cpu_baseline": {"value": v, "unit": "vectors/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def mds_cpu_baseline(tri, n, dim, sample_n=3000):
    """the reference's mds() (numpy + ARPACK, oracle port of mds.py:10-35) on the first sample_n points of the
    set the GPU line used"""
    from oracle import oracle as O
    m = min(sample_n, n)
    # pairs (i, j) with i < j < m: the first m-1-i entries of row i of the packed triangle
    sub = np.concatenate([tri[i * n - i * (i + 1) // 2: i * n - i * (i + 1) // 2 + (m - 1 - i)] for i in range(m - 1)])
    delta = O.full_from_triangle(sub, m)
    t0 = time.perf_counter()
    O.mds_reference_arpack(delta, dim)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": "points/s", "kind": "port", "sample": "first %d of the %d points: numpy centring + "
            "scipy eigsh (mds.py:10-35)" % (m, n), "ms": dt * 1e3}


def distances_cpu_baseline(n=4096, k=6, length=9000):
    """generation_dists' row tasks on all host threads, on n count vectors of D = 4^k: the reference's OWN
    row-lambda machine code (`kind: reference`: generation_dists_windows_avx2.exe VA 0x1400253c0, run from the
    image the build extracted into oracle/_ref, oracle/refcode/refdists.c; checked here against the port) or,
    without it, the threaded restatement (`kind: port`); and scipy's pdist for the metrics the reference does
    not have."""
    from oracle import oracle as O
    threads = O.host_threads()
    chars, start = synth_chars(n, length, SEED + 7)
    x = O.cgr_batch(chars, start, k, 16)
    pairs = n * (n - 1) // 2
    kind, run = "port", lambda m, i: O.dists_triangle(x, m, i, threads=threads)
    try:
        if O.refdists() is not None:
            a = O.refdists_triangle(x[:64], True, True, threads=2)
            b = O.dists_triangle(x[:64], True, True)
            if np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)):
                kind, run = "reference", lambda m, i: O.refdists_triangle(x, m, i, threads=threads)
            else:
                print("bench: the reference's row-task machine code and its port disagree; timing the port",
                      file=sys.stderr)
    except (OSError, RuntimeError) as e:      # a baseline problem must not cost the bench line
        print("bench: reference machine code unavailable (%s); timing the port" % e, file=sys.stderr)
    t0 = time.perf_counter()
    run(True, False)
    tm = time.perf_counter() - t0
    t0 = time.perf_counter()
    run(False, True)
    ti = time.perf_counter() - t0
    m = 256
    t0 = time.perf_counter()
    O.cdist_triangle(x[:m], "cosine")
    tc = time.perf_counter() - t0
    sample = "%d count vectors (%d bp, k=%d, D=%d): %d pairs" % (n, length, k, 4 ** k, pairs)
    return {"manhat": {"value": pairs / tm, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
            "info": {"value": pairs / ti, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
            "cosine_scipy_pdist": {"value": m * (m - 1) // 2 / tc, "unit": "pairs/s", "cores": 1, "kind": "port",
                                   "sample": "scipy.spatial.distance.pdist(cosine) on %d of those vectors (the "
                                             "definition of the metric; not in the reference)" % m}}


def preprocess_cpu_baseline(host, sample_n=2000):
    """the reference's two normalisers (classify.py:29-50) as scikit-learn runs them on the host cores
    (`kind: reference`: scikit-learn IS the reference's implementation of this arithmetic), on the first
    sample_n vectors of the set the GPU line used"""
    from sklearn.decomposition import TruncatedSVD
    from sklearn.preprocessing import StandardScaler
    x = host[:sample_n].astype(np.float64)
    ncomp = int(np.ceil(sum(np.count_nonzero(f) for f in x[:256]) / 256 * 0.1))
    t0 = time.perf_counter()
    TruncatedSVD(n_components=ncomp, random_state=0).fit_transform(StandardScaler(with_mean=False).fit_transform(x))
    dt = time.perf_counter() - t0
    return {"value": len(x) / dt, "unit": "vectors/s", "kind": "reference", "cores": len(os.sched_getaffinity(0)),
            "ms": dt * 1e3, "sample": "first %d of the %d vectors: sklearn StandardScaler(with_mean=False) + "
            "TruncatedSVD(n_components=%d) fit_transform" % (len(x), len(host), ncomp)}


def guarded(fn, *a, **kw):
    """a secondary measurement must never cost the headline line"""
    try:
        return fn(*a, **kw)
    except Exception as e:   # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def compact(d, keys=("value", "unit", "ms", "achieved", "peak", "frac", "kernel", "bound", "error")):
    """the figures of a secondary line that go under roofline.also / e2e.also"""
    if not isinstance(d, dict):
        return d
    src = dict(d)
    src.update(d.get("roofline", {}))
    return {k: src[k] for k in keys if k in src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-distances", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short-read / long-sequence / mds / classify lines")
    ap.add_argument("--no-ring", action="store_true", help="skip the multi-GPU distances legs and the self-check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # ONE JSON line on stdout: libraries that print to file descriptor 1 (MAGMA / NCCL banners) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args, rank, world)
        sys.stdout.flush()
        return

    import torch
    import torch.distributed as dist_pg
    from kameris_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep NCCL's version banner off stdout (ONE JSON line)
        dist_pg.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    hbm_peak, tf_peak, peak_src = load_peaks()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def barrier():
        if world > 1:
            dist_pg.barrier()
        torch.cuda.synchronize(dev)

    # ---- inputs: every rank its own shard (weak scaling), ASCII resident in HBM -------------------------------
    chars, start = synth_chars(N_SEQS, SEQ_LEN, SEED + rank)
    nchars = len(chars)
    h_chars = torch.from_numpy(chars.copy()).pin_memory()
    d_chars = torch.zeros(nchars + 16, dtype=torch.uint8, device=dev)
    d_chars[:nchars] = h_chars.to(dev)
    d_start = torch.from_numpy(start.astype(np.int64)).to(dev)
    planes = torch.empty(lib.kam_planes_words(nchars), dtype=torch.int64, device=dev)
    base = torch.empty(N_SEQS + 1, dtype=torch.int64, device=dev)
    head = torch.empty(N_SEQS, dtype=torch.int32, device=dev)
    pws = torch.empty(max(lib.kam_pack_workspace_bytes(nchars, N_SEQS), 256), dtype=torch.uint8, device=dev)
    kws = torch.empty(lib.kam_kmer_workspace_bytes(N_SEQS, K), dtype=torch.uint8, device=dev)
    out = torch.empty((N_SEQS, BINS), dtype=torch.float32, device=dev)      # 10.5 GB
    PACK_LAUNCHES, COUNT_LAUNCHES = 3, 1            # tile_first_seq + pack_fused + head_mask; cgr_tile_kernel

    def pack():
        _lib.check(lib.kam_pack_ascii(d_chars.data_ptr(), d_start.data_ptr(), N_SEQS, nchars, planes.data_ptr(),
                                      base.data_ptr(), head.data_ptr(), pws.data_ptr(), pws.numel(), stream))

    def count():
        _lib.check(lib.kam_kmer_count(planes.data_ptr(), base.data_ptr(), head.data_ptr(), N_SEQS, nchars, K, 16,
                                      _lib.KAM_OUT_FREQ_F32, out.data_ptr(), BINS, kws.data_ptr(), kws.numel(), stream))

    for _ in range(args.warmup):
        pack()
        count()
    # sanity before timing: row sums of counts == L-k+1  <=>  frequencies sum to 1
    s = out[:64].sum(dim=1)
    assert torch.allclose(s, torch.ones_like(s), atol=1e-4), "frequency rows do not sum to 1"

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        pack()
        ev[2 * i + 1].record()
        count()
        ev[2 * i + 2].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]
    pack_ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    del out, planes, base, head, pws, kws, d_chars
    torch.cuda.empty_cache()

    # ---- e2e: host ASCII -> host vectors through the C ABI ------------------------------------------------------
    # the step's 10 000 vectors (21 GB of float64) return through a reused pinned buffer, SUB rows a call
    SUB = 2500
    h_out = torch.empty((SUB, BINS), dtype=torch.float64).pin_memory()
    h_start_u64 = start.astype(np.uint64)
    sub_starts = [(h_start_u64[i:i + SUB + 1] - h_start_u64[i]).copy() for i in range(0, N_SEQS, SUB)]
    KINDS = {"f64": (_lib.KAM_OUT_FREQ_F64, 8), "f32": (_lib.KAM_OUT_FREQ_F32, 4), "u16": (_lib.KAM_OUT_COUNTS, 2)}

    def e2e_step(kind):
        code = KINDS[kind][0]                            # every kind reuses the one pinned buffer (rows packed)
        for bi, i in enumerate(range(0, N_SEQS, SUB)):
            _lib.check(lib.kam_kmers_host(h_chars.data_ptr() + int(h_start_u64[i]), sub_starts[bi].ctypes.data,
                                          min(SUB, N_SEQS - i), K, 16, code, h_out.data_ptr()))

    def e2e_time(kind, mode, steps):
        lib.kam_kmers_host_set_sparse(mode)
        try:
            e2e_step(kind)                               # warm-up: allocates the pipeline's buffers
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step(kind)
            barrier()
            dt = (time.perf_counter() - t0) / steps
        finally:
            lib.kam_kmers_host_set_sparse(1)
        if kind == "f64":
            chk = float(h_out[SUB // 2].sum())
            assert abs(chk - 1.0) < 1e-9, "e2e output row does not sum to 1"
        elif kind == "u16":
            r = SUB // 2
            chk = int(h_out.view(torch.int16).reshape(-1)[r * BINS:(r + 1) * BINS].to(torch.int64).sum())
            assert chk == SEQ_LEN - K + 1, "e2e count row does not sum to L-k+1"
        return dt

    e2e_s = e2e_time("f64", 1, args.e2e_steps)
    e2e_counts_s = e2e_time("u16", 1, args.e2e_steps)
    variants = {}
    if world == 1:
        variants["float32"] = e2e_time("f32", 1, 1)
        variants["float64_dense_copy"] = e2e_time("f64", 0, 1)
        variants["counts_dense_copy"] = e2e_time("u16", 0, 1)
    host_threads = int(lib.kam_host_threads())
    del h_out
    # ---- max over ranks -----------------------------------------------------------------------------------
    t = torch.tensor([total_ms, e2e_s, e2e_counts_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist_pg.all_reduce(t, op=dist_pg.ReduceOp.MAX)
    total_ms, e2e_s, e2e_counts_s = float(t[0]), float(t[1]), float(t[2])

    # ---- multi-GPU distances legs (every N, also 1) and the sharded-steps self-check (N > 1) -----------------
    also, e2e_also, selfcheck = {}, {}, None
    if not args.no_ring:
        from kameris_b200 import ring_bench
        ring_sampler = ClockSampler(dev.index)                     # a 100-200 ms tensor-core launch: was it power-capped?
        if rank == 0:
            ring_sampler.start()
        also["gram_ring"] = guarded(ring_bench.run_gram_ring, dist_pg, dev, rank, world, tf_peak)
        if rank == 0 and isinstance(also["gram_ring"], dict):
            also["gram_ring"]["clocks"] = ring_sampler.stop()
        torch.cuda.empty_cache()
        also["manhat_ring"] = guarded(ring_bench.run_manhat_ring, dist_pg, dev, rank, world, tf_peak)
        torch.cuda.empty_cache()
        also["manhat_sharded"] = guarded(ring_bench.run_manhat_sharded, dist_pg, dev, rank, world)
        torch.cuda.empty_cache()
        if world > 1:
            from kameris_b200 import selfcheck as SC
            selfcheck = guarded(SC.run, dist_pg, rank, world)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * N_SEQS / (ms_per_step * 1e-3)
        k_ms = float(np.mean(kernel_ms))
        p_ms = float(np.mean(pack_ms))
        achieved = ALG_BYTES_PER_VEC * N_SEQS / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes per launch from the ncu capture
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("cgr_tile_kernel_f32_k9_bytes_per_launch")
        pack_bytes = nchars + nchars // 4
        also["pack"] = {"kernel": "pack_fused_kernel (+ tile_first_seq, head_mask)", "ms": p_ms, "bound": "issue",
                        "achieved": pack_bytes / (p_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": pack_bytes / (p_ms * 1e-3) / 1e9 / hbm_peak}
        h2d = int(nchars + (N_SEQS + 1) * 8)
        entry_bytes = int(nchars * 4 + N_SEQS * 4)

        def e2e_line(sec, out_bytes, sparse):
            return {"value": world * N_SEQS / sec, "unit": "vectors/s", "ms": sec * 1e3, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": entry_bytes if sparse else out_bytes, "host_bytes_written_per_step": out_bytes}
        e2e_also["counts_mode"] = dict(e2e_line(e2e_counts_s, N_SEQS * BINS * 2, True),
                                       out="uint16 counts (`mode: counts`, bits_per_element 16), bit-exact")
        if "float32" in variants:
            e2e_also["float32"] = e2e_line(variants["float32"], N_SEQS * BINS * 4, True)
            e2e_also["float64_dense_copy"] = dict(e2e_line(variants["float64_dense_copy"], N_SEQS * BINS * 8, False),
                                                  what="kam_kmers_host_set_sparse(0): dense rows over PCIe (round 1 path)")
            e2e_also["counts_dense_copy"] = e2e_line(variants["counts_dense_copy"], N_SEQS * BINS * 2, False)
        line = {
            "metric": "cgr_vectors_per_sec_k9", "value": value, "unit": "vectors/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_seqs_per_gpu": N_SEQS, "seq_len": SEQ_LEN, "k": K, "bins": BINS,
                       "step": "ASCII records resident in HBM -> kam_pack_ascii -> kam_kmer_count (float32 frequencies)",
                       "out": "float32 on device for value/roofline (SURVEY.md 8d); e2e returns float64",
                       "l2": "each step writes 10.5 GB of vectors (>> 126 MB L2), nothing is re-read",
                       "parallelism": "kmers: sequence shards, no collective (weak); distances legs under "
                                      "roofline.also: fixed set, ring / all-gather (strong)",
                       "host_threads_per_rank": host_threads},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "kernel": "cgr_tile_kernel<uint8,float>",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_vector": ALG_BYTES_PER_VEC,
                         "peak_source": peak_src, "also": also},
            "e2e": {"value": world * N_SEQS / e2e_s, "unit": "vectors/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": entry_bytes, "ms_per_step": e2e_s * 1e3,
                    "host_bytes_written_per_step": int(N_SEQS * BINS * 8),
                    "out": "float64 frequencies, bit-exact cgr/cgr.sum() (the reference's .mm-repr value type): "
                           "%.1f GB of dense rows written into the caller's host buffer per step" % (N_SEQS * BINS * 8 / 1e9),
                    "how": "the kernel returns each vector as a sorted (bin, count) list (<= 4 B per base); only the "
                           "lists cross PCIe; %d host threads expand them into the dense rows with non-temporal "
                           "stores and do the IEEE division.  Bound: host DRAM write bandwidth" % host_threads,
                    "host_write_GBps": world * N_SEQS * BINS * 8 / e2e_s / 1e9,
                    "api": "kam_kmers_host (C ABI, pinned host buffers), %d vectors per call" % SUB,
                    "also": e2e_also},
            "gpu_launches": args.steps * (PACK_LAUNCHES + COUNT_LAUNCHES),
            "clocks": clocks,
            "host": {"cores": len(os.sched_getaffinity(0))},
        }
        if selfcheck is not None:
            line["config"]["multi_gpu_self_check"] = selfcheck
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = guarded(cpu_baseline)
            co = line["cpu_baseline"].get("counts_only", {}).get("value")
            if co:
                e2e_also["counts_mode"]["vs_reference_counts_only"] = e2e_also["counts_mode"]["value"] / co
        torch.cuda.empty_cache()
        if world == 1 and not args.no_extra:
            from kameris_b200 import repr_bench
            shapes = guarded(repr_bench.run, dev, hbm_peak)
            line["kmers_other_shapes"] = shapes
            for key in ("short_reads", "long_sequences"):
                if key in shapes:
                    also[key] = compact(shapes[key])
            if "error" in shapes:
                also["kmers_other_shapes"] = shapes
            torch.cuda.empty_cache()
        if world == 1 and not args.no_distances:
            from kameris_b200 import dist_bench
            d = guarded(dist_bench.run, torch, dev, hbm_peak, tf_peak)
            line["distances"] = d
            if "error" in d:
                also["distances"] = d
            else:
                for key, val in d.get("rooflines", {}).items():
                    also[key] = val
                if "pipeline" in d:
                    e2e_also["kmers_dists_pipeline"] = compact(d["pipeline"], ("value", "unit", "ms", "h2d_bytes_per_step",
                                                                                "d2h_bytes_per_step"))
                for key, val in d.get("e2e", {}).items():
                    if isinstance(val, dict) and "value" in val:
                        e2e_also["dists_host_" + key] = compact(val, ("value", "unit", "ms", "h2d_bytes_per_step",
                                                                      "d2h_bytes_per_step"))
                if not args.no_cpu_baseline:
                    d["cpu_baseline"] = guarded(distances_cpu_baseline)
                    for key in ("manhat", "info"):
                        if key in d["cpu_baseline"]:
                            line["cpu_baseline"]["distances_" + key] = d["cpu_baseline"][key]
        if world == 1 and not args.no_extra:
            from kameris_b200 import mds_bench, preprocess_bench
            torch.cuda.empty_cache()
            try:                                   # the first consumer of the matrices
                line["mds"], tri_sample = mds_bench.run(dev, hbm_peak)
                if not args.no_cpu_baseline:
                    line["mds"]["cpu_baseline"] = guarded(mds_cpu_baseline, tri_sample, 10000, 10)
            except Exception as e:   # noqa: BLE001
                line["mds"] = {"error": "%s: %s" % (type(e).__name__, e)}
            also["mds"] = compact(line["mds"])
            torch.cuda.empty_cache()
            try:                                   # classify pre-processing
                line["classify_preprocess"], host = preprocess_bench.run(torch, dev, hbm_peak)
                if not args.no_cpu_baseline:
                    line["classify_preprocess"]["cpu_baseline"] = guarded(preprocess_cpu_baseline, host)
            except Exception as e:   # noqa: BLE001
                line["classify_preprocess"] = {"error": "%s: %s" % (type(e).__name__, e)}
            also["classify_preprocess"] = compact(line["classify_preprocess"])
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        dist_pg.destroy_process_group()


if __name__ == "__main__":
    main()
