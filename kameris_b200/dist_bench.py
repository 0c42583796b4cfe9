"""Secondary benchmark lines for bench.py: pairs/s of the distance kernels with their rooflines.

  gram      : N=8192 count vectors of D=4^9 (9 kb genomes, k=9): cosine+pearson+euclid from one
              tcgen05 Gram pass.  Algorithmic flops = 2*D per unique pair.  Counts
              <= 255 run as uint8 x uint8 -> int32 (peak = 2 x the measured bf16 throughput of
              MEASURED_PEAKS.json); the fp16 path on the same data is reported against the bf16 peak.
  manhattan : BASELINE config 4 shape -- M reads x C=1000 centroids, D=4^6, uint32-exact manhat
              (byte path, VABSDIFF4) ; bound = CUDA-core issue rate, reported as pair-elements/s.
"""
import time

import numpy as np


def _time(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def synth_counts(torch, dev, n, L, k, seed, kind="counts"):
    """count (or frequency) vectors of n synthetic genomes, produced by the repr path itself."""
    from . import repr as R
    rng = np.random.Generator(np.random.PCG64(seed))
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    chars = alpha[rng.choice(4, size=n * L, p=[0.36, 0.18, 0.24, 0.22])]
    start = np.arange(n + 1, dtype=np.int64) * L
    d_chars = torch.zeros(len(chars) + 16, dtype=torch.uint8, device=dev)
    d_chars[:len(chars)] = torch.from_numpy(chars.copy()).to(dev)
    batch = R.pack_chars(d_chars, torch.from_numpy(start).to(dev), n, len(chars))
    return R.kmer_vectors(batch, k, 16, kind)


def run(torch, dev, hbm_peak, tf_peak, n_gram=8192, reps=2):
    from . import _lib
    lib = _lib.load()
    out = {}

    # ---- Gram metrics at k=9 -------------------------------------------------------------------------
    n, k = n_gram, 9
    d = 4 ** k
    x = synth_counts(torch, dev, n, 9000, k, 20261003)
    pairs = n * (n - 1) // 2
    tri = {m: torch.empty(pairs, dtype=torch.float32, device=dev) for m in ("euclid", "cosine", "pearson")}
    ws = torch.empty(lib.kam_gram_workspace_bytes(n, d), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def gram():
        _lib.check(lib.kam_gram_tri(x.data_ptr(), 1, n, d, x.stride(0), 0, n, 4 | 8 | 16, 1, tri["euclid"].data_ptr(),
                                    tri["cosine"].data_ptr(), tri["pearson"].data_ptr(), ws.data_ptr(), ws.numel(),
                                    stream))
    t = _time(torch, gram, reps)                  # counts <= 255: uint8 x uint8 -> int32 (tcgen05 kind::i8)
    lib.kam_gram_force_f16(1)
    try:
        t16 = _time(torch, gram, reps)            # same data through the fp16 -> fp32 path (kind::f16)
    finally:
        lib.kam_gram_force_f16(0)
    tflops, tflops16 = 2.0 * d * pairs / t / 1e12, 2.0 * d * pairs / t16 / 1e12
    i8_peak = 2.0 * tf_peak                       # int8 dense tensor rate is 2x the bf16 rate on B200
    out["gram"] = {"metric": "distance_pairs_per_sec", "value": pairs / t, "unit": "pairs/s",
                   "config": {"workload": "N=%d count vectors, D=4^9, euclid+cosine+pearson from one Gram pass" % n,
                              "includes": "operand/statistics pre-pass + tile list + tcgen05 kernel + epilogue",
                              "path": "uint8 operands, int32 accumulation in TMEM (exact); CTA pairs, tcgen05.mma.cta_group::2 (M256 x N256)"},
                   "ms": t * 1e3,
                   "roofline": {"bound": "tensor", "achieved": tflops, "peak": i8_peak, "unit": "TFLOP/s",
                                "frac": tflops / i8_peak, "traffic": None,
                                "peak_note": "2 x measured bf16 peak (%.1f): int8 tensor rate; achieved/bf16 peak = %.2f"
                                             % (tf_peak, tflops / tf_peak),
                                "algorithmic_flops_per_pair": 2 * d, "kernel": "gram_pair_kernel<i8>"},
                   "fp16_path": {"value": pairs / t16, "unit": "pairs/s", "ms": t16 * 1e3,
                                 "roofline": {"bound": "tensor", "achieved": tflops16, "peak": tf_peak,
                                              "unit": "TFLOP/s", "frac": tflops16 / tf_peak,
                                              "kernel": "gram_pair_kernel<f16>"}}}
    del x

    # ---- the same metrics on a `frequencies` matrix (float64 cgr / cgr.sum(), what backend.py:45-56 writes) ----
    # integer recovery fused into the pre-pass (freq_gram.cu), then the same tensor-core Gram
    xf = synth_counts(torch, dev, n, 9000, k, 20261003, kind="freq64")

    def gram_f():
        _lib.check(lib.kam_gram_tri(xf.data_ptr(), 5, n, d, xf.stride(0), 0, n, 4 | 8 | 16, 1, tri["euclid"].data_ptr(),
                                    tri["cosine"].data_ptr(), tri["pearson"].data_ptr(), ws.data_ptr(), ws.numel(),
                                    stream))
    tf = _time(torch, gram_f, reps)
    out["gram"]["frequencies_input"] = {
        "value": pairs / tf, "unit": "pairs/s", "ms": tf * 1e3,
        "what": "float64 frequency rows: counts recovered and verified bit for bit in the pre-pass "
                "(1/16 + 1 sweeps of %.1f GB), then the uint8 tensor-core Gram" % (n * d * 8 / 1e9),
        "extra_ms_vs_counts_input": (tf - t) * 1e3,
        "prepass_read_GBps_lower_bound": (1.0 + 1.0 / 16) * n * d * 8 / max(tf - t, 1e-9) / 1e9}
    del xf
    # rows that are not frequency vectors: float64 Gram on the CUDA cores (DFMA-bound)
    nf, df = 4096, 4096
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    xr = torch.rand((nf, df), dtype=torch.float64, device=dev, generator=g)
    pf = nf * (nf - 1) // 2
    wsf = torch.empty(lib.kam_gram_workspace_bytes(nf, df), dtype=torch.uint8, device=dev)

    def gram_r():
        _lib.check(lib.kam_gram_tri(xr.data_ptr(), 5, nf, df, xr.stride(0), 0, nf, 4 | 8 | 16, 1,
                                    tri["euclid"].data_ptr(), tri["cosine"].data_ptr(), tri["pearson"].data_ptr(),
                                    wsf.data_ptr(), wsf.numel(), stream))
    tr = _time(torch, gram_r, reps)
    dfma_peak = 148 * 58.8 * 1.965e9              # measured FP64 rate (scratch/dfma_bench.cu): 58.8 FMA/clk/SM
    out["gram"]["float64_fallback"] = {
        "value": pf / tr, "unit": "pairs/s", "ms": tr * 1e3,
        "config": {"workload": "N=%d arbitrary float64 rows, D=%d" % (nf, df),
                   "includes": "failed recovery pre-pass + row statistics + gram_f64_kernel"},
        "roofline": {"bound": "fp64 issue", "achieved": pf * df / tr / 1e12, "peak": dfma_peak / 1e12,
                     "unit": "T DFMA/s", "frac": pf * df / tr / dfma_peak, "kernel": "gram_f64_kernel<double>"}}
    del xr, wsf, tri, ws

    # ---- Gram at D = 4096 (configs[0] / [3] shape): 32 K-steps per tile, bound by the epilogue and its stores ----
    ng, kg = 32768, 6
    dg = 4 ** kg
    xg = synth_counts(torch, dev, ng, 9000, kg, 20261008)
    pg = ng * (ng - 1) // 2
    trig = {m: torch.empty(pg, dtype=torch.float32, device=dev) for m in ("euclid", "cosine", "pearson")}
    wsg = torch.empty(lib.kam_gram_workspace_bytes(ng, dg), dtype=torch.uint8, device=dev)

    def gram_small(mask):
        _lib.check(lib.kam_gram_tri(xg.data_ptr(), 1, ng, dg, xg.stride(0), 0, ng, mask, 1, trig["euclid"].data_ptr(),
                                    trig["cosine"].data_ptr(), trig["pearson"].data_ptr(), wsg.data_ptr(), wsg.numel(),
                                    stream))
    t3 = _time(torch, lambda: gram_small(4 | 8 | 16), reps)
    t1 = _time(torch, lambda: gram_small(8), reps)
    out["gram_d4096"] = {
        "metric": "distance_pairs_per_sec", "value": pg / t3, "unit": "pairs/s", "ms": t3 * 1e3,
        "config": {"workload": "N=%d count vectors (9 kb genomes, k=6, D=4096), euclid+cosine+pearson, packed triangles "
                               "on the device" % ng},
        "roofline": {"bound": "hbm (writes)", "achieved": 12.0 * pg / t3 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": 12.0 * pg / t3 / 1e9 / hbm_peak, "algorithmic_bytes_per_pair": 12,
                     "kernel": "gram_pair_kernel<i8> (K = 32 steps: float64 epilogue + triangle stores dominate)",
                     "tensor_TFLOPs": 2.0 * dg * pg / t3 / 1e12},
        "cosine_only": {"value": pg / t1, "unit": "pairs/s", "ms": t1 * 1e3, "GBps": 4.0 * pg / t1 / 1e9}}
    del xg, trig, wsg
    torch.cuda.empty_cache()

    # ---- limb-split path (counts above 255 / sums of squares beyond the fp16 envelope: configs[4]-like vectors) ----
    nl, dl = 4096, 4 ** 9
    gl = torch.Generator(device=dev)
    gl.manual_seed(3)
    xl = torch.poisson(torch.full((nl, dl), 19.0, device=dev), generator=gl)
    xl[:, ::977] *= 40
    xl = xl.to(torch.int32).to(torch.uint16)
    pl = nl * (nl - 1) // 2
    tril = {m: torch.empty(pl, dtype=torch.float32, device=dev) for m in ("cosine", "pearson")}
    wsl = torch.empty(lib.kam_gram_workspace_bytes_int(nl, dl), dtype=torch.uint8, device=dev)

    def gram_limb():
        _lib.check(lib.kam_gram_tri(xl.data_ptr(), 1, nl, dl, xl.stride(0), 0, nl, 8 | 16, 1, None,
                                    tril["cosine"].data_ptr(), tril["pearson"].data_ptr(), wsl.data_ptr(), wsl.numel(),
                                    stream))
    tl = _time(torch, gram_limb, reps)
    lib.kam_gram_set_limb(0)
    try:
        tl0 = _time(torch, gram_limb, 1)
    finally:
        lib.kam_gram_set_limb(1)
    out["gram_limb"] = {
        "metric": "distance_pairs_per_sec", "value": pl / tl, "unit": "pairs/s", "ms": tl * 1e3,
        "config": {"workload": "N=%d count vectors of a 5 Mb genome at k=9 (Poisson(19) + a tail to ~760, D=4^9): "
                               "sum of squares ~1e8 per row, outside the uint8 and fp16 envelopes" % nl,
                   "path": "c = 256 hi + lo: three kind::i8 Grams (4x the MMA work of the uint8 path), float64 combine"},
        "roofline": {"bound": "tensor", "achieved": 2.0 * dl * pl / tl / 1e12, "peak": 2.0 * tf_peak, "unit": "TFLOP/s",
                     "frac": 2.0 * dl * pl / tl / 1e12 / (2.0 * tf_peak), "kernel": "gram_pair_kernel<i8> x3 + combine",
                     "issued_TFLOPs": 4 * 2.0 * dl * pl / tl / 1e12},
        "cuda_core_integer_path": {"value": pl / tl0, "unit": "pairs/s", "ms": tl0 * 1e3,
                                   "what": "the path these vectors took in round 1 (IMAD.WIDE Gram, kam_gram_set_limb(0))"}}
    del xl, tril, wsl
    torch.cuda.empty_cache()

    # ---- manhattan / info / float L1, config-4 shape ---------------------------------------------------------
    # peaks: the issue rate of each kernel's inner operation, measured with scratch/pairop_bench.cu on this GPU
    # type (profiles/r2_pairop_microbench.log), in pair-elements per clock and SM
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    op_peak = {"u8": 199.5, "u32": 50.4, "info": 243.9, "f32": 50.6}
    m, c, k6 = 131072, 1000, 6
    d6 = 4 ** k6
    reads = synth_counts(torch, dev, m, 150, k6, 20261004)
    cent = synth_counts(torch, dev, c, 150, k6, 3)
    o = torch.empty((m, c), dtype=torch.int32, device=dev)
    ws = torch.empty(lib.kam_manhattan_workspace_bytes(m, d6) + lib.kam_manhattan_workspace_bytes(c, d6),
                     dtype=torch.uint8, device=dev)

    def manh(x, y):
        _lib.check(lib.kam_manhattan_rect(x.data_ptr(), y.data_ptr(), 1, m, c, d6, x.stride(0), y.stride(0),
                                          o.data_ptr(), ws.data_ptr(), ws.numel(), stream))

    def pair_line(t, pairs, dd, key, kernel, workload, includes="operand transposition pre-pass + tile kernel"):
        peak = op_peak[key] * sms * 1.965e9
        return {"metric": "distance_pairs_per_sec", "value": pairs / t, "unit": "pairs/s", "ms": t * 1e3,
                "config": {"workload": workload, "includes": includes},
                "roofline": {"bound": "cuda-core issue (%s)" % kernel.split(":")[1].strip(),
                             "achieved": pairs * dd / t / 1e12, "peak": peak / 1e12, "unit": "T pair-elements/s",
                             "frac": pairs * dd / t / peak, "kernel": kernel.split(":")[0],
                             "peak_source": "measured issue rate of the inner op (scratch/pairop_bench.cu): %.1f "
                                            "pair-elements per clk and SM" % op_peak[key]}}
    def tensor_line(t, pairs, dd, k_total, workload, kernel, t_core, includes):
        """manhat / info as a Gram of thermometer rows: the tensor pipe does K = sum of column maxima MACs per pair"""
        kpad = (k_total + 127) // 128 * 128
        fl = 2.0 * kpad * pairs
        return {"metric": "distance_pairs_per_sec", "value": pairs / t, "unit": "pairs/s", "ms": t * 1e3,
                "config": {"workload": workload, "includes": includes, "operand_bytes_per_row": kpad},
                "roofline": {"bound": "tensor", "achieved": fl / t / 1e12, "peak": 2.0 * tf_peak, "unit": "TFLOP/s",
                             "frac": fl / t / 1e12 / (2.0 * tf_peak), "kernel": kernel, "K": kpad, "K_over_d": kpad / dd,
                             "pair_elements_per_s": pairs * dd / t,
                             "speedup_vs_cuda_core_kernel": (t_core / t) if t_core else None}}

    def col_total(*mats):            # sum over columns of the largest count any row holds (counts far below 2^15)
        cm = None
        for x in mats:
            a = x.view(torch.int16).amax(dim=0)
            cm = a if cm is None else torch.maximum(cm, a)
        return int(cm.to(torch.int64).sum()), int(cm.max())

    ktot, top = col_total(reads, cent)
    avg = -(-ktot // d6)
    lib.kam_dist_set_tensor(0)
    t = _time(torch, lambda: manh(reads, cent), reps)
    out["manhattan"] = pair_line(t, m * c, d6, "u8", "pair_tile_kernel<SAD8X4>: VABSDIFF4.U8.ACC",
                                 "%d reads (150 bp, k=6 counts) x %d centroids, D=4096, uint32 manhat, "
                                 "CUDA-core path (kam_dist_set_tensor(0))" % (m, c))
    lib.kam_dist_set_tensor(1)
    if top <= 255 and avg <= 12:
        ws_t = torch.empty(lib.kam_manhattan_workspace_bytes_levels(m, d6, avg) +
                           lib.kam_manhattan_workspace_bytes_levels(c, d6, avg), dtype=torch.uint8, device=dev)

        def manh_t():
            _lib.check(lib.kam_manhattan_rect(reads.data_ptr(), cent.data_ptr(), 1, m, c, d6, reads.stride(0),
                                              cent.stride(0), o.data_ptr(), ws_t.data_ptr(), ws_t.numel(), stream))
        o_core = o.clone()
        tt = _time(torch, manh_t, reps)
        assert torch.equal(o, o_core), "tensor-path manhat differs from the CUDA-core kernel"
        out["manhattan_tensor"] = tensor_line(
            tt, m * c, d6, ktot, "the same %d x %d uint32 manhat on the tensor cores (default path; largest count %d, "
            "column maxima sum to %.2f d)" % (m, c, top, ktot / d6),
            "thermo_colmax / scan / encode kernels + gram_pair_kernel<i8> (integer epilogue)", t,
            "column-maxima pre-pass, thermometer operand rows of both sides, Gram launch")
        del ws_t, o_core
    big = cent.clone()
    big.view(torch.int16)[0, 0] = 300                 # one count above 255 -> the exact 32-bit path
    t = _time(torch, lambda: manh(reads, big), reps)
    out["manhattan_u32"] = pair_line(t, m * c, d6, "u32", "pair_tile_kernel<SAD32>: VABSDIFF.U32",
                                     "same shape with a count above 255: exact mod 2^32 path")
    centf = cent.to(torch.float32) / 145.0
    of = torch.empty((m, c), dtype=torch.float32, device=dev)

    def l1f():
        _lib.check(lib.kam_manhattan_rect_f32(reads.data_ptr(), 1, centf.data_ptr(), m, c, d6, reads.stride(0),
                                              centf.stride(0), of.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    t = _time(torch, l1f, reps)
    out["manhattan_f32_centroids"] = pair_line(t, m * c, d6, "f32", "pair_tile_kernel<L1F32>: 2 x FADD",
                                               "configs[3]'s real workload: uint16 count rows x float32 centroids, float L1")
    del reads, cent, o, of, big, centf
    ni = 32768
    xi = synth_counts(torch, dev, ni, 9000, k6, 20261009)
    pi_ = ni * (ni - 1) // 2
    ti_ = torch.empty(pi_, dtype=torch.float32, device=dev)
    wsi = torch.empty(lib.kam_info_workspace_bytes(ni, d6), dtype=torch.uint8, device=dev)

    def info():
        _lib.check(lib.kam_info_tri(xi.data_ptr(), 1, ni, d6, xi.stride(0), 0, ni, ti_.data_ptr(), wsi.data_ptr(),
                                    wsi.numel(), stream))
    lib.kam_dist_set_tensor(0)
    t = _time(torch, info, reps)
    out["info"] = pair_line(t, pi_, d6, "info", "pair_tile_kernel<JACC>: LOP3 + POPC",
                            "N=%d count vectors (9 kb genomes, k=6, D=4096), reference `info` triangle, CUDA-core path "
                            "(kam_dist_set_tensor(0))" % ni, "presence-bitmap pre-pass + tile kernel")
    lib.kam_dist_set_tensor(1)
    t_core = ti_.clone()
    tt = _time(torch, info, reps)
    assert torch.equal(ti_, t_core), "tensor-path info differs from the CUDA-core kernel"
    out["info_tensor"] = tensor_line(tt, pi_, d6, d6, "the same `info` triangle on the tensor cores (default path): Gram of "
                                     "the presence bytes", "thermo_presence_kernel + gram_pair_kernel<i8> (float32 epilogue)",
                                     t, "presence-byte operand rows + Gram launch")
    del xi, ti_, wsi, ws, t_core
    torch.cuda.empty_cache()
    # the reference's own workload: `manhat` between 9 kb genomes at k = 9 (D = 4^9; counts rarely pass 3)
    n9, d9 = 8192, 4 ** 9
    x9 = synth_counts(torch, dev, n9, 9000, 9, 20261012)
    ktot9, top9 = col_total(x9)
    avg9 = -(-ktot9 // d9)
    p9 = n9 * (n9 - 1) // 2
    tri9 = torch.empty(p9, dtype=torch.int32, device=dev)
    ws9 = torch.empty(lib.kam_manhattan_workspace_bytes_levels(n9, d9, avg9), dtype=torch.uint8, device=dev)

    def man9():
        _lib.check(lib.kam_manhattan_tri(x9.data_ptr(), 1, n9, d9, x9.stride(0), 0, n9, tri9.data_ptr(), ws9.data_ptr(),
                                         ws9.numel(), stream))
    lib.kam_dist_set_tensor(0)
    t = _time(torch, man9, 2)
    out["manhattan_k9"] = pair_line(t, p9, d9, "u8", "pair_tile_kernel<SAD8X4>: VABSDIFF4.U8.ACC",
                                    "N=%d genomes (9 kb, k=9, D=%d) uint32 manhat triangle, CUDA-core path" % (n9, d9))
    lib.kam_dist_set_tensor(1)
    if top9 <= 255 and avg9 <= 12:
        t_core = tri9.clone()
        tt = _time(torch, man9, reps)
        assert torch.equal(tri9, t_core), "tensor-path manhat (k=9) differs from the CUDA-core kernel"
        out["manhattan_k9_tensor"] = tensor_line(
            tt, p9, d9, ktot9, "the same triangle on the tensor cores (default path; largest count %d, column maxima "
            "sum to %.2f d)" % (top9, ktot9 / d9),
            "thermo_colmax / scan / encode kernels + gram_pair_kernel<i8> (integer epilogue)", t,
            "column-maxima pre-pass, thermometer operand rows, Gram launch")
        del t_core
    del x9, tri9, ws9
    torch.cuda.empty_cache()

    # ---- e2e: the whole `distances` step on HOST buffers through kam_dists_host ---------------------------
    # (what run_backend_dists calls between reading the .mm-repr and writing the .mm-dist files): H2D of the
    # count vectors, kernels, D2H of the packed triangles, device buffers allocated and freed inside the call
    ne, ke = 16384, 6
    de = 4 ** ke
    xe = synth_counts(torch, dev, ne, 9000, ke, 20261006).cpu().pin_memory()
    pe = ne * (ne - 1) // 2
    h_man = torch.empty(pe, dtype=torch.int32).pin_memory()
    h_f = [torch.empty(pe, dtype=torch.float32).pin_memory() for _ in range(3)]

    def host_step(mask):
        _lib.check(lib.kam_dists_host(xe.data_ptr(), 1, ne, de, mask,
                                      h_man.data_ptr() if mask & 1 else None, None,
                                      h_f[0].data_ptr() if mask & 4 else None,
                                      h_f[1].data_ptr() if mask & 8 else None,
                                      h_f[2].data_ptr() if mask & 16 else None))

    e2e = {}
    for name, mask, n_out in (("manhat", 1, 1), ("euclid+cosine+pearson", 4 | 8 | 16, 3)):
        host_step(mask)
        host_step(mask)                 # the first calls of a process pay for mapping fresh device memory
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            host_step(mask)
            ts.append(time.perf_counter() - t0)
        dt = sorted(ts)[2]              # median of 5 (cudaMalloc / cudaFree inside the call make single calls noisy)
        e2e[name] = {"value": pe / dt, "unit": "pairs/s", "ms": dt * 1e3, "ms_min_max": [min(ts) * 1e3, max(ts) * 1e3],
                     "h2d_bytes_per_step": ne * de * 2, "d2h_bytes_per_step": pe * 4 * n_out}
    def flat(line):
        return dict(line["roofline"], value=line["value"], ms=line["ms"])
    out["rooflines"] = {"gram_i8": flat(out["gram"]), "gram_f16": flat(out["gram"]["fp16_path"]),
                        "gram_d4096": flat(out["gram_d4096"]), "gram_limb": flat(out["gram_limb"]),
                        "manhattan_u8": flat(out["manhattan"]), "manhattan_u32": flat(out["manhattan_u32"]),
                        "manhattan_f32_centroids": flat(out["manhattan_f32_centroids"]), "info": flat(out["info"])}
    for key in ("manhattan_tensor", "info_tensor", "manhattan_k9", "manhattan_k9_tensor"):
        if key in out:
            out["rooflines"][key] = flat(out[key])
    # ---- e2e of a kmers -> distances job (run_job.py:175-180): host ASCII records -> host count vectors AND
    # host distance triangles, the count matrix handed over on the device (kam_kmers_host_keep + kam_dists_device)
    npl, kpl, Lp = 10000, 6, 9000
    dpl = 4 ** kpl
    rng = np.random.Generator(np.random.PCG64(20261010))
    pch = torch.from_numpy(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=npl * Lp, p=[0.36, 0.18, 0.24, 0.22])]).pin_memory()
    pst = (np.arange(npl + 1, dtype=np.uint64) * np.uint64(Lp))
    h_vec = torch.empty((npl, dpl), dtype=torch.uint16).pin_memory()
    ppl = npl * (npl - 1) // 2
    h_tri = [torch.empty(ppl, dtype=torch.int32).pin_memory(), torch.empty(ppl, dtype=torch.float32).pin_memory()]
    d_keep = torch.empty((npl, dpl), dtype=torch.uint16, device=dev)

    def pipeline(handover):
        if handover:
            _lib.check(lib.kam_kmers_host_keep(pch.data_ptr(), pst.ctypes.data, npl, kpl, 16, 0, h_vec.data_ptr(),
                                               d_keep.data_ptr(), dpl))
            _lib.check(lib.kam_dists_device(d_keep.data_ptr(), 1, npl, dpl, dpl, 1 | 8, 0, h_tri[0].data_ptr(), None, None,
                                            h_tri[1].data_ptr(), None))
        else:
            _lib.check(lib.kam_kmers_host(pch.data_ptr(), pst.ctypes.data, npl, kpl, 16, 0, h_vec.data_ptr()))
            _lib.check(lib.kam_dists_host(h_vec.data_ptr(), 1, npl, dpl, 1 | 8, h_tri[0].data_ptr(), None, None,
                                          h_tri[1].data_ptr(), None))
    pl_res = {}
    for name, ho in (("device_handover", True), ("reupload", False)):
        pipeline(ho)
        pipeline(ho)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            pipeline(ho)
            ts.append(time.perf_counter() - t0)
        dt = sorted(ts)[2]
        pl_res[name] = {"ms": dt * 1e3, "vectors_per_s": npl / dt, "pairs_per_s": ppl / dt}
    out["pipeline"] = {"api": "kam_kmers_host_keep + kam_dists_device (device hand-over) vs kam_kmers_host + kam_dists_host",
                       "config": {"workload": "%d sequences x %d bp, k=%d uint16 counts to the host + manhat and cosine "
                                              "triangles (%d pairs each) to the host" % (npl, Lp, kpl, ppl)},
                       "value": npl / (pl_res["device_handover"]["ms"] * 1e-3), "unit": "vectors/s",
                       "ms": pl_res["device_handover"]["ms"], **pl_res,
                       "h2d_bytes_per_step": npl * Lp, "d2h_bytes_per_step": npl * dpl * 2 + 2 * ppl * 4}
    del pch, h_vec, h_tri, d_keep
    out["e2e"] = {"api": "kam_dists_host (C ABI, pinned host buffers)",
                  "config": {"workload": "%d count vectors (9 kb genomes, k=%d, D=%d, uint16): %d pairs" % (ne, ke, de, pe)},
                  **e2e}
    return out


if __name__ == "__main__":
    import json
    import os
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    pth = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks.update(json.load(open(pth)))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    print(json.dumps(run(torch, dev, peaks["hbm_gbs"], peaks["bf16_tflops"])))
