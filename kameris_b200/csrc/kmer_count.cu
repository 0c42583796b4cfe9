// kmer_count.cu -- planes -> 4^k CGR count / frequency vectors.
//
// Replaces the per-file CGR fill of the reference (generation_cgr_mac_avx2 @0x10000d330 u16,
// @0x100011320 u32; loop @0x10000e1d2-0x10000e304, increment @0x10000ec41) and, for the FREQ
// outputs, the numpy normalisation kameris/job_steps/backend.py:42-56.
//
// Index of the k-mer ending at base j of a sequence: with the x/y bit planes
// LSB-first, x = the k x-bits of bases j-k+1..j read as an integer (oldest base in bit 0), y
// likewise, bin = (y << k) | x.  For the first k-1 bases the CGR "centre" bit is still inside
// the window: x = (bits << (k-m)) | 1 << (k-m-1) with m = j+1 bases seen; the reference emits
// those only when non-base characters pushed the raw index to >= k-1 (quirk Q1), which
// head_mask encodes: base j emits iff j >= popc(head_mask & ((1<<(k-1))-1)).
//
// Path A (cgr_tile_kernel): one CTA per sequence, persistent over a dynamic queue.  The 4^k
// table is produced in tiles of <= 32768 bins held in shared memory (uint8 counters for k >= 7,
// uint32 for k <= 6); per tile the CTA re-scans the (L1-resident) planes of its sequence,
// increments the bins that fall in the tile, then streams the tile out with 16-byte stores in
// the output type, fusing the frequency division.  The output (4^k * sizeof(out) per sequence,
// >96% zeros for 9 kb genomes at k=9) is written exactly once and never read: HBM-write-bound.
// A uint8 counter that would pass 255, or a sequence too long for this path, puts the sequence
// on the overflow list.
//
// Path A for short reads at k <= 6 (cgr_warp_kernel): one WARP per read, private uint8 table.
//
// Path L (cgr_long_kernel): the overflow list of path A -- megabase sequences, uint8 overflows --
// counted in shared-memory tiles of 65536 packed uint16 counters, one CTA per (sequence, tile);
// shared-memory atomics are ~18x faster than global RED.ADD on B200, so scanning a sequence once
// per tile is the cheaper trade.
//
// Path B (cgr_spill_*): whatever overflows even uint16 counters (and k >= 12): uint32 tables in
// global memory (L2-resident: a few sequences at a time), RED.ADD increments from many CTAs per
// sequence, then a finalize pass (wrap to 16 bits, normalise).
//
// Sweep (cgr_sweep_kernel): every order k_lo..k_hi of a batch; orders <= 7 from one scan at k=7 and
// 2x2 pooling of the shared-memory tile.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int TILE_BINS_LOG2 = 15;
constexpr uint32_t TILE_BINS = 1u << TILE_BINS_LOG2;   // bins per shared-memory tile (uint8 counters)
constexpr uint64_t A_MAX_LEN = 1ull << 17;              // longer sequences go to path B when tiled
constexpr int B_THREADS = 256;
constexpr uint32_t B_CHUNK = 32768;                     // bases per CTA in path B

struct SeqInfo {
    uint64_t p0, n;     // first flat base, number of bases
    uint32_t lead;      // bases [lead, n) emit
};

__device__ __forceinline__ SeqInfo seq_info(const uint64_t *__restrict__ base_start,
                                            const uint32_t *__restrict__ head_mask, uint32_t s, int k) {
    SeqInfo q;
    q.p0 = base_start[s];
    q.n = base_start[s + 1] - q.p0;
    q.lead = __popc(head_mask[s] & ((k > 1 ? (1u << (k - 1)) : 1u) - 1u));
    return q;
}

// x and y windows of the k-mer ending at flat position p (needs p - (k-1) >= seq start)
__device__ __forceinline__ uint32_t kmer_bin(unsigned long long vx, unsigned long long vy, uint32_t sh, int k,
                                             uint32_t mask) {
    const uint32_t xw = (uint32_t)(vx >> sh) & mask;
    const uint32_t yw = (uint32_t)(vy >> sh) & mask;
    return (yw << k) | xw;
}

// The <= k-1 partial k-mers at the head of a dirty sequence (quirk Q1).  Serial, executed by one
// thread; `emit(bin)` is called for each.
template <class F>
__device__ __forceinline__ void head_partials(const uint2 *__restrict__ planes, const SeqInfo &q, int k, F emit) {
    const uint64_t jend = q.n < (uint64_t)(k - 1) ? q.n : (uint64_t)(k - 1);
    uint32_t x = 1u << (k - 1), y = x;   // the literal reference recurrence
    for (uint64_t j = 0; j < jend; ++j) {
        const uint64_t p = q.p0 + j;
        const uint2 w = planes[p >> 5];
        const uint32_t b = (uint32_t)(p & 31);
        x = (x >> 1) | (((w.x >> b) & 1u) << (k - 1));
        y = (y >> 1) | (((w.y >> b) & 1u) << (k - 1));
        if (j >= q.lead) emit((y << k) | x);
    }
}

template <typename OUT>
__device__ __forceinline__ OUT convert_count(uint32_t c, double inv_unused, double sum) {
    (void)inv_unused;
    if constexpr (sizeof(OUT) == 8 && !std::is_integral<OUT>::value) return (double)c / sum;   // numpy true division
    else if constexpr (std::is_same<OUT, float>::value) return (float)((double)c / sum);
    else return (OUT)c;   // wraps mod 2^bits like `inc WORD/DWORD PTR`
}

// ---- path A ------------------------------------------------------------------------------------
// CT: shared-memory counter type (uint8_t with overflow detection, or uint32_t).
// OUT: uint16_t / uint32_t (counts), float / double (frequencies).

// one increment of local bin lb (uint8 counters packed four to a word, or uint32 counters)
template <typename CT>
__device__ __forceinline__ void bump(uint32_t *tab32, uint32_t lb, uint32_t *s_flag) {
    if constexpr (sizeof(CT) == 1) {
        const uint32_t bsh = (lb & 3u) * 8u;
        const uint32_t old = atomicAdd(&tab32[lb >> 2], 1u << bsh);
        if (((old >> bsh) & 0xffu) == 0xffu) *s_flag = 1;   // 255 -> 256 carried into the neighbour: redo on path B
    } else {
        atomicAdd(&tab32[lb], 1u);
    }
}


// Scan of one sequence for one shared-memory tile, 32 k-mer end positions (one plane word) per thread
// and trip: the in-tile mask of all 32 positions is formed with whole-word operations (the tile index is
// the top `tbits` bits of the y window = the y-bits of the newest tbits bases), then one increment per set
// bit.  Replaces a scan that handled 8 positions per thread and trip (4x the per-position overhead);
// genome-length sequences only -- a 150 bp read has too few words to occupy a CTA this way.
template <typename CT, int THREADS>
__device__ __forceinline__ void scan_tile_words(const uint2 *__restrict__ planes, uint64_t pf, uint64_t pe, int k,
                                                uint32_t t, uint32_t tbits, uint32_t ylow, uint32_t *tab32,
                                                uint32_t *s_flag) {
    const uint32_t mask = (1u << k) - 1u;
    const uint32_t shc = 32u - (uint32_t)(k - 1);
    const uint64_t W0 = pf >> 5, W1 = ((pe - 1) >> 5) + 1;
    for (uint64_t W = W0 + threadIdx.x; W < W1; W += THREADS) {
        const uint2 hi = __ldg(planes + W);
        const uint2 lo = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
        const unsigned long long vx = ((unsigned long long)hi.x << 32) | lo.x;
        const unsigned long long vy = ((unsigned long long)hi.y << 32) | lo.y;
        const uint64_t pbase = W << 5;
        uint32_t m = 0xffffffffu;
        if (pbase < pf) m &= 0xffffffffu << (uint32_t)(pf - pbase);
        if (pbase + 32 > pe) m &= 0xffffffffu >> (uint32_t)(pbase + 32 - pe);
#pragma unroll 1
        for (uint32_t i = 0; i < tbits; ++i) {
            const uint32_t want = 0u - ((t >> (tbits - 1 - i)) & 1u);
            m &= ~((uint32_t)(vy >> (32u - i)) ^ want);
        }
        while (m) {
            uint32_t b;
            asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(m));
            m ^= 1u << b;
            const uint32_t sh = b + shc;
            const uint32_t xw = (uint32_t)(vx >> sh) & mask, yw = (uint32_t)(vy >> sh) & ylow;
            bump<CT>(tab32, (yw << k) | xw, s_flag);
        }
    }
}

// THREADS: 256 for genome-length sequences; 64 for short reads, where a sequence gives a CTA only a
// few dozen k-mer groups and the per-sequence barriers dominate (more, smaller CTAs per SM hide them).
template <typename CT, typename OUT, int THREADS>
__global__ void __launch_bounds__(THREADS) cgr_tile_kernel(const uint2 *__restrict__ planes,
                                                             const uint64_t *__restrict__ base_start,
                                                             const uint32_t *__restrict__ head_mask, uint32_t n_seqs,
                                                             int k, int count_bits, OUT *__restrict__ out,
                                                             uint64_t out_stride, uint32_t *__restrict__ queue,
                                                             uint32_t *__restrict__ spill_count,
                                                             uint32_t *__restrict__ spill_list,
                                                             unsigned long long *__restrict__ spill_maxlen,
                                                             int word_scan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t s_seq, s_next, s_end;
    __shared__ uint32_t s_flag;
    __shared__ unsigned long long s_red[THREADS / 32];
    constexpr uint32_t QB = THREADS == 64 ? 8u : 1u;
    if (threadIdx.x == 0) s_next = s_end = 0;
    // quotient of every possible uint8 count for the current sequence: the (slow, exact) double
    // division runs 256 times per sequence instead of 4^k times
    __shared__ OUT s_lut[256];
    constexpr bool IS_FREQ = !std::is_integral<OUT>::value;

    const uint32_t bins = 1u << (2 * k);
    const uint32_t tile_bins = bins < TILE_BINS ? bins : TILE_BINS;
    const uint32_t n_tiles = bins / tile_bins;
    const uint32_t tile_words = tile_bins * sizeof(CT) / 4;   // >= 1
    const uint32_t mask = (1u << k) - 1u;                      // k <= 15 here
    // tile index of a bin = its top bits = top bits of y:  bin >> 15 == y >> ysh
    const uint32_t ysh = n_tiles > 1 ? (uint32_t)(TILE_BINS_LOG2 - k) : (uint32_t)k;
    const uint32_t ylow = (1u << ysh) - 1u;
    const uint32_t tbits = n_tiles > 1 ? (uint32_t)(2 * k - TILE_BINS_LOG2) : 0u;   // log2(n_tiles)
    constexpr int BPS = 16 / (int)sizeof(OUT);                // bins per 16-byte store

    for (uint32_t i = threadIdx.x; i < tile_words; i += THREADS) tab32[i] = 0;

    for (;;) {
        __syncthreads();   // table is clear; previous s_seq/s_flag/s_lut consumed
        if (threadIdx.x == 0) {
            if (s_next >= s_end) {            // short reads take QB sequences per global atomic
                s_next = atomicAdd(queue, QB);
                s_end = s_next + QB;
            }
            s_seq = s_next++;
            s_flag = 0;
        }
        __syncthreads();
        const uint32_t s = s_seq;
        if (s >= n_seqs) break;
        const SeqInfo q = seq_info(base_start, head_mask, s, k);
        OUT *row = out + (uint64_t)s * out_stride;

        bool spill = (n_tiles > 1 && q.n > A_MAX_LEN);
        if (sizeof(CT) == 1 && q.n > 64ull * bins && q.n > 255) spill = true;   // mean count > 64: uint8 will overflow
        if (spill) {
            if (threadIdx.x == 0) {
                spill_list[atomicAdd(spill_count, 1u)] = s;
                atomicMax(spill_maxlen, (unsigned long long)q.n);
            }
            continue;
        }

        // groups of 8 flat positions covering the full k-mers, i.e. flat positions [pf, pe)
        const uint64_t pf = q.p0 + (uint64_t)(k - 1);
        const uint64_t pe = q.p0 + q.n;
        const bool has_full = q.n >= (uint64_t)k;
        const uint64_t g0 = pf >> 3, g1 = has_full ? ((pe - 1) >> 3) + 1 : g0;
        // sum of the counts = number of emitting bases (no wrap on this path unless handled below)
        double sum = (double)(q.n > q.lead ? q.n - q.lead : 0);
        if constexpr (IS_FREQ) {   // (recomputed below when the divisor turns out to be a wrapped sum)
            for (uint32_t i = threadIdx.x; i < 256; i += THREADS) s_lut[i] = convert_count<OUT>(i, 0.0, sum);
        }

        for (uint32_t t = 0; t < n_tiles; ++t) {
            if (THREADS == 256 && has_full && word_scan) {
                scan_tile_words<CT, THREADS>(planes, pf, pe, k, t, tbits, ylow, tab32, &s_flag);
            } else
            for (uint64_t g = g0 + threadIdx.x; g < g1; g += THREADS) {
                const uint64_t W = g >> 2;                  // word holding the group
                const uint2 hi = __ldg(planes + W);
                const uint2 lo = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
                // bits [8g-(k-1), 8g+8) of each plane, i.e. the 8 windows of this group
                const uint32_t sh0 = 8u * (uint32_t)(g & 3) + 32u - (uint32_t)(k - 1);
                const uint32_t gx = (uint32_t)((((unsigned long long)hi.x << 32) | lo.x) >> sh0);
                const uint32_t gy = (uint32_t)((((unsigned long long)hi.y << 32) | lo.y) >> sh0);
                // which of the 8 positions are inside [pf, pe)
                const uint64_t pbase = g << 3;
                uint32_t m = 0xffu;
                if (pbase < pf) m &= 0xffu << (uint32_t)(pf - pbase);
                if (pbase + 8 > pe) m &= 0xffu >> (uint32_t)(pbase + 8 - pe);
                // ... and fall into tile t: the tile index is the top `tbits` bits of y, i.e. the y-bits
                // of the newest `tbits` bases of each k-mer -> one whole-word compare per tile bit
#pragma unroll 1
                for (uint32_t i = 0; i < tbits; ++i) {
                    const uint32_t want = 0u - ((t >> (tbits - 1 - i)) & 1u);        // all-ones / all-zeros
                    m &= ~((gy >> (uint32_t)(k - 1 - i)) ^ want);
                }
                while (m) {                                   // on average one k-mer of the 8 is in the tile
                    const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1u;
                    const uint32_t yw = (gy >> b) & mask, xw = (gx >> b) & mask;
                    bump<CT>(tab32, ((yw & ylow) << k) | xw, &s_flag);
                }
            }
            if (threadIdx.x == 0 && q.lead < (uint32_t)(k - 1)) {   // dirty head (rare)
                head_partials(planes, q, k, [&](uint32_t bin) {
                    if (n_tiles > 1 && (bin >> TILE_BINS_LOG2) != t) return;
                    bump<CT>(tab32, bin & (tile_bins - 1), &s_flag);
                });
            }
            __syncthreads();

            if constexpr (sizeof(CT) == 4) {
                // frequencies of 16-bit counts that wrapped: the divisor is the sum of the WRAPPED values
                if (t == 0 && IS_FREQ && count_bits == 16 && q.n > 65535) {
                    unsigned long long part = 0;
                    for (uint32_t i = threadIdx.x; i < tile_bins; i += THREADS) part += tab32[i] & 0xffffu;
                    part = warp_sum64(part);
                    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
                    __syncthreads();
                    unsigned long long tot = 0;
#pragma unroll
                    for (int w = 0; w < THREADS / 32; ++w) tot += s_red[w];
                    sum = (double)tot;
                    __syncthreads();
                    for (uint32_t i = threadIdx.x; i < 256; i += THREADS) s_lut[i] = convert_count<OUT>(i, 0.0, sum);
                    __syncthreads();
                }
            }

            // stream the tile out (and clear it).  Store c covers bins [c*BPS, c*BPS+BPS).
            OUT *trow = row + (uint64_t)t * tile_bins;
            const bool vec_ok = (tile_bins % BPS == 0) && ((reinterpret_cast<uintptr_t>(trow) & 15) == 0);
            const uint32_t wrapmask = (count_bits == 16) ? 0xffffu : 0xffffffffu;
            if (vec_ok) {
                const uint32_t n_stores = tile_bins / BPS;
                for (uint32_t c = threadIdx.x; c < n_stores; c += THREADS) {
                    union {
                        uint4 v;
                        OUT e[BPS];
                    } u;
                    if constexpr (sizeof(CT) == 1) {
                        // BPS uint8 counters = BPS/4 words (BPS==2: half a word); branch-free
                        uint32_t w0 = 0, w1 = 0;
                        if constexpr (BPS == 8) {
                            const uint2 w = *reinterpret_cast<uint2 *>(&tab32[c * 2]);
                            w0 = w.x, w1 = w.y;
                            *reinterpret_cast<uint2 *>(&tab32[c * 2]) = make_uint2(0, 0);
                        } else if constexpr (BPS == 4) {
                            w0 = tab32[c];
                            tab32[c] = 0;
                        } else {
                            unsigned short *t16 = reinterpret_cast<unsigned short *>(tab32);
                            w0 = t16[c];
                            t16[c] = 0;
                        }
                        if constexpr (IS_FREQ) {              // >96% of the lanes read s_lut[0]: a broadcast
#pragma unroll
                            for (int i = 0; i < BPS; ++i) u.e[i] = s_lut[((i < 4 ? w0 : w1) >> (8 * (i & 3))) & 0xffu];
                        } else if constexpr (BPS == 8) {      // uint16 counts: two bytes -> two halfwords per PRMT
                            u.v = make_uint4(__byte_perm(w0, 0, 0x4140), __byte_perm(w0, 0, 0x4342),
                                             __byte_perm(w1, 0, 0x4140), __byte_perm(w1, 0, 0x4342));
                        } else {                              // uint32 counts
                            u.v = make_uint4(w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu, w0 >> 24);
                        }
                    } else {
                        // uint32 counters: vector read-and-clear, then quotient LUT for counts < 256
                        uint32_t cnt[BPS];
                        if constexpr (BPS == 2) {
                            const uint2 w = *reinterpret_cast<uint2 *>(&tab32[c * 2]);
                            *reinterpret_cast<uint2 *>(&tab32[c * 2]) = make_uint2(0, 0);
                            cnt[0] = w.x & wrapmask, cnt[1] = w.y & wrapmask;
                        } else {
#pragma unroll
                            for (int q4 = 0; q4 < BPS / 4; ++q4) {
                                const uint4 w = *reinterpret_cast<uint4 *>(&tab32[c * BPS + 4 * q4]);
                                *reinterpret_cast<uint4 *>(&tab32[c * BPS + 4 * q4]) = make_uint4(0, 0, 0, 0);
                                cnt[4 * q4] = w.x & wrapmask, cnt[4 * q4 + 1] = w.y & wrapmask;
                                cnt[4 * q4 + 2] = w.z & wrapmask, cnt[4 * q4 + 3] = w.w & wrapmask;
                            }
                        }
                        if constexpr (IS_FREQ) {
#pragma unroll
                            for (int i = 0; i < BPS; ++i)
                                u.e[i] = cnt[i] < 256u ? s_lut[cnt[i]] : convert_count<OUT>(cnt[i], 0.0, sum);
                        } else if constexpr (BPS == 8) {      // uint16 counts: low halves of two counters per PRMT
                            u.v = make_uint4(__byte_perm(cnt[0], cnt[1], 0x5410), __byte_perm(cnt[2], cnt[3], 0x5410),
                                             __byte_perm(cnt[4], cnt[5], 0x5410), __byte_perm(cnt[6], cnt[7], 0x5410));
                        } else {
                            u.v = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
                        }
                    }
                    st_stream16(trow + (uint64_t)c * BPS, u.v);
                }
            } else {   // tiny or misaligned rows (k=1 with 16-bit counts)
                for (uint32_t i = threadIdx.x; i < tile_bins; i += THREADS) {
                    uint32_t c;
                    if constexpr (sizeof(CT) == 1) c = reinterpret_cast<unsigned char *>(tab32)[i];
                    else c = tab32[i] & wrapmask;
                    trow[i] = convert_count<OUT>(c, 0.0, sum);
                }
                __syncthreads();
                for (uint32_t i = threadIdx.x; i < tile_words; i += THREADS) tab32[i] = 0;
            }
            __syncthreads();
        }
        if (sizeof(CT) == 1 && threadIdx.x == 0 && s_flag) {   // a uint8 counter overflowed: redo on path B
            spill_list[atomicAdd(spill_count, 1u)] = s;
            atomicMax(spill_maxlen, (unsigned long long)q.n);
        }
    }
}

// ---- path A with a SPARSE read-out ---------------------------------------------------------------------
// A 9 kb genome at k=9 touches <= 8992 of 262144 bins: as a dense float64 row it is 2 MB of which 96.6 % are
// zeros, and the host-buffer form of the step (kam_kmers_host) used to spend 99.6 % of its time moving those
// zeros over PCIe.  Here the same shared-memory tiles are read out as a list of packed entries
// (bin << 8 | count), in ascending bin order, <= one entry per emitting base: the CTA's warps own consecutive
// runs of counter words, a first sweep counts the non-zero counters per warp, a second writes them at the
// warp's offset (ballot to skip all-zero groups, shuffle scan inside a group) and clears the tile.  The host
// expands a list into the caller's dense row with non-temporal stores (host_kmers.cu).  A sequence this path
// declines (too long for the tiles, or a uint8 counter overflowed) gets nnz = KAM_SPARSE_DECLINED.
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v) {   // 0x80 in every non-zero byte
    return (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) cgr_sparse_kernel(const uint2 *__restrict__ planes,
                                                               const uint64_t *__restrict__ base_start,
                                                               const uint32_t *__restrict__ head_mask, uint32_t n_seqs,
                                                               int k, const uint64_t *__restrict__ entry_start,
                                                               uint32_t *__restrict__ entries,
                                                               uint32_t *__restrict__ nnz_out,
                                                               uint32_t *__restrict__ queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t s_seq, s_flag;
    __shared__ uint32_t s_wtot[THREADS / 32];
    constexpr uint32_t WARPS = THREADS / 32;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t bins = 1u << (2 * k);
    const uint32_t tile_bins = bins < TILE_BINS ? bins : TILE_BINS;
    const uint32_t n_tiles = bins / tile_bins;
    const uint32_t tile_words = tile_bins / 4;                 // uint8 counters, four to a word
    const uint32_t wpw = tile_words / WARPS;                   // words per warp (k >= 7: a multiple of 32)
    const uint32_t ysh = n_tiles > 1 ? (uint32_t)(TILE_BINS_LOG2 - k) : (uint32_t)k;
    const uint32_t ylow = (1u << ysh) - 1u;
    const uint32_t tbits = n_tiles > 1 ? (uint32_t)(2 * k - TILE_BINS_LOG2) : 0u;

    for (uint32_t i = threadIdx.x; i < tile_words; i += THREADS) tab32[i] = 0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_seq = atomicAdd(queue, 1u);
            s_flag = 0;
        }
        __syncthreads();
        const uint32_t s = s_seq;
        if (s >= n_seqs) break;
        const SeqInfo q = seq_info(base_start, head_mask, s, k);
        if (q.n > A_MAX_LEN || (q.n > 64ull * bins && q.n > 255)) {
            if (threadIdx.x == 0) nnz_out[s] = KAM_SPARSE_DECLINED;
            continue;
        }
        uint32_t *ent = entries + entry_start[s];
        const uint64_t pf = q.p0 + (uint64_t)(k - 1), pe = q.p0 + q.n;
        const bool has_full = q.n >= (uint64_t)k;
        uint32_t total = 0;                                    // CTA-uniform
        for (uint32_t t = 0; t < n_tiles; ++t) {
            if (has_full) scan_tile_words<uint8_t, THREADS>(planes, pf, pe, k, t, tbits, ylow, tab32, &s_flag);
            if (threadIdx.x == 0 && q.lead < (uint32_t)(k - 1)) {   // dirty head (rare)
                head_partials(planes, q, k, [&](uint32_t bin) {
                    if (n_tiles > 1 && (bin >> TILE_BINS_LOG2) != t) return;
                    bump<uint8_t>(tab32, bin & (tile_bins - 1), &s_flag);
                });
            }
            __syncthreads();
            const uint32_t w0 = warp * wpw;
            uint32_t c = 0;
            for (uint32_t it = 0; it < wpw; it += 32) c += __popc(nonzero_bytes(tab32[w0 + it + lane]));
            c = warp_sum(c);
            if (lane == 0) s_wtot[warp] = c;
            __syncthreads();
            uint32_t base = total, tile_total = 0;
#pragma unroll
            for (uint32_t w = 0; w < WARPS; ++w) {
                const uint32_t x = s_wtot[w];
                if (w < warp) base += x;
                tile_total += x;
            }
            if (c) {
                for (uint32_t it = 0; it < wpw; it += 32) {
                    const uint32_t widx = w0 + it + lane;
                    const uint32_t v = tab32[widx];
                    const uint32_t cnt = __popc(nonzero_bytes(v));
                    if (__ballot_sync(0xffffffffu, cnt != 0) == 0u) continue;   // > 85 % of the groups at k=9
                    uint32_t incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (uint32_t)o) incl += y;
                    }
                    if (cnt) {
                        tab32[widx] = 0;
                        uint32_t pos = base + incl - cnt;
                        const uint32_t bin0 = t * tile_bins + widx * 4u;
#pragma unroll
                        for (uint32_t b = 0; b < 4; ++b) {
                            const uint32_t byte = (v >> (8u * b)) & 0xffu;
                            if (byte) ent[pos++] = ((bin0 + b) << 8) | byte;
                        }
                    }
                    base += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            total += tile_total;
            __syncthreads();   // tile clear again, s_wtot consumed
        }
        if (threadIdx.x == 0) nnz_out[s] = s_flag ? KAM_SPARSE_DECLINED : total;
    }
}

// ---- path A for short reads at k <= 6: one WARP per read ---------------------------------------------
// A 150 bp read has ~18 groups of 8 k-mers and a 4096-bin table: a CTA per read spends its time in
// barriers and dependent global-load latencies, and uint32 counters cost 32 KB of shared-memory
// traffic per read for the read-out alone.  Here every warp owns a private table of 4^k uint8 counters
// (4 KB at k=6; a read of <= 255 bases cannot overflow them), grabs QB reads per global atomic, fetches
// the QB+1 offsets and head masks with ONE coalesced load (lane i <- read s0+i) and only ever needs
// __syncwarp.  The read-out is 16-byte streaming stores in the output type, exactly one write per byte
// of the output.  Longer reads in the batch: counters are checked for the 255 -> 256 carry and the read
// is redone on path B.
constexpr int WK_WARPS = 8;
constexpr uint32_t WK_QB = 8;
int g_word_scan = 1;                    // test hook: 0 = path A scans 8 positions per thread and trip (the first version)
int g_warp_path = 1;                    // test hook: 0 = short reads at k <= 6 use the CTA-per-read kernel

template <bool CHECK>
__device__ __forceinline__ uint32_t warp_scan_read(const uint2 *__restrict__ planes, const SeqInfo &q, int k,
                                                   uint32_t *tab32, uint32_t lane) {
    const uint32_t mask = (1u << k) - 1u;
    const uint64_t pf = q.p0 + (uint64_t)(k - 1), pe = q.p0 + q.n;
    uint32_t over = 0;
    if (q.n >= (uint64_t)k) {
        const uint64_t g0 = pf >> 3, g1 = ((pe - 1) >> 3) + 1;
        for (uint64_t g = g0 + lane; g < g1; g += 32) {
            const uint64_t W = g >> 2;
            const uint2 hi = __ldg(planes + W);
            const uint2 lo = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
            const uint32_t sh0 = 8u * (uint32_t)(g & 3) + 32u - (uint32_t)(k - 1);
            const uint32_t gx = (uint32_t)((((unsigned long long)hi.x << 32) | lo.x) >> sh0);
            const uint32_t gy = (uint32_t)((((unsigned long long)hi.y << 32) | lo.y) >> sh0);
            const uint64_t pbase = g << 3;
            uint32_t m = 0xffu;
            if (pbase < pf) m &= 0xffu << (uint32_t)(pf - pbase);
            if (pbase + 8 > pe) m &= 0xffu >> (uint32_t)(pbase + 8 - pe);
            while (m) {
                const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                m &= m - 1u;
                const uint32_t lb = (((gy >> b) & mask) << k) | ((gx >> b) & mask);
                const uint32_t bsh = (lb & 3u) * 8u;
                if constexpr (CHECK) {
                    const uint32_t old = atomicAdd(&tab32[lb >> 2], 1u << bsh);
                    over |= (((old >> bsh) & 0xffu) == 0xffu);
                } else {
                    atomicAdd(&tab32[lb >> 2], 1u << bsh);   // result unused
                }
            }
        }
    }
    if (lane == 0 && q.lead < (uint32_t)(k - 1)) {   // dirty head (rare)
        head_partials(planes, q, k, [&](uint32_t bin) {
            const uint32_t bsh = (bin & 3u) * 8u;
            const uint32_t old = atomicAdd(&tab32[bin >> 2], 1u << bsh);
            over |= (((old >> bsh) & 0xffu) == 0xffu);
        });
    }
    return over;
}

template <typename OUT>
__global__ void __launch_bounds__(WK_WARPS * 32) cgr_warp_kernel(const uint2 *__restrict__ planes,
                                                                 const uint64_t *__restrict__ base_start,
                                                                 const uint32_t *__restrict__ head_mask,
                                                                 uint32_t n_seqs, int k, OUT *__restrict__ out,
                                                                 uint64_t out_stride, uint32_t *__restrict__ queue,
                                                                 uint32_t *__restrict__ spill_count,
                                                                 uint32_t *__restrict__ spill_list,
                                                                 unsigned long long *__restrict__ spill_maxlen) {
    constexpr bool IS_FREQ = !std::is_integral<OUT>::value;
    constexpr int BPS = 16 / (int)sizeof(OUT);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t bins = 1u << (2 * k);
    const uint32_t tab_bytes = bins < 16u ? 16u : bins;
    const uint32_t per_warp = tab_bytes + (IS_FREQ ? 256u * (uint32_t)sizeof(OUT) : 0u);
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(smem_raw + (size_t)warp * per_warp);
    OUT *lut = reinterpret_cast<OUT *>(smem_raw + (size_t)warp * per_warp + tab_bytes);
    (void)lut;
    for (uint32_t i = lane; i < tab_bytes / 4; i += 32) tab32[i] = 0;
    __syncwarp();

    for (;;) {
        uint32_t s0 = 0;
        if (lane == 0) s0 = atomicAdd(queue, WK_QB);
        s0 = __shfl_sync(0xffffffffu, s0, 0);
        if (s0 >= n_seqs) break;
        const uint32_t nb = n_seqs - s0 < WK_QB ? n_seqs - s0 : WK_QB;
        const unsigned long long bs_l = lane <= nb ? base_start[s0 + lane] : 0ull;
        const uint32_t hm_l = lane < nb ? head_mask[s0 + lane] : 0u;
        for (uint32_t r = 0; r < nb; ++r) {
            const uint32_t s = s0 + r;
            SeqInfo q;
            q.p0 = __shfl_sync(0xffffffffu, bs_l, (int)r);
            q.n = __shfl_sync(0xffffffffu, bs_l, (int)r + 1) - q.p0;
            q.lead = __popc(__shfl_sync(0xffffffffu, hm_l, (int)r) & ((k > 1 ? (1u << (k - 1)) : 1u) - 1u));
            OUT *row = out + (uint64_t)s * out_stride;
            if (q.n > 64ull * bins && q.n > 255) {   // mean count > 64: uint8 would overflow, straight to path B
                if (lane == 0) {
                    spill_list[atomicAdd(spill_count, 1u)] = s;
                    atomicMax(spill_maxlen, (unsigned long long)q.n);
                }
                continue;
            }
            const double sum = (double)(q.n > q.lead ? q.n - q.lead : 0);
            uint32_t maxc = 255u;
            if constexpr (IS_FREQ) {   // quotient of every count this read can produce (<= its number of k-mers)
                if (sum < 255.0) maxc = (uint32_t)sum;
                for (uint32_t i = lane; i <= maxc; i += 32) lut[i] = convert_count<OUT>(i, 0.0, sum);
            }
            uint32_t over = q.n <= 255 ? warp_scan_read<false>(planes, q, k, tab32, lane)
                                       : warp_scan_read<true>(planes, q, k, tab32, lane);
            over = __any_sync(0xffffffffu, over);
            __syncwarp();                            // table / LUT writes visible to every lane before the read-out
            if (over && lane == 0) {
                spill_list[atomicAdd(spill_count, 1u)] = s;
                atomicMax(spill_maxlen, (unsigned long long)q.n);
            }
            // read-out (and clear); a row that path B will redo is still written, then overwritten
            const bool vec_ok = (bins % BPS == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
            if (vec_ok) {
                const uint32_t n_stores = bins / BPS;
                for (uint32_t c = lane; c < n_stores; c += 32) {
                    union {
                        uint4 v;
                        OUT e[BPS];
                    } u;
                    uint32_t w0 = 0, w1 = 0;
                    if constexpr (BPS == 8) {
                        const uint2 w = *reinterpret_cast<uint2 *>(&tab32[c * 2]);
                        w0 = w.x, w1 = w.y;
                        if (w0 | w1) *reinterpret_cast<uint2 *>(&tab32[c * 2]) = make_uint2(0, 0);
                    } else if constexpr (BPS == 4) {
                        w0 = tab32[c];
                        if (w0) tab32[c] = 0;
                    } else {
                        unsigned short *t16 = reinterpret_cast<unsigned short *>(tab32);
                        w0 = t16[c];
                        if (w0) t16[c] = 0;
                    }
                    if constexpr (IS_FREQ) {
#pragma unroll
                        for (int i = 0; i < BPS; ++i) {
                            const uint32_t cnt = ((i < 4 ? w0 : w1) >> (8 * (i & 3))) & 0xffu;
                            u.e[i] = lut[cnt < maxc ? cnt : maxc];   // cnt > maxc only on an overflowed (redone) row
                        }
                    } else if constexpr (BPS == 8) {
                        u.v = make_uint4(__byte_perm(w0, 0, 0x4140), __byte_perm(w0, 0, 0x4342),
                                         __byte_perm(w1, 0, 0x4140), __byte_perm(w1, 0, 0x4342));
                    } else {
                        u.v = make_uint4(w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu, w0 >> 24);
                    }
                    st_stream16(row + (uint64_t)c * BPS, u.v);
                }
            } else {   // tiny or misaligned rows (k = 1, or odd strides)
                unsigned char *t8 = reinterpret_cast<unsigned char *>(tab32);
                for (uint32_t i = lane; i < bins; i += 32) {
                    const uint32_t cnt = t8[i];
                    if constexpr (IS_FREQ) row[i] = lut[cnt < maxc ? cnt : maxc];
                    else row[i] = (OUT)cnt;
                }
                __syncwarp();
                for (uint32_t i = lane; i < tab_bytes / 4; i += 32) tab32[i] = 0;
            }
            __syncwarp();
        }
    }
}

// ---- fused sweep: every order k_lo..k_hi from ONE scan at k_hi ---------------------
// The 2^k x 2^k CGR image of order k is the 2x2 sum-pool of the order-(k+1) image plus the one k-mer
// that ends at base k-1 (it has no (k+1)-mer ending there).  A shared-memory tile of the k_hi table
// holds whole rows, so each tile can be pooled down level by level inside the CTA and every level's
// slice streamed out while the tile is resident: the sequence is scanned once per tile of the
// LARGEST table only, and the lower orders cost their output bytes.  Levels that shrink below one
// row per tile (k < 2*k_hi - 15) are accumulated across tiles in a small table and finished at the
// end of the sequence.  Dirty heads (quirk Q1): the partial k-mers of a level are added just before
// that level is written and removed again before it is pooled.  Sequences longer than 65535 bases
// (pooled uint16 counters) or with a uint8 overflow go to the spill list and are redone per order.
constexpr int SW_THREADS = 256;
constexpr int SW_MAX_K = 10;            // fused top order limit: the first order below one row per tile stays <= 256 bins
int g_sweep_top = 7;                    // top order of the fused part (tuning / test hook)
constexpr uint32_t SW_NONE = 0xffffffffu;

struct SweepPtrs {
    void *out[16];                      // indexed by k
};

template <typename OUT>
__device__ __forceinline__ OUT sweep_convert(uint32_t c, const OUT *lut, double sum) {
    if constexpr (std::is_integral<OUT>::value) return (OUT)c;
    else return c < 256u ? lut[c] : convert_count<OUT>(c, 0.0, sum);
}

template <typename OUT>
__global__ void __launch_bounds__(SW_THREADS) cgr_sweep_kernel(const uint2 *__restrict__ planes,
                                                               const uint64_t *__restrict__ base_start,
                                                               const uint32_t *__restrict__ head_mask, uint32_t n_seqs,
                                                               int k_hi, int k_lo, SweepPtrs so,
                                                               uint32_t *__restrict__ queue,
                                                               uint32_t *__restrict__ spill_count,
                                                               uint32_t *__restrict__ spill_list,
                                                               unsigned long long *__restrict__ spill_maxlen) {
    constexpr bool IS_FREQ = !std::is_integral<OUT>::value;
    constexpr int BPS = 16 / (int)sizeof(OUT);
    // tile geometry: the whole table when it has <= 32768 bins (k_hi <= 7), else 32768-bin row blocks
    const uint32_t bins_hi = 1u << (2 * k_hi);
    const uint32_t tile_log2 = 2 * k_hi < TILE_BINS_LOG2 ? (uint32_t)(2 * k_hi) : (uint32_t)TILE_BINS_LOG2;
    const uint32_t tile_bins = 1u << tile_log2;
    const uint32_t n_tiles = bins_hi / tile_bins;
    const uint32_t tbits = (uint32_t)(2 * k_hi) - tile_log2;     // log2(n_tiles)
    const int k1 = (int)tbits;                                    // order at which a tile holds exactly one row
    const uint32_t mask = (1u << k_hi) - 1u;
    const uint32_t ysh = tile_log2 - (uint32_t)k_hi, ylow = (1u << ysh) - 1u;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab32 = reinterpret_cast<uint32_t *>(smem_raw);                                  // tile_bins uint8
    unsigned short *bufA = reinterpret_cast<unsigned short *>(smem_raw + tile_bins);           // tile_bins/4 uint16
    unsigned short *bufB = bufA + tile_bins / 4;                                               // tile_bins/16 uint16
    uint32_t *deepA = reinterpret_cast<uint32_t *>(smem_raw + tile_bins + tile_bins / 2 + tile_bins / 8);   // 256
    uint32_t *deepB = deepA + 256;                                                             // 64
    OUT *lut = reinterpret_cast<OUT *>(deepB + 64);                                            // [levels][256]
    __shared__ uint32_t s_seq, s_flag;
    __shared__ uint32_t s_fix[16], s_lead[16];
    __shared__ double s_sum[16];

    for (uint32_t i = threadIdx.x; i < tile_bins / 4; i += SW_THREADS) tab32[i] = 0;
    for (uint32_t i = threadIdx.x; i < 256 + 64; i += SW_THREADS) deepA[i] = 0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_seq = atomicAdd(queue, 1u);
            s_flag = 0;
        }
        __syncthreads();
        const uint32_t s = s_seq;
        if (s >= n_seqs) break;
        const uint64_t p0 = base_start[s], n = base_start[s + 1] - p0;
        const uint32_t hm = head_mask[s];
        if (n > 65535) {
            if (threadIdx.x == 0) {
                spill_list[atomicAdd(spill_count, 1u)] = s;
                atomicMax(spill_maxlen, (unsigned long long)n);
            }
            continue;
        }
        // per-level constants: emit threshold, divisor, the k-mer ending at base k-1, quotient LUT
        if (threadIdx.x >= (uint32_t)k_lo && threadIdx.x <= (uint32_t)k_hi) {
            const int kk = (int)threadIdx.x;
            const uint32_t lead = __popc(hm & ((kk > 1 ? (1u << (kk - 1)) : 1u) - 1u));
            s_lead[kk] = lead;
            s_sum[kk] = (double)(n > lead ? n - lead : 0);
            uint32_t fix = SW_NONE;
            if (kk < k_hi && n >= (uint64_t)kk) {
                uint32_t x = 0, y = 0;                           // full kk-mer of bases 0..kk-1, oldest in bit 0
                for (int t = 0; t < kk; ++t) {
                    const uint64_t p = p0 + t;
                    const uint2 w = planes[p >> 5];
                    const uint32_t b = (uint32_t)(p & 31);
                    x |= ((w.x >> b) & 1u) << t;
                    y |= ((w.y >> b) & 1u) << t;
                }
                fix = (y << kk) | x;
            }
            s_fix[kk] = fix;
        }
        __syncthreads();
        if constexpr (IS_FREQ) {
            for (int kk = k_lo; kk <= k_hi; ++kk)
                lut[(kk - k_lo) * 256 + threadIdx.x] = convert_count<OUT>(threadIdx.x, 0.0, s_sum[kk]);
        }
        SeqInfo q;
        q.p0 = p0;
        q.n = n;
        bool any_dirty = false;
        for (int kk = k_lo; kk <= k_hi; ++kk) any_dirty |= s_lead[kk] < (uint32_t)(kk - 1);

        const uint64_t pf = p0 + (uint64_t)(k_hi - 1), pe = p0 + n;
        const bool has_full = n >= (uint64_t)k_hi;
        const uint64_t g0 = pf >> 3, g1 = has_full ? ((pe - 1) >> 3) + 1 : g0;

        for (uint32_t t = 0; t < n_tiles; ++t) {
            // ---- scan: full k_hi-mers that fall into tile t
            for (uint64_t g = g0 + threadIdx.x; g < g1; g += SW_THREADS) {
                const uint64_t W = g >> 2;
                const uint2 hi = __ldg(planes + W);
                const uint2 lo = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
                const uint32_t sh0 = 8u * (uint32_t)(g & 3) + 32u - (uint32_t)(k_hi - 1);
                const uint32_t gx = (uint32_t)((((unsigned long long)hi.x << 32) | lo.x) >> sh0);
                const uint32_t gy = (uint32_t)((((unsigned long long)hi.y << 32) | lo.y) >> sh0);
                const uint64_t pbase = g << 3;
                uint32_t m = 0xffu;
                if (pbase < pf) m &= 0xffu << (uint32_t)(pf - pbase);
                if (pbase + 8 > pe) m &= 0xffu >> (uint32_t)(pbase + 8 - pe);
#pragma unroll 1
                for (uint32_t i = 0; i < tbits; ++i) {
                    const uint32_t want = 0u - ((t >> (tbits - 1 - i)) & 1u);
                    m &= ~((gy >> (uint32_t)(k_hi - 1 - i)) ^ want);
                }
                while (m) {
                    const uint32_t b = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1u;
                    const uint32_t yw = (gy >> b) & mask, xw = (gx >> b) & mask;
                    bump<uint8_t>(tab32, ((yw & ylow) << k_hi) | xw, &s_flag);
                }
            }
            __syncthreads();

            // ---- level chain: kk = k_hi (the uint8 tile itself), then pooled uint16 levels down to
            //      max(k_lo, k1); `cur` holds the slice [t*cur_bins, (t+1)*cur_bins) of level kk.
            //      One barrier per level: level kk-1 is pooled (into the other buffer) by threads that
            //      may still be overtaken by threads writing level kk out -- both only READ level kk.
            unsigned short *cur16 = nullptr;
            uint32_t cur_bins = tile_bins;
            for (int kk = k_hi; kk >= k_lo && kk >= k1; --kk) {
                if (kk < k_hi) {
                    const uint32_t nb = cur_bins >> 2, ncols_log = (uint32_t)kk;        // destination: 2^kk columns
                    unsigned short *dst = ((k_hi - kk) & 1) ? bufA : bufB;
                    const uint32_t fix = s_fix[kk];
                    const uint32_t fix_l = (fix != SW_NONE && fix / nb == t) ? fix % nb : SW_NONE;   // the k-mer ending at base kk-1
                    if (kk == k_hi - 1) {                                                 // from the uint8 tile
                        const unsigned char *P = reinterpret_cast<const unsigned char *>(tab32);
                        for (uint32_t i = threadIdx.x; i < nb; i += SW_THREADS) {
                            const uint32_t y = i >> ncols_log, x = i & ((1u << ncols_log) - 1u);
                            const uint32_t a = *reinterpret_cast<const unsigned short *>(P + ((2 * y) << (kk + 1)) + 2 * x);
                            const uint32_t b = *reinterpret_cast<const unsigned short *>(P + ((2 * y + 1) << (kk + 1)) + 2 * x);
                            dst[i] = (unsigned short)((a & 0xffu) + (a >> 8) + (b & 0xffu) + (b >> 8) + (i == fix_l ? 1u : 0u));
                        }
                    } else {
                        const unsigned short *P = cur16;
                        for (uint32_t i = threadIdx.x; i < nb; i += SW_THREADS) {
                            const uint32_t y = i >> ncols_log, x = i & ((1u << ncols_log) - 1u);
                            const uint32_t a = *reinterpret_cast<const uint32_t *>(P + ((2 * y) << (kk + 1)) + 2 * x);
                            const uint32_t b = *reinterpret_cast<const uint32_t *>(P + ((2 * y + 1) << (kk + 1)) + 2 * x);
                            dst[i] = (unsigned short)((a & 0xffffu) + (a >> 16) + (b & 0xffffu) + (b >> 16) + (i == fix_l ? 1u : 0u));
                        }
                    }
                    cur16 = dst;
                    cur_bins = nb;
                    __syncthreads();
                }
                const bool dirty = any_dirty && s_lead[kk] < (uint32_t)(kk - 1);   // CTA-uniform
                if (dirty) {
                    if (threadIdx.x == 0) {
                        q.lead = s_lead[kk];
                        head_partials(planes, q, kk, [&](uint32_t bin) {
                            if (bin / cur_bins != t) return;
                            const uint32_t lb = bin % cur_bins;
                            if (kk == k_hi) bump<uint8_t>(tab32, lb, &s_flag);
                            else cur16[lb] += 1;
                        });
                    }
                    __syncthreads();
                }
                // ---- write this level's slice
                OUT *dstrow = reinterpret_cast<OUT *>(so.out[kk]) + (uint64_t)s * ((uint64_t)1 << (2 * kk)) + (uint64_t)t * cur_bins;
                const OUT *lk = lut + (kk - k_lo) * 256;
                const double sum = s_sum[kk];
                const bool vec_ok = cur_bins % BPS == 0 && ((reinterpret_cast<uintptr_t>(dstrow) & 15) == 0);
                if (kk == k_hi && vec_ok) {
                    const uint32_t n_stores = cur_bins / BPS;
                    for (uint32_t c = threadIdx.x; c < n_stores; c += SW_THREADS) {
                        union {
                            uint4 v;
                            OUT e[BPS];
                        } u;
                        uint32_t w0 = 0, w1 = 0;
                        if constexpr (BPS == 8) {
                            const uint2 w = *reinterpret_cast<uint2 *>(&tab32[c * 2]);
                            w0 = w.x, w1 = w.y;
                        } else if constexpr (BPS == 4) {
                            w0 = tab32[c];
                        } else {
                            w0 = reinterpret_cast<unsigned short *>(tab32)[c];
                        }
                        if constexpr (IS_FREQ) {
#pragma unroll
                            for (int i = 0; i < BPS; ++i) u.e[i] = lk[((i < 4 ? w0 : w1) >> (8 * (i & 3))) & 0xffu];
                        } else if constexpr (BPS == 8) {
                            u.v = make_uint4(__byte_perm(w0, 0, 0x4140), __byte_perm(w0, 0, 0x4342),
                                             __byte_perm(w1, 0, 0x4140), __byte_perm(w1, 0, 0x4342));
                        } else {
                            u.v = make_uint4(w0 & 0xffu, (w0 >> 8) & 0xffu, (w0 >> 16) & 0xffu, w0 >> 24);
                        }
                        st_stream16(dstrow + (uint64_t)c * BPS, u.v);
                    }
                } else if (kk < k_hi && vec_ok) {
                    const uint32_t n_stores = cur_bins / BPS;
                    for (uint32_t c = threadIdx.x; c < n_stores; c += SW_THREADS) {
                        union {
                            uint4 v;
                            OUT e[BPS];
                        } u;
#pragma unroll
                        for (int i = 0; i < BPS; ++i) u.e[i] = sweep_convert<OUT>(cur16[c * BPS + i], lk, sum);
                        st_stream16(dstrow + (uint64_t)c * BPS, u.v);
                    }
                } else {
                    for (uint32_t i = threadIdx.x; i < cur_bins; i += SW_THREADS) {
                        const uint32_t c = kk == k_hi ? reinterpret_cast<const unsigned char *>(tab32)[i] : cur16[i];
                        dstrow[i] = sweep_convert<OUT>(c, lk, sum);
                    }
                }
                if (dirty) {                                      // remove the partial k-mers before pooling further
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        q.lead = s_lead[kk];
                        head_partials(planes, q, kk, [&](uint32_t bin) {
                            if (bin / cur_bins != t) return;
                            const uint32_t lb = bin % cur_bins;
                            if (kk == k_hi) atomicSub(&tab32[lb >> 2], 1u << ((lb & 3u) * 8u));
                            else cur16[lb] -= 1;
                        });
                    }
                    __syncthreads();
                }
            }
            // ---- below one row per tile: add this tile's order-k1 row into the order-(k1-1) table
            if (k_lo < k1) {
                const uint32_t cols = 1u << k1;                  // cur16 holds order k1: one row, y = t
                for (uint32_t x = threadIdx.x; x < cols; x += SW_THREADS)
                    atomicAdd(&deepA[(t >> 1) * (cols >> 1) + (x >> 1)], (uint32_t)cur16[x]);
            }
            __syncthreads();                                      // every reader of the tile is done
            for (uint32_t i = threadIdx.x; i < tile_bins / 4; i += SW_THREADS) tab32[i] = 0;
            __syncthreads();
        }

        // ---- orders below k1: tiny tables, pooled sequentially
        if (k_lo < k1) {
            uint32_t *tbl = deepA, *oth = deepB;
            for (int kk = k1 - 1; kk >= k_lo; --kk) {
                const uint32_t nb = 1u << (2 * kk);
                if (threadIdx.x == 0) {
                    if (s_fix[kk] != SW_NONE) tbl[s_fix[kk]] += 1;
                    if (s_lead[kk] < (uint32_t)(kk - 1)) {
                        q.lead = s_lead[kk];
                        head_partials(planes, q, kk, [&](uint32_t bin) { tbl[bin] += 1; });
                    }
                }
                __syncthreads();
                OUT *dstrow = reinterpret_cast<OUT *>(so.out[kk]) + (uint64_t)s * nb;
                for (uint32_t i = threadIdx.x; i < nb; i += SW_THREADS)
                    dstrow[i] = sweep_convert<OUT>(tbl[i], lut + (kk - k_lo) * 256, s_sum[kk]);
                __syncthreads();
                if (threadIdx.x == 0 && s_lead[kk] < (uint32_t)(kk - 1)) {
                    q.lead = s_lead[kk];
                    head_partials(planes, q, kk, [&](uint32_t bin) { tbl[bin] -= 1; });
                }
                __syncthreads();
                if (kk > k_lo) {
                    const uint32_t nd = nb >> 2;
                    for (uint32_t i = threadIdx.x; i < nd; i += SW_THREADS) {
                        const uint32_t y = i >> (kk - 1), x = i & ((1u << (kk - 1)) - 1u);
                        oth[i] = tbl[((2 * y) << kk) + 2 * x] + tbl[((2 * y) << kk) + 2 * x + 1] +
                                 tbl[((2 * y + 1) << kk) + 2 * x] + tbl[((2 * y + 1) << kk) + 2 * x + 1];
                    }
                    __syncthreads();
                    uint32_t *tmp = tbl;
                    tbl = oth;
                    oth = tmp;
                }
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < 256 + 64; i += SW_THREADS) deepA[i] = 0;
        }
        if (threadIdx.x == 0 && s_flag) {
            spill_list[atomicAdd(spill_count, 1u)] = s;
            atomicMax(spill_maxlen, (unsigned long long)n);
        }
    }
}


// ---- path L: long sequences, shared-memory tiles of packed uint16 counters ----------------------------
// Measured on B200 (scratch/atoms_bench.cu): shared-memory atomic increments on random bins run at
// ~12 per clock per SM (3.5e12/s per GPU), global RED.ADD at 0.66 (1.9e11/s) -- 18x apart.  So a
// megabase sequence is counted in shared memory too, even though its table does not fit: work item =
// (sequence, tile of 65536 bins = 128 KB of uint16 counters packed two to a word); the CTA scans the
// WHOLE sequence and increments the k-mers of its tile (4 passes over the planes at k=9, which cost
// 1.25 MB of L2 reads each against 5e6 increments).  Per thread: one plane word = 32 k-mer end
// positions; the in-tile mask of all 32 is formed with a few whole-word ops (the tile index is the
// y-bits of the newest bases), then one atomic per set bit.  Overflow of a uint16 counter is caught
// without looking at the atomics' results: the sum of the counters read out must equal the number of
// increments issued (every wrap lowers the sum by 65535 or 65536), else the sequence is redone on
// path B.  The tile is converted and streamed out with 16-byte stores exactly once.
constexpr int LG_THREADS = 1024;
constexpr int LG_TILE_LOG2 = 16;
constexpr int LG_MAX_K = 11;            // 64 passes at k=11 still beat RED.ADD; beyond that path B
int g_long_path = 1;                    // test hook: 0 = everything path A declines goes straight to path B

template <typename OUT>
__global__ void __launch_bounds__(LG_THREADS, 1) cgr_long_kernel(const uint2 *__restrict__ planes,
                                                                 const uint64_t *__restrict__ base_start,
                                                                 const uint32_t *__restrict__ head_mask,
                                                                 const uint32_t *__restrict__ list, uint32_t n_list,
                                                                 int k, OUT *__restrict__ out, uint64_t out_stride,
                                                                 uint32_t *__restrict__ queue2,
                                                                 uint32_t *__restrict__ ovf_flag,
                                                                 uint32_t *__restrict__ spill2_count,
                                                                 uint32_t *__restrict__ spill2_list,
                                                                 unsigned long long *__restrict__ spill2_maxlen) {
    constexpr bool IS_FREQ = !std::is_integral<OUT>::value;
    constexpr int BPS = 16 / (int)sizeof(OUT);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t s_item;
    __shared__ unsigned long long s_diff;
    __shared__ OUT s_lut[256];
    (void)s_lut;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tile_log2 = 2 * k < LG_TILE_LOG2 ? (uint32_t)(2 * k) : (uint32_t)LG_TILE_LOG2;
    const uint32_t tile_bins = 1u << tile_log2;
    const uint32_t tbits = (uint32_t)(2 * k) - tile_log2;            // log2(number of tiles)
    const uint32_t n_tiles = 1u << tbits;
    const uint32_t ylow = (1u << (tile_log2 - (uint32_t)k)) - 1u;    // y bits that index inside a tile
    const uint32_t tile_words = tile_bins >> 1;                      // k >= 2 here
    const uint32_t n_items = n_list * n_tiles;

    for (uint32_t i = threadIdx.x; i < tile_words; i += LG_THREADS) tab[i] = 0;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_item = atomicAdd(queue2, 1u);
            s_diff = 0;
        }
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const uint32_t li = item >> tbits, t = item & (n_tiles - 1u);
        const uint32_t s = list[li];
        const SeqInfo q = seq_info(base_start, head_mask, s, k);
        const double sum = (double)(q.n > q.lead ? q.n - q.lead : 0);
        if constexpr (IS_FREQ) {
            if (threadIdx.x < 256) s_lut[threadIdx.x] = convert_count<OUT>(threadIdx.x, 0.0, sum);
        }
        unsigned long long issued = 0;
        if (q.n >= (uint64_t)k) {
            const uint64_t pa = q.p0 + (uint64_t)(k - 1), pe = q.p0 + q.n;   // k-mer end positions [pa, pe)
            const uint64_t W0 = pa >> 5, W1 = ((pe - 1) >> 5) + 1;
            const uint32_t shc = 32u - (uint32_t)(k - 1);
            uint32_t cnt = 0;
            const uint32_t shc1 = shc - 1u;
            const uint32_t xm4 = ((1u << (k - 1)) - 1u) << 2;          // x bits 1..k-1 of the window, as a byte offset
            const uint32_t ym4 = ylow << (k + 1);                      // in-tile y bits above them
            for (uint64_t Wb = W0 + (threadIdx.x & ~31u); Wb < W1; Wb += LG_THREADS) {   // warp-uniform trip count
                const uint64_t W = Wb + lane;
                uint2 hi = make_uint2(0, 0);
                if (W < W1) hi = __ldg(planes + W);
                uint2 lo;
                lo.x = __shfl_up_sync(0xffffffffu, hi.x, 1);
                lo.y = __shfl_up_sync(0xffffffffu, hi.y, 1);
                if (lane == 0) lo = Wb ? __ldg(planes + Wb - 1) : make_uint2(0, 0);
                if (W < W1) {
                    const unsigned long long vx = ((unsigned long long)hi.x << 32) | lo.x;
                    const unsigned long long vy = ((unsigned long long)hi.y << 32) | lo.y;
                    const uint64_t pbase = W << 5;
                    uint32_t m = 0xffffffffu;
                    if (pbase < pa) m &= 0xffffffffu << (uint32_t)(pa - pbase);
                    if (pbase + 32 > pe) m &= 0xffffffffu >> (uint32_t)(pbase + 32 - pe);
                    // tile index = top tbits bits of the y window = y-bits of the newest tbits bases
#pragma unroll 1
                    for (uint32_t i = 0; i < tbits; ++i) {
                        const uint32_t want = 0u - ((t >> (tbits - 1 - i)) & 1u);
                        m &= ~((uint32_t)(vy >> (32u - i)) ^ want);
                    }
                    cnt += __popc(m);
                    // bin = (y window << k) | x window; counter word = bin >> 1, half = bin & 1 = x-bit of the
                    // OLDEST base of the k-mer, which for end position b is bit b of selw
                    const uint32_t selw = (uint32_t)(vx >> shc);
                    const unsigned long long vy2 = vy >> 1;
                    auto bin_off = [&](uint32_t b) {
                        const uint32_t s1 = b + shc1;
                        return ((uint32_t)(vx >> s1) & xm4) | (((uint32_t)(vy2 >> s1) << (k + 1)) & ym4);
                    };
                    while (m) {   // two independent increments per trip: the highest and the lowest set bit
                        uint32_t b1, b2;
                        asm("bfind.u32 %0, %1;" : "=r"(b1) : "r"(m));
                        const uint32_t bit1 = 1u << b1, bit2 = m & (0u - m);
                        asm("bfind.u32 %0, %1;" : "=r"(b2) : "r"(bit2));
                        m &= ~(bit1 | bit2);
                        const uint32_t off1 = bin_off(b1), off2 = bin_off(b2);
                        const uint32_t inc1 = (selw & bit1) ? 0x10000u : 1u;
                        const uint32_t inc2 = bit2 == bit1 ? 0u : ((selw & bit2) ? 0x10000u : 1u);   // one bit left: adds 0
                        atomicAdd(reinterpret_cast<uint32_t *>(smem_raw + off1), inc1);   // results unused
                        atomicAdd(reinterpret_cast<uint32_t *>(smem_raw + off2), inc2);
                    }
                }
                __syncwarp();   // lanes leave the bit loop at different times: reconverge before the next shuffle
            }
            issued = cnt;
        }
        if (threadIdx.x == 0 && q.lead < (uint32_t)(k - 1)) {   // dirty head (rare)
            head_partials(planes, q, k, [&](uint32_t bin) {
                if ((bin >> tile_log2) != t) return;
                const uint32_t lb = bin & (tile_bins - 1u);
                atomicAdd(&tab[lb >> 1], (lb & 1u) ? 0x10000u : 1u);
                issued += 1;
            });
        }
        __syncthreads();

        // ---- read-out: convert, stream, clear, and add up what the counters hold
        OUT *trow = out + (uint64_t)s * out_stride + (uint64_t)t * tile_bins;
        unsigned long long held = 0;
        const bool vec_ok = (tile_bins % BPS == 0) && ((reinterpret_cast<uintptr_t>(trow) & 15) == 0);
        if (vec_ok) {
            const uint32_t n_stores = tile_bins / BPS;
            for (uint32_t c = threadIdx.x; c < n_stores; c += LG_THREADS) {
                union {
                    uint4 v;
                    OUT e[BPS];
                } u;
                uint32_t cn[BPS];
                if constexpr (BPS == 8) {
                    const uint4 w = *reinterpret_cast<uint4 *>(&tab[c * 4]);
                    if (w.x | w.y | w.z | w.w) *reinterpret_cast<uint4 *>(&tab[c * 4]) = make_uint4(0, 0, 0, 0);
                    cn[0] = w.x & 0xffffu, cn[1] = w.x >> 16, cn[2] = w.y & 0xffffu, cn[3] = w.y >> 16;
                    cn[4] = w.z & 0xffffu, cn[5] = w.z >> 16, cn[6] = w.w & 0xffffu, cn[7] = w.w >> 16;
                    u.v = w;                                     // uint16 counts: the packed counters as they are
                } else if constexpr (BPS == 4) {
                    const uint2 w = *reinterpret_cast<uint2 *>(&tab[c * 2]);
                    if (w.x | w.y) *reinterpret_cast<uint2 *>(&tab[c * 2]) = make_uint2(0, 0);
                    cn[0] = w.x & 0xffffu, cn[1] = w.x >> 16, cn[2] = w.y & 0xffffu, cn[3] = w.y >> 16;
                } else {
                    const uint32_t w = tab[c];
                    if (w) tab[c] = 0;
                    cn[0] = w & 0xffffu, cn[1] = w >> 16;
                }
#pragma unroll
                for (int i = 0; i < BPS; ++i) held += cn[i];
                if constexpr (IS_FREQ) {
#pragma unroll
                    for (int i = 0; i < BPS; ++i) u.e[i] = cn[i] < 256u ? s_lut[cn[i]] : convert_count<OUT>(cn[i], 0.0, sum);
                } else if constexpr (BPS == 4) {
                    u.v = make_uint4(cn[0], cn[1], cn[2], cn[3]);
                }
                st_stream16(trow + (uint64_t)c * BPS, u.v);
            }
        } else {   // misaligned rows (odd strides)
            for (uint32_t i = threadIdx.x; i < tile_bins; i += LG_THREADS) {
                const uint32_t c = (tab[i >> 1] >> (16u * (i & 1u))) & 0xffffu;
                held += c;
                trow[i] = IS_FREQ ? (c < 256u ? s_lut[c] : convert_count<OUT>(c, 0.0, sum)) : convert_count<OUT>(c, 0.0, sum);
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < tile_words; i += LG_THREADS) tab[i] = 0;
        }
        // a wrapped counter makes held < issued: redo the sequence on path B (once, whichever tile sees it)
        const unsigned long long d = warp_sum64(issued - held);
        if (lane == 0 && d) atomicAdd(&s_diff, d);
        __syncthreads();
        if (threadIdx.x == 0 && s_diff != 0 && atomicExch(&ovf_flag[li], 1u) == 0u) {
            spill2_list[atomicAdd(spill2_count, 1u)] = s;
            atomicMax(spill2_maxlen, (unsigned long long)q.n);
        }
    }
}

// ---- path L, one pass (k <= 9): the whole 4^k table in shared memory --------------------------------
// cgr_long_kernel above is bound by what it spends AROUND an increment: at k = 9 only one k-mer in four
// falls into the tile a CTA holds, so every plane word is fetched, masked and bit-looped four times, the
// bit loop diverges, and ncu shows the ALU pipe at 79 % for 1.9 increments per clock and SM against the
// 12 the shared-memory atomics can do.  Here the table of ONE sequence is complete in the shared memory of
// ONE CTA, so every k-mer is an increment: no tile mask, no bit loop, no divergence -- each thread takes a
// plane word and issues its 32 increments straight-line.
//   k <= 8: 4^k uint16 counters packed two to a word (<= 128 KB), exactly the counters of the kernel above.
//   k == 9: 262144 counters do not fit at 16 or 8 bits (227 KB per CTA) but do at SIX: five 6-bit fields
//           per 32-bit word, word = bin / 5, 52429 words = 205 KB (l1_locate: one widening multiply gives the
//           word and the field's bit offset).  A field overflows
//           at 64 while the mean count of a 5 Mb genome is 19, so a field is kept below 32 + slack by its
//           own incrementers: ATOMS returns the old word (same rate as without, scratch/atoms_bench.cu);
//           the thread that sees its field at exactly 32 takes 32 back out (one more ATOMS) and adds 1 to
//           the sequence's uint32 side table in global memory (RED.ADD, once per 32 increments of a hot
//           bin).  count = field + 32 * side.  All word arithmetic is addition mod 2^32, so the final word
//           is right whatever the interleaving, PROVIDED no field ever carried into its neighbour; a
//           field can only reach 64 through an increment that observes 63 (by induction from field 0 up:
//           carries only come from a lower field that itself overflowed), and that thread raises the
//           abort flag -- the sequence is then redone by the tiled kernel above (uint16, no such limit).
//           It takes > 30 increments of ONE bin in flight between an ATOMS and its follow-up, i.e. a
//           homopolymer / tandem-repeat run of several hundred bases.  The sum of everything read out
//           must equal the number of k-mers as well (belt and braces, as above).
// Work item = sequence; one CTA per SM, persistent over the list.
constexpr uint32_t L1_MAX_CTAS = 160;   // side tables in the workspace: one per CTA
constexpr int L1_SIX_K = 9;             // the order whose table needs 6-bit fields; below it uint16 counters fit
int g_long_one_pass = 1;                // test hook: 0 = tiled kernel only, 2 = one pass however few the sequences

// shared-memory atomics by 32-bit shared address (the table base stays in one register instead of being
// re-derived from the generic pointer around every call)
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Field of a bin inside its word.  One 32 x 32 -> 64 multiply by ceil(2^32 / 5) gives both: the high word is
// bin / 5, the low word is (bin % 5) / 5 as a binary fraction (+ < 2^18), whose top five bits are
// 0, 6, 12, 19, 25 -- five bit offsets at least six apart, so the fields sit THERE (bits 18 and 31 unused) and
// no second multiply is needed for "6 * (bin % 5)".
constexpr uint32_t L1_DIV5 = 0x33333334u;
__device__ __forceinline__ void l1_locate(uint32_t bin, uint32_t &word, uint32_t &sh) {
    uint32_t lo;   // spelled in PTX: left to itself the compiler shifts the 64-bit product about instead
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}" : "=r"(lo), "=r"(word) : "r"(bin), "n"(L1_DIV5));
    sh = lo >> 27;
}

// G increments at once: the address arithmetic of the G bins interleaves, the G results are tested together.
// flags: bit 0 = the side table was used, bit 1 = a field carried into its neighbour (abort)
// s_blk: one bit per 64 bins, set where the side table holds something (the read-out looks only there)
template <bool SIX, int G>
__device__ __forceinline__ void l1_count(const uint32_t (&bin)[G], uint32_t tab_addr, uint32_t *side, uint32_t *s_blk,
                                         uint32_t &flags) {
    if constexpr (SIX) {
        uint32_t a[G], sh[G], old[G], hot = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            uint32_t q;
            l1_locate(bin[j], q, sh[j]);
            asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a[j]) : "r"(q), "r"(tab_addr));
            old[j] = atoms_add(a[j], 1u << sh[j]);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) hot |= (old[j] >> sh[j]) & 32u;
        if (hot) {   // some field already held >= 32 (rare on a bin of average count, routine on a hot bin)
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const uint32_t t = (old[j] >> sh[j]) & 63u;
                if (t >= 32u) {
                    if (t == 32u) {   // this thread brings the field back under 32 and credits the side table
                        reds_add(a[j], 0u - (32u << sh[j]));
                        atomicAdd(&side[bin[j]], 1u);
                        atomicOr(&s_blk[bin[j] >> 11], 1u << ((bin[j] >> 6) & 31u));
                        flags |= 1u;
                    } else if (t == 63u) {
                        flags |= 2u;   // this increment carried into the neighbouring field
                    }
                }
            }
        }
    } else {
        (void)side, (void)flags, (void)s_blk;
#pragma unroll
        for (int j = 0; j < G; ++j) reds_add(tab_addr + ((bin[j] << 1) & ~3u), 1u << ((bin[j] & 1u) << 4));
    }
}

template <typename OUT, bool SIX>
__global__ void __launch_bounds__(LG_THREADS, 1) cgr_long1_kernel(const uint2 *__restrict__ planes,
                                                                  const uint64_t *__restrict__ base_start,
                                                                  const uint32_t *__restrict__ head_mask,
                                                                  const uint32_t *__restrict__ list,
                                                                  const uint32_t *__restrict__ n_list_ptr, int k,
                                                                  OUT *__restrict__ out, uint64_t out_stride,
                                                                  uint32_t *__restrict__ queue2,
                                                                  uint32_t *__restrict__ side_tables,
                                                                  uint32_t *__restrict__ spill2_count,
                                                                  uint32_t *__restrict__ spill2_list,
                                                                  unsigned long long *__restrict__ spill2_maxlen) {
    constexpr bool IS_FREQ = !std::is_integral<OUT>::value;
    constexpr int BPS = 16 / (int)sizeof(OUT);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *tab = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t s_item;
    __shared__ uint32_t s_flags[2];   // [0] the side table was used, [1] abort
    __shared__ uint32_t s_blk[SIX ? (1u << (2 * L1_SIX_K - 11)) : 1];
    __shared__ unsigned long long s_diff;
    __shared__ OUT s_lut[256];
    (void)s_lut;
    const uint32_t n_list = *n_list_ptr;
    if (n_list == 0) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bins = 1u << (2 * k);
    const uint32_t n_words = SIX ? (bins + 4u) / 5u : bins >> 1;
    uint32_t *side = SIX ? side_tables + (size_t)blockIdx.x * bins : nullptr;
    const uint32_t mask = (1u << k) - 1u;
    const uint32_t shc = 32u - (uint32_t)(k - 1);
    const uint32_t ymul = 1u << (2 * k);   // mulhi by 2^(2k) = shift right by 32-2k, on the FMA pipe
    uint32_t tab_addr;   // opaque to the compiler, or it re-derives the shared window base around every group
    asm volatile("mov.u32 %0, %1;" : "=r"(tab_addr) : "r"((uint32_t)__cvta_generic_to_shared(tab)));

    for (uint32_t i = threadIdx.x; i < n_words; i += LG_THREADS) tab[i] = 0;
    if constexpr (SIX) {
        for (uint32_t i = threadIdx.x; i < (bins >> 11); i += LG_THREADS) s_blk[i] = 0;
        uint4 *s4 = reinterpret_cast<uint4 *>(side);
        for (uint32_t i = threadIdx.x; i < bins / 4; i += LG_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
    }

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_item = atomicAdd(queue2, 1u);
            s_diff = 0;
            s_flags[0] = 0;
            s_flags[1] = 0;
        }
        __syncthreads();
        const uint32_t li = s_item;
        if (li >= n_list) break;
        const uint32_t s = list[li];
        const SeqInfo q = seq_info(base_start, head_mask, s, k);
        const double sum = (double)(q.n > q.lead ? q.n - q.lead : 0);
        if constexpr (IS_FREQ) {
            if (threadIdx.x < 256) s_lut[threadIdx.x] = convert_count<OUT>(threadIdx.x, 0.0, sum);
        }
        unsigned long long issued = 0;
        uint32_t flags = 0;
        if (q.n >= (uint64_t)k) {
            const uint64_t pa = q.p0 + (uint64_t)(k - 1), pe = q.p0 + q.n;   // k-mer end positions [pa, pe)
            if (threadIdx.x == 0) issued = pe - pa;
            const uint64_t Wf0 = (pa + 31) >> 5, Wf1 = pe >> 5;              // plane words lying wholly inside
            // x window of end position b: bits b+shc.. of (hi.x:lo.x); y window pushed to the top of a register by
            // a funnel shift left of 31-b, brought down to bits k..2k-1 (zeros above), merged under the x window
            auto bin_of = [&](unsigned long long vxs, uint32_t ylo, uint32_t yhi, uint32_t b) {
                const uint32_t xw = (uint32_t)(vxs >> b);
                const uint32_t yw = __umulhi(__funnelshift_l(ylo, yhi, 31u - b), ymul);
                return (xw & mask) | (yw & ~mask);
            };
            constexpr int G = SIX ? 8 : 8;
            const uint64_t Wfirst = Wf0 + (threadIdx.x & ~31u);
            uint2 nxt = make_uint2(0, 0), nxt0 = make_uint2(0, 0);   // the next trip's words are fetched a trip ahead
            if (Wfirst + lane < Wf1) nxt = __ldg(planes + Wfirst + lane);
            if (lane == 0 && Wfirst < Wf1 && Wfirst) nxt0 = __ldg(planes + Wfirst - 1);
            for (uint64_t Wb = Wfirst; Wb < Wf1; Wb += LG_THREADS) {   // warp-uniform trips
                const uint64_t W = Wb + lane;
                const uint2 hi = nxt;
                uint2 lo;
                lo.x = __shfl_up_sync(0xffffffffu, hi.x, 1);
                lo.y = __shfl_up_sync(0xffffffffu, hi.y, 1);
                if (lane == 0) lo = nxt0;
                const uint64_t Wn = W + LG_THREADS;
                nxt = make_uint2(0, 0);
                if (Wn < Wf1) nxt = __ldg(planes + Wn);
                if (lane == 0 && Wn < Wf1) nxt0 = __ldg(planes + Wn - 1);
                if (W < Wf1) {
                    const unsigned long long vxs = (((unsigned long long)hi.x << 32) | lo.x) >> shc;
#pragma unroll
                    for (uint32_t b = 0; b < 32; b += G) {
                        uint32_t bin[G];
#pragma unroll
                        for (int j = 0; j < G; ++j) bin[j] = bin_of(vxs, lo.y, hi.y, b + j);
                        l1_count<SIX, G>(bin, tab_addr, side, s_blk, flags);
                    }
                }
                __syncwarp();
            }
            // the (at most two) ragged words at the ends: warp 0, lane = end position
            if (threadIdx.x < 32) {
                auto edge = [&](uint64_t W, uint32_t m) {
                    const uint2 hi = __ldg(planes + W);
                    const uint2 lo = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
                    if ((m >> lane) & 1u) {
                        const unsigned long long vxs = (((unsigned long long)hi.x << 32) | lo.x) >> shc;
                        const uint32_t bin[1] = {bin_of(vxs, lo.y, hi.y, lane)};
                        l1_count<SIX, 1>(bin, tab_addr, side, s_blk, flags);
                    }
                };
                const uint32_t a5 = (uint32_t)(pa & 31), e5 = (uint32_t)(pe & 31);
                if (Wf0 > Wf1) {
                    edge(pa >> 5, (0xffffffffu << a5) & ~(0xffffffffu << e5));   // a5 < e5, both inside one word
                } else {
                    if (a5) edge(Wf0 - 1, 0xffffffffu << a5);
                    if (e5) edge(Wf1, ~(0xffffffffu << e5));
                }
            }
        }
        if (threadIdx.x == 0 && q.lead < (uint32_t)(k - 1)) {   // dirty head (rare)
            head_partials(planes, q, k, [&](uint32_t bin) {
                const uint32_t b1[1] = {bin};
                l1_count<SIX, 1>(b1, tab_addr, side, s_blk, flags);
                issued += 1;
            });
        }
        if (flags & 1u) s_flags[0] = 1;
        if (flags & 2u) s_flags[1] = 1;
        __syncthreads();
        const bool use_side = SIX && s_flags[0] != 0;
        const bool aborted = SIX && s_flags[1] != 0;

        // ---- read-out: convert, stream out, add up what the counters hold
        OUT *row = out + (uint64_t)s * out_stride;
        unsigned long long held = 0;
        const bool vec_ok = (bins % BPS == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
        auto counter = [&](uint32_t bin) -> uint32_t {
            if constexpr (SIX) {
                uint32_t qq, sh;
                l1_locate(bin, qq, sh);
                return (tab[qq] >> sh) & 63u;
            } else {
                return (tab[bin >> 1] >> ((bin & 1u) << 4)) & 0xffffu;
            }
        };
        if (!aborted) {
            if (vec_ok) {
                const uint32_t n_stores = bins / BPS;
                for (uint32_t c = threadIdx.x; c < n_stores; c += LG_THREADS) {
                    union {
                        uint4 v;
                        OUT e[BPS];
                    } u;
                    uint32_t cn[BPS];
#pragma unroll
                    for (int i = 0; i < BPS; ++i) cn[i] = counter(c * BPS + i);
                    if constexpr (SIX) {
                        const uint32_t b0 = c * BPS;
                        if (use_side && ((s_blk[b0 >> 11] >> ((b0 >> 6) & 31u)) & 1u)) {   // BPS divides 64
                            uint32_t h[BPS];
                            if constexpr (BPS == 8) {
                                const uint4 h0 = *reinterpret_cast<const uint4 *>(side + b0);
                                const uint4 h1 = *reinterpret_cast<const uint4 *>(side + b0 + 4);
                                h[0] = h0.x, h[1] = h0.y, h[2] = h0.z, h[3] = h0.w, h[4] = h1.x, h[5] = h1.y, h[6] = h1.z, h[7] = h1.w;
                            } else if constexpr (BPS == 4) {
                                const uint4 h0 = *reinterpret_cast<const uint4 *>(side + b0);
                                h[0] = h0.x, h[1] = h0.y, h[2] = h0.z, h[3] = h0.w;
                            } else {
                                const uint2 h0 = *reinterpret_cast<const uint2 *>(side + b0);
                                h[0] = h0.x, h[1] = h0.y;
                            }
                            uint32_t any = 0;
#pragma unroll
                            for (int i = 0; i < BPS; ++i) cn[i] += h[i] << 5, any |= h[i];
                            if (any) {
#pragma unroll
                                for (int i = 0; i < BPS; ++i) side[b0 + i] = 0;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < BPS; ++i) held += cn[i];
#pragma unroll
                    for (int i = 0; i < BPS; ++i) {
                        if constexpr (IS_FREQ) u.e[i] = cn[i] < 256u ? s_lut[cn[i]] : convert_count<OUT>(cn[i], 0.0, sum);
                        else u.e[i] = (OUT)cn[i];
                    }
                    st_stream16(row + (uint64_t)c * BPS, u.v);
                }
            } else {   // misaligned rows (odd strides)
                for (uint32_t i = threadIdx.x; i < bins; i += LG_THREADS) {
                    uint32_t c = counter(i);
                    if constexpr (SIX) {
                        if (use_side) {
                            c += side[i] << 5;
                            side[i] = 0;
                        }
                    }
                    held += c;
                    row[i] = IS_FREQ ? (c < 256u ? s_lut[c] : convert_count<OUT>(c, 0.0, sum)) : convert_count<OUT>(c, 0.0, sum);
                }
            }
        } else if constexpr (SIX) {
            if (use_side) {
                uint4 *s4 = reinterpret_cast<uint4 *>(side);
                for (uint32_t i = threadIdx.x; i < bins / 4; i += LG_THREADS) s4[i] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_words; i += LG_THREADS) tab[i] = 0;
        if constexpr (SIX) {
            if (use_side)
                for (uint32_t i = threadIdx.x; i < (bins >> 11); i += LG_THREADS) s_blk[i] = 0;
        }
        // a wrapped / carried counter: redo the sequence elsewhere
        const unsigned long long d = warp_sum64(issued - held);
        if (lane == 0 && d) atomicAdd(&s_diff, d);
        __syncthreads();
        if (threadIdx.x == 0 && (aborted || s_diff != 0)) {
            spill2_list[atomicAdd(spill2_count, 1u)] = s;
            atomicMax(spill2_maxlen, (unsigned long long)q.n);
        }
    }
}

// ---- path B ------------------------------------------------------------------------------------
// grid (chunks, batch): CTA (c, b) counts bases [c*B_CHUNK, (c+1)*B_CHUNK) of sequence
// spill_list[first + b] into tables[b] (4^k uint32, zeroed beforehand) with RED.ADD.
__global__ void __launch_bounds__(B_THREADS) cgr_spill_count_kernel(const uint2 *__restrict__ planes,
                                                                    const uint64_t *__restrict__ base_start,
                                                                    const uint32_t *__restrict__ head_mask,
                                                                    const uint32_t *__restrict__ spill_list,
                                                                    uint32_t first, int k,
                                                                    uint32_t *__restrict__ tables,
                                                                    uint64_t table_stride) {
    const uint32_t s = spill_list[first + blockIdx.y];
    const SeqInfo q = seq_info(base_start, head_mask, s, k);
    uint32_t *tab = tables + (uint64_t)blockIdx.y * table_stride;
    const uint32_t mask = (1u << k) - 1u;
    const uint64_t c0 = (uint64_t)blockIdx.x * B_CHUNK;
    if (c0 >= q.n) return;
    const uint64_t c1 = c0 + B_CHUNK < q.n ? c0 + B_CHUNK : q.n;
    // this CTA emits k-mers ending at sequence bases [c0, c1) that are full and >= lead
    uint64_t lo = c0;
    if (lo < (uint64_t)(k - 1)) lo = (uint64_t)(k - 1);
    if (lo < q.lead) lo = q.lead;
    const uint64_t pa = q.p0 + lo, pe = q.p0 + c1;
    if (blockIdx.x == 0 && threadIdx.x == 0 && q.lead < (uint32_t)(k - 1))
        head_partials(planes, q, k, [&](uint32_t bin) { atomicAdd(&tab[bin], 1u); });
    if (pa >= pe) return;
    const uint64_t g0 = pa >> 3, g1 = ((pe - 1) >> 3) + 1;
    for (uint64_t g = g0 + threadIdx.x; g < g1; g += B_THREADS) {
        const uint64_t W = g >> 2;
        const uint2 hi = __ldg(planes + W);
        const uint2 lw = W ? __ldg(planes + W - 1) : make_uint2(0, 0);
        const unsigned long long vx = ((unsigned long long)hi.x << 32) | lw.x;
        const unsigned long long vy = ((unsigned long long)hi.y << 32) | lw.y;
        const uint64_t pbase = g << 3;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint64_t p = pbase + b;
            if (p < pa || p >= pe) continue;
            const uint32_t sh = (uint32_t)(p - (uint64_t)(k - 1) + 32 - (W << 5));
            atomicAdd(&tab[kmer_bin(vx, vy, sh, k, mask)], 1u);   // result unused -> RED.ADD
        }
    }
}

// per-table sum of the (wrapped) counts, needed as the frequency divisor
__global__ void __launch_bounds__(256) cgr_spill_sum_kernel(const uint32_t *__restrict__ tables, uint64_t table_stride,
                                                            uint32_t bins, uint32_t wrapmask,
                                                            unsigned long long *__restrict__ sums) {
    const uint32_t *tab = tables + (uint64_t)blockIdx.y * table_stride;
    unsigned long long part = 0;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < bins; i += gridDim.x * 256) part += tab[i] & wrapmask;
    part = warp_sum64(part);
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&sums[blockIdx.y], part);
}

template <typename OUT>
__global__ void __launch_bounds__(256) cgr_spill_finalize_kernel(const uint32_t *__restrict__ tables,
                                                                 uint64_t table_stride, uint32_t bins,
                                                                 uint32_t wrapmask,
                                                                 const unsigned long long *__restrict__ sums,
                                                                 const uint32_t *__restrict__ spill_list,
                                                                 uint32_t first, OUT *__restrict__ out,
                                                                 uint64_t out_stride) {
    const uint32_t *tab = tables + (uint64_t)blockIdx.y * table_stride;
    OUT *row = out + (uint64_t)spill_list[first + blockIdx.y] * out_stride;
    const double sum = (double)sums[blockIdx.y];
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < bins; i += gridDim.x * 256)
        row[i] = convert_count<OUT>(tab[i] & wrapmask, 0.0, sum);
}

struct KmerWs {
    uint32_t *queue;                 // [1]
    uint32_t *spill_count;           // [1]
    unsigned long long *spill_maxlen;// [1]
    uint32_t *spill_list;            // [n_seqs]
    uint32_t *queue2;                // [1]   path L: work-item queue, and the list of sequences it hands to path B
    uint32_t *spill2_count;          // [1]
    unsigned long long *spill2_maxlen;// [1]
    uint32_t *spill2_list;           // [n_seqs]
    uint32_t *ovf_flag;              // [n_seqs]
    uint32_t *queue3;                // [1]   the tiled path L kernel when it runs behind the one-pass kernel
    uint32_t *spill3_count;          // [1]
    unsigned long long *spill3_maxlen;// [1]
    uint32_t *spill3_list;           // [n_seqs]
    unsigned long long *sums;        // [batch]
    uint32_t *tables;                // [batch][4^k]; the one-pass path L kernel's side tables alias it
    uint32_t batch;
};

constexpr size_t SPILL_TABLE_BUDGET = 48ull << 20;   // keep the live uint32 tables well inside the 126 MB L2

inline uint32_t spill_batch(int k) {
    const size_t tb = ((size_t)1 << (2 * k)) * 4;
    size_t b = SPILL_TABLE_BUDGET / tb;
    if (b < 1) b = 1;
    if (b > 1024) b = 1024;
    return (uint32_t)b;
}

inline size_t ws_layout(uint32_t n_seqs, int k, void *base, KmerWs *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? (char *)base + off : nullptr;
        off += kam_align_up(bytes, 256);
        return p;
    };
    const uint32_t batch = spill_batch(k);
    void *hdr = take(64);
    void *list = take((size_t)(n_seqs ? n_seqs : 1) * 4);
    void *list2 = take((size_t)(n_seqs ? n_seqs : 1) * 4);
    void *flags = take((size_t)(n_seqs ? n_seqs : 1) * 4);
    void *list3 = take((size_t)(n_seqs ? n_seqs : 1) * 4);
    void *sums = take((size_t)1024 * 8);   // spill_batch() never exceeds 1024 (any k: the sweep reuses this layout)
    size_t n_tables = batch;
    if (k == L1_SIX_K) {   // one uint32 side table per CTA of cgr_long1_kernel<.., true>
        const size_t ctas = n_seqs < L1_MAX_CTAS ? (n_seqs ? n_seqs : 1) : L1_MAX_CTAS;
        if (ctas > n_tables) n_tables = ctas;
    }
    void *tables = take(n_tables * (((size_t)1 << (2 * k)) * 4));
    if (w) {
        w->queue = (uint32_t *)hdr;
        w->spill_count = (uint32_t *)hdr + 1;
        w->spill_maxlen = (unsigned long long *)((char *)hdr + 8);
        w->spill_list = (uint32_t *)list;
        w->queue2 = (uint32_t *)hdr + 4;
        w->spill2_count = (uint32_t *)hdr + 5;
        w->spill2_maxlen = (unsigned long long *)((char *)hdr + 24);
        w->spill2_list = (uint32_t *)list2;
        w->ovf_flag = (uint32_t *)flags;
        w->queue3 = (uint32_t *)hdr + 8;
        w->spill3_count = (uint32_t *)hdr + 9;
        w->spill3_maxlen = (unsigned long long *)((char *)hdr + 40);
        w->spill3_list = (uint32_t *)list3;
        w->sums = (unsigned long long *)sums;
        w->tables = (uint32_t *)tables;
        w->batch = batch;
    }
    return off;
}

template <typename CT, typename OUT, int THREADS>
int launch_tile(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, uint32_t n_seqs, int k,
                int count_bits, OUT *out, uint64_t out_stride, const KmerWs &w, cudaStream_t st) {
    const uint32_t bins = 1u << (2 * k);
    const uint32_t tile_bins = bins < TILE_BINS ? bins : TILE_BINS;
    size_t smem = (size_t)tile_bins * sizeof(CT);
    if (smem < 16) smem = 16;
    auto kern = cgr_tile_kernel<CT, OUT, THREADS>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    KAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)kam_sm_count() * per_sm;   // persistent: a whole number of CTAs per SM
    if (grid > n_seqs) grid = n_seqs;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(planes, base_start, head_mask, n_seqs, k, count_bits, out,
                                                  out_stride, w.queue, w.spill_count, w.spill_list, w.spill_maxlen,
                                                  g_word_scan);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

template <typename OUT>
int launch_warp(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, uint32_t n_seqs, int k,
                OUT *out, uint64_t out_stride, const KmerWs &w, cudaStream_t st) {
    const uint32_t bins = 1u << (2 * k);
    const size_t per_warp = (bins < 16u ? 16u : bins) + (std::is_integral<OUT>::value ? 0 : 256 * sizeof(OUT));
    const size_t smem = per_warp * WK_WARPS;
    auto kern = cgr_warp_kernel<OUT>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    KAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WK_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)kam_sm_count() * per_sm;   // persistent: a whole number of CTAs per SM
    const uint64_t need = ((uint64_t)n_seqs + WK_QB * WK_WARPS - 1) / (WK_QB * WK_WARPS);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, WK_WARPS * 32, smem, st>>>(planes, base_start, head_mask, n_seqs, k, out, out_stride, w.queue,
                                                      w.spill_count, w.spill_list, w.spill_maxlen);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

struct SpillInfo {
    uint32_t queue, count;
    unsigned long long maxlen;
};

inline int read_spill(const KmerWs &w, SpillInfo *h, cudaStream_t st) {
    KAM_CUDA(cudaMemcpyAsync(h, w.queue, 16, cudaMemcpyDeviceToHost, st));
    KAM_CUDA(cudaStreamSynchronize(st));
    return KAM_OK;
}

// path B over the spill list, `batch` sequences at a time so the live tables stay in L2
template <typename OUT>
int run_spill(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, int k, int count_bits,
              OUT *out, uint64_t out_stride, const KmerWs &w, const uint32_t *list, const SpillInfo &h,
              cudaStream_t st) {
    if (h.count == 0) return KAM_OK;
    const uint32_t bins = 1u << (2 * k);
    const uint32_t wrapmask = count_bits == 16 ? 0xffffu : 0xffffffffu;
    const uint32_t chunks = (uint32_t)((h.maxlen + B_CHUNK - 1) / B_CHUNK);
    uint32_t fin_blocks = bins / 256 / 8;
    if (fin_blocks < 1) fin_blocks = 1;
    if (fin_blocks > 1024) fin_blocks = 1024;
    const uint32_t batch = spill_batch(k);
    for (uint32_t first = 0; first < h.count; first += batch) {
        const uint32_t nb = h.count - first < batch ? h.count - first : batch;
        KAM_CUDA(cudaMemsetAsync(w.tables, 0, (size_t)nb * bins * 4, st));
        KAM_CUDA(cudaMemsetAsync(w.sums, 0, (size_t)nb * 8, st));
        cgr_spill_count_kernel<<<dim3(chunks ? chunks : 1, nb), B_THREADS, 0, st>>>(
            planes, base_start, head_mask, list, first, k, w.tables, bins);
        cgr_spill_sum_kernel<<<dim3(fin_blocks, nb), 256, 0, st>>>(w.tables, bins, bins, wrapmask, w.sums);
        cgr_spill_finalize_kernel<OUT><<<dim3(fin_blocks, nb), 256, 0, st>>>(w.tables, bins, bins, wrapmask, w.sums,
                                                                             list, first, out, out_stride);
        KAM_CUDA(cudaGetLastError());
    }
    return KAM_OK;
}

// the tiled path L kernel over `list` (h.count entries); sequences whose uint16 counters wrap go to `out_*`
template <typename OUT>
int launch_long_tiled(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, int k, OUT *out,
                      uint64_t out_stride, const KmerWs &w, const uint32_t *list, uint32_t count, uint32_t *queue,
                      uint32_t *out_count, uint32_t *out_list, unsigned long long *out_maxlen, cudaStream_t st) {
    KAM_CUDA(cudaMemsetAsync(w.ovf_flag, 0, (size_t)count * 4, st));
    const uint32_t tile_log2 = 2 * k < LG_TILE_LOG2 ? 2 * k : LG_TILE_LOG2;
    const size_t smem = ((size_t)1 << tile_log2) * 2;
    const uint64_t items = (uint64_t)count << (2 * k - tile_log2);
    uint64_t grid = (uint64_t)kam_sm_count();   // 128 KB of counters: one CTA per SM, persistent over the item queue
    if (grid > items) grid = items;
    auto kern = cgr_long_kernel<OUT>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, LG_THREADS, smem, st>>>(planes, base_start, head_mask, list, count, k, out, out_stride, queue,
                                                   w.ovf_flag, out_count, out_list, out_maxlen);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

// the one-pass path L kernel over the spill list of path A; the number of entries is read on the device
// (count_hint only sizes the grid), so the launch does not have to wait for path A's read-back
template <typename OUT>
int launch_long_one_pass(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, int k, OUT *out,
                         uint64_t out_stride, const KmerWs &w, uint32_t count_hint, cudaStream_t st) {
    const bool six = k == L1_SIX_K;
    const uint32_t bins = 1u << (2 * k);
    size_t smem = six ? (size_t)((bins + 4) / 5) * 4 : (size_t)bins * 2;
    smem = kam_align_up(smem, 16);
    uint64_t grid = (uint64_t)kam_sm_count();
    if (grid > L1_MAX_CTAS) grid = L1_MAX_CTAS;
    if (grid > count_hint) grid = count_hint;
    auto kern = six ? cgr_long1_kernel<OUT, true> : cgr_long1_kernel<OUT, false>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, LG_THREADS, smem, st>>>(planes, base_start, head_mask, w.spill_list, w.spill_count, k, out,
                                                   out_stride, w.queue2, w.tables, w.spill2_count, w.spill2_list,
                                                   w.spill2_maxlen);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

inline int read_hdr(const uint32_t *p, SpillInfo *h, cudaStream_t st) {
    KAM_CUDA(cudaMemcpyAsync(h, p, 16, cudaMemcpyDeviceToHost, st));
    KAM_CUDA(cudaStreamSynchronize(st));
    return KAM_OK;
}

// whether a batch of `count` long sequences at order k is better off with the one-pass kernel: at k = 9 it
// occupies one SM per sequence where the tiled kernel has four work items per sequence
inline bool one_pass_pays(int k, uint64_t count) {
    if (k < 7 || k > L1_SIX_K || !g_long_path || !g_long_one_pass) return false;
    return k < L1_SIX_K || g_long_one_pass == 2 || count * 2 >= (uint64_t)kam_sm_count();
}

// stages behind the one-pass kernel: h2 = what it gave up on
template <typename OUT>
int run_after_one_pass(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, int k, int count_bits,
                       OUT *out, uint64_t out_stride, const KmerWs &w, const SpillInfo &h2, cudaStream_t st) {
    if (h2.count == 0) return KAM_OK;
    if (k < L1_SIX_K)   // uint16 counters wrapped: the tiled kernel has the same counters
        return run_spill<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, w.spill2_list, h2, st);
    int rc;
    if ((rc = launch_long_tiled<OUT>(planes, base_start, head_mask, k, out, out_stride, w, w.spill2_list, h2.count,
                                     w.queue3, w.spill3_count, w.spill3_list, w.spill3_maxlen, st)))
        return rc;
    SpillInfo h3;
    if ((rc = read_hdr(w.queue3, &h3, st))) return rc;
    return run_spill<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, w.spill3_list, h3, st);
}

// the sequences path A declined (too long for its tiles, or a uint8 counter overflowed): path L where it
// applies -- in one pass with the whole table in shared memory for k <= 9, what that kernel gives up on
// (k = 9: a run of one k-mer that outpaces the 6-bit fields) and k = 10, 11 in tiles -- and path B for
// whatever overflows even uint16 counters (one 16-byte read-back per stage that found something)
template <typename OUT>
int run_overflow(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, int k, int count_bits,
                 OUT *out, uint64_t out_stride, const KmerWs &w, const SpillInfo &h, cudaStream_t st) {
    if (h.count == 0) return KAM_OK;
    if (k < 7 || k > LG_MAX_K || !g_long_path)
        return run_spill<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, w.spill_list, h, st);
    int rc;
    KAM_CUDA(cudaMemsetAsync(w.queue2, 0, 48, st));   // the queues / counts / maxima of stages two and three
    if (one_pass_pays(k, h.count)) {
        if ((rc = launch_long_one_pass<OUT>(planes, base_start, head_mask, k, out, out_stride, w, h.count, st))) return rc;
        SpillInfo h2;
        if ((rc = read_hdr(w.queue2, &h2, st))) return rc;
        return run_after_one_pass<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, h2, st);
    }
    if ((rc = launch_long_tiled<OUT>(planes, base_start, head_mask, k, out, out_stride, w, w.spill_list, h.count,
                                     w.queue2, w.spill2_count, w.spill2_list, w.spill2_maxlen, st)))
        return rc;
    SpillInfo h2;
    if ((rc = read_hdr(w.queue2, &h2, st))) return rc;
    return run_spill<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, w.spill2_list, h2, st);
}

template <typename OUT>
int run_kmer(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, uint32_t n_seqs, int k,
             int count_bits, OUT *out, uint64_t out_stride, const KmerWs &w, uint64_t total_bases_hint,
             cudaStream_t st) {
    KAM_CUDA(cudaMemsetAsync(w.queue, 0, 64, st));
    // mean sequence length picks the CTA size: the caller's hint, else one 8-byte read of the last offset
    unsigned long long total_bases = total_bases_hint;
    if (total_bases == 0) {
        KAM_CUDA(cudaMemcpyAsync(&total_bases, base_start + n_seqs, 8, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
    }
    const bool short_reads = total_bases / n_seqs < 2048;
    int rc;
    if (k <= 6 && short_reads && g_warp_path)
        rc = launch_warp<OUT>(planes, base_start, head_mask, n_seqs, k, out, out_stride, w, st);
    else if (k <= 6)
        rc = short_reads ? launch_tile<uint32_t, OUT, 64>(planes, base_start, head_mask, n_seqs, k, count_bits, out,
                                                          out_stride, w, st)
                         : launch_tile<uint32_t, OUT, 256>(planes, base_start, head_mask, n_seqs, k, count_bits, out,
                                                           out_stride, w, st);
    else
        rc = short_reads ? launch_tile<uint8_t, OUT, 64>(planes, base_start, head_mask, n_seqs, k, count_bits, out,
                                                         out_stride, w, st)
                         : launch_tile<uint8_t, OUT, 256>(planes, base_start, head_mask, n_seqs, k, count_bits, out,
                                                          out_stride, w, st);
    if (rc) return rc;
    if (k <= 6 && !(short_reads && g_warp_path)) return KAM_OK;   // uint32 counters, single tile: nothing can spill

    // A batch of megabase sequences: path A has just put (nearly) all of them on its list, so the one-pass
    // path L kernel is launched right behind it -- it reads the length of the list on the device -- and the
    // host reads both kernels' verdicts back in one go instead of waiting in between.
    if (total_bases / n_seqs > A_MAX_LEN && one_pass_pays(k, n_seqs)) {
        if ((rc = launch_long_one_pass<OUT>(planes, base_start, head_mask, k, out, out_stride, w, n_seqs, st))) return rc;
        SpillInfo h2;
        if ((rc = read_hdr(w.queue2, &h2, st))) return rc;
        return run_after_one_pass<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, h2, st);
    }
    SpillInfo h;
    if ((rc = read_spill(w, &h, st))) return rc;
    return run_overflow<OUT>(planes, base_start, head_mask, k, count_bits, out, out_stride, w, h, st);
}

template <typename OUT>
int run_sweep(const uint2 *planes, const uint64_t *base_start, const uint32_t *head_mask, uint32_t n_seqs, int k_lo,
              int k_hi, int count_bits, void *const *outs, const KmerWs &w, cudaStream_t st) {
    SweepPtrs so;
    for (int i = 0; i < 16; ++i) so.out[i] = nullptr;
    for (int k = k_lo; k <= k_hi; ++k) so.out[k] = outs[k - k_lo];
    KAM_CUDA(cudaMemsetAsync(w.queue, 0, 64, st));
    const bool is_freq = !std::is_integral<OUT>::value;
    const size_t tile_bins = (size_t)1 << (2 * k_hi < TILE_BINS_LOG2 ? 2 * k_hi : TILE_BINS_LOG2);
    const size_t smem = tile_bins + tile_bins / 2 + tile_bins / 8 + 320 * 4 +
                        (is_freq ? (size_t)(k_hi - k_lo + 1) * 256 * sizeof(OUT) : 16);
    auto kern = cgr_sweep_kernel<OUT>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    KAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SW_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)kam_sm_count() * per_sm;
    if (grid > n_seqs) grid = n_seqs;
    kern<<<(unsigned)grid, SW_THREADS, smem, st>>>(planes, base_start, head_mask, n_seqs, k_hi, k_lo, so, w.queue,
                                                   w.spill_count, w.spill_list, w.spill_maxlen);
    KAM_CUDA(cudaGetLastError());
    SpillInfo h;
    int rc = read_spill(w, &h, st);
    if (rc) return rc;
    for (int k = k_lo; k <= k_hi && h.count; ++k)      // sequences the fused kernel declined: every order, path B
        if ((rc = run_overflow<OUT>(planes, base_start, head_mask, k, count_bits, (OUT *)outs[k - k_lo],
                                 (uint64_t)1 << (2 * k), w, h, st)))
            return rc;
    return KAM_OK;
}

}  // namespace

extern "C" {

size_t kam_kmer_workspace_bytes(uint32_t n_seqs, int k) {
    if (k < 1 || k > 15) return 0;
    return ws_layout(n_seqs, k, nullptr, nullptr);
}

int kam_kmer_count(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                   uint32_t n_seqs, uint64_t total_bases_hint, int k, int count_bits, int out_kind, void *d_out,
                   uint64_t out_stride, void *d_ws, size_t ws_bytes, void *stream) {
    if (k < 1 || 2 * k > 64) {
        kam_set_error("k is too large to fit in the index");   // the reference's message (cgr @0x10000e15f)
        return KAM_E_INVALID;
    }
    if (k > 15) {
        kam_set_error("k=%d: 4^k bins exceed what this build addresses (k <= 15)", k);
        return KAM_E_UNSUPPORTED;
    }
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind >= KAM_OUT_COUNTS && out_kind <= KAM_OUT_FREQ_F64, "bad out_kind");
    KAM_REQUIRE(d_planes && d_base_start && d_head_mask && d_out && d_ws, "null pointer");
    KAM_REQUIRE(out_stride >= (1ull << (2 * k)), "out_stride smaller than 4^k");
    if (ws_bytes < kam_kmer_workspace_bytes(n_seqs, k)) {
        kam_set_error("kam_kmer_count: workspace too small");
        return KAM_E_WORKSPACE;
    }
    if (n_seqs == 0) return KAM_OK;
    KmerWs w;
    ws_layout(n_seqs, k, d_ws, &w);
    cudaStream_t st = (cudaStream_t)stream;
    const uint2 *pl = (const uint2 *)d_planes;
    if (out_kind == KAM_OUT_COUNTS) {
        if (count_bits == 16)
            return run_kmer<uint16_t>(pl, d_base_start, d_head_mask, n_seqs, k, 16, (uint16_t *)d_out, out_stride, w, total_bases_hint, st);
        return run_kmer<uint32_t>(pl, d_base_start, d_head_mask, n_seqs, k, 32, (uint32_t *)d_out, out_stride, w, total_bases_hint, st);
    }
    if (out_kind == KAM_OUT_FREQ_F32)
        return run_kmer<float>(pl, d_base_start, d_head_mask, n_seqs, k, count_bits, (float *)d_out, out_stride, w, total_bases_hint, st);
    return run_kmer<double>(pl, d_base_start, d_head_mask, n_seqs, k, count_bits, (double *)d_out, out_stride, w, total_bases_hint, st);
}

int kam_kmer_count_sparse(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                          uint32_t n_seqs, int k, const uint64_t *d_entry_start, uint32_t *d_entries, uint32_t *d_nnz,
                          void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(k >= KAM_SPARSE_MIN_K && k <= KAM_SPARSE_MAX_K, "sparse vectors need 7 <= k <= 12");
    KAM_REQUIRE(d_planes && d_base_start && d_head_mask && d_entry_start && d_entries && d_nnz && d_ws, "null pointer");
    if (ws_bytes < kam_kmer_workspace_bytes(n_seqs, k)) {
        kam_set_error("kam_kmer_count_sparse: workspace too small");
        return KAM_E_WORKSPACE;
    }
    if (n_seqs == 0) return KAM_OK;
    KmerWs w;
    ws_layout(n_seqs, k, d_ws, &w);
    cudaStream_t st = (cudaStream_t)stream;
    KAM_CUDA(cudaMemsetAsync(w.queue, 0, 64, st));
    const uint32_t bins = 1u << (2 * k);
    const size_t smem = bins < TILE_BINS ? bins : TILE_BINS;
    auto kern = cgr_sparse_kernel<256>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    KAM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) per_sm = 1;
    uint64_t grid = (uint64_t)kam_sm_count() * per_sm;
    if (grid > n_seqs) grid = n_seqs;
    kern<<<(unsigned)grid, 256, smem, st>>>((const uint2 *)d_planes, d_base_start, d_head_mask, n_seqs, k, d_entry_start,
                                            d_entries, d_nnz, w.queue);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

int kam_kmer_sweep(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                   uint32_t n_seqs, uint64_t total_bases_hint, int k_lo, int k_hi, int count_bits, int out_kind,
                   void *const *h_outs, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(k_lo >= 1 && k_lo <= k_hi, "need 1 <= k_lo <= k_hi");
    if (2 * k_hi > 64) {
        kam_set_error("k is too large to fit in the index");
        return KAM_E_INVALID;
    }
    KAM_REQUIRE(k_hi <= 15, "k > 15 not supported by this build");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind >= KAM_OUT_COUNTS && out_kind <= KAM_OUT_FREQ_F64, "bad out_kind");
    KAM_REQUIRE(d_planes && d_base_start && d_head_mask && h_outs && d_ws, "null pointer");
    for (int k = k_lo; k <= k_hi; ++k) KAM_REQUIRE(h_outs[k - k_lo], "null output pointer");
    if (ws_bytes < kam_kmer_workspace_bytes(n_seqs, k_hi)) {
        kam_set_error("kam_kmer_sweep: workspace too small (needs kam_kmer_workspace_bytes(n_seqs, k_hi))");
        return KAM_E_WORKSPACE;
    }
    if (n_seqs == 0) return KAM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    // orders >= 8 (a vector is >= 256 KB: output-bound) run the ordinary tiled kernel; the orders below
    // come from ONE scan at order min(k_hi, 7) whose table fits a single shared-memory tile, pooled down
    // inside the CTA.  The fused part needs genome-length sequences (pooled uint16 counters, uint8 top).
    const int top = g_sweep_top < 6 ? 6 : (g_sweep_top > SW_MAX_K ? SW_MAX_K : g_sweep_top);
    const int kf = k_hi < top ? k_hi : top;                   // top order of the fused part
    const bool fused = kf >= 6 && k_lo < kf && (total_bases_hint == 0 || total_bases_hint / n_seqs <= 65535);
    for (int k = k_hi; k >= k_lo && (k > kf || !fused); --k) {
        int rc = kam_kmer_count(d_planes, d_base_start, d_head_mask, n_seqs, total_bases_hint, k, count_bits, out_kind,
                                h_outs[k - k_lo], (uint64_t)1 << (2 * k), d_ws, ws_bytes, stream);
        if (rc) return rc;
    }
    if (!fused) return KAM_OK;
    k_hi = kf;
    KmerWs w;
    ws_layout(n_seqs, k_hi, d_ws, &w);
    const uint2 *pl = (const uint2 *)d_planes;
    if (out_kind == KAM_OUT_COUNTS)
        return count_bits == 16 ? run_sweep<uint16_t>(pl, d_base_start, d_head_mask, n_seqs, k_lo, k_hi, 16, h_outs, w, st)
                                : run_sweep<uint32_t>(pl, d_base_start, d_head_mask, n_seqs, k_lo, k_hi, 32, h_outs, w, st);
    if (out_kind == KAM_OUT_FREQ_F32)
        return run_sweep<float>(pl, d_base_start, d_head_mask, n_seqs, k_lo, k_hi, count_bits, h_outs, w, st);
    return run_sweep<double>(pl, d_base_start, d_head_mask, n_seqs, k_lo, k_hi, count_bits, h_outs, w, st);
}

// tuning / test hook: top order of the fused part of kam_kmer_sweep (default 7; 8..10 fuse the tiled
// orders too, which was measured slower: fewer CTAs fit per SM next to the pooling buffers)
void kam_kmer_sweep_set_fused_top(int k) { g_sweep_top = k; }
void kam_kmer_set_warp_path(int on) { g_warp_path = on; }
void kam_kmer_set_word_scan(int on) { g_word_scan = on; }
void kam_kmer_set_long_path(int on) { g_long_path = on; }
void kam_kmer_set_long_one_pass(int on) { g_long_one_pass = on; }

}  // extern "C"
