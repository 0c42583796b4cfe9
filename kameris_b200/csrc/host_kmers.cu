// host_kmers.cu -- kam_kmers_host: the whole `kmers` step on HOST buffers.
//
// Replaces `generation_cgr cgr <dir> <out> <k> <bits>` + the numpy frequency pass as driven by
// kameris/job_steps/backend.py:34-56 (file reading/writing stays with the caller): ASCII records
// in host memory -> count / frequency vectors in host memory.  The batch is cut into chunks of
// sequences; each chunk goes H2D -> pack -> k-mer count -> D2H on one of NSLOT streams with its
// own device buffers, so the device-to-host copy of chunk c overlaps the kernels and the
// host-to-device copy of chunk c+1.
//
// Two ways back to the host:
//   dense   the kernel writes the vectors, the copy engine moves them (PCIe: ~57 GB/s measured).  Pinned
//           caller buffers are copied directly; pageable ones go through internal pinned staging.
//   sparse  (k >= 7, chunks whose vectors are mostly zeros -- 9 kb genomes at k = 9 are 96.6 % zeros):
//           the kernel emits sorted (bin, count) lists (kam_kmer_count_sparse, 4 B per non-zero bin
//           instead of 2-8 B per bin), only those cross PCIe, and a pool of host threads expands each list
//           into the caller's dense row with non-temporal stores, writing every cache line of the row
//           exactly once (zero lines and composed lines alike; ~200 GB/s on 8+ cores of the GPU box against
//           57 GB/s for the copy engine, scratch/hostbw.cu).  The frequency division (backend.py:49) runs
//           on the host in IEEE double precision -- the same operation numpy performs -- once per distinct
//           count of a vector.  A chunk in which the sparse kernel declined a sequence is redone densely.
#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#include "common.cuh"

namespace {

constexpr int NSLOT = 3;
constexpr size_t OUT_CHUNK_BYTES = 128ull << 20;
constexpr size_t CHAR_CHUNK_BYTES = 256ull << 20;
constexpr size_t SPARSE_CHAR_CHUNK = 8ull << 20;      // 4 B of entry space per character: 32 MB of staging per slot
constexpr uint32_t SPARSE_MAX_ROWS = 1024;
constexpr int MAX_DEVICES = 64;

// ---- host thread pool (the expansion's workers; the calling thread takes part) --------------------------
class HostPool {
   public:
    ~HostPool() { stop(); }
    int threads() {
        std::lock_guard<std::mutex> g(cfg_);
        return want();
    }
    void set_threads(int n) {
        std::lock_guard<std::mutex> g(cfg_);
        override_ = n > 0 ? n : 0;
    }
    // fn(i) for every i in [0, n), on the pool's threads and the caller
    void parallel_for(uint32_t n, const std::function<void(uint32_t)> &fn) {
        if (n == 0) return;
        std::lock_guard<std::mutex> g(cfg_);
        const int t = want();
        if (t <= 1 || n == 1) {
            for (uint32_t i = 0; i < n; ++i) fn(i);
            return;
        }
        ensure(t - 1);
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            active_ = (int)workers_.size();
            ++gen_;
        }
        cv_.notify_all();
        for (uint32_t i; (i = next_.fetch_add(1)) < n;) fn(i);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [&] { return active_ == 0; });
        fn_ = nullptr;
    }

   private:
    int want() {
        if (override_) return override_;
        if (const char *e = getenv("KAM_HOST_THREADS")) {
            const int v = atoi(e);
            if (v > 0) return v;
        }
        cpu_set_t set;
        int cores = 0;
        if (sched_getaffinity(0, sizeof set, &set) == 0) cores = CPU_COUNT(&set);
        if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
        if (cores <= 0) cores = 1;
        // one process per GPU (torchrun): the ranks of a box share its cores
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
            const int lw = atoi(e);
            if (lw > 1) cores = std::max(2, cores / lw);
        }
        return std::min(cores, 64);
    }
    void ensure(int n) {
        if ((int)workers_.size() == n) return;
        stop();
        quit_ = false;
        const uint64_t g0 = gen_;            // no generation is published while cfg_ is held by our caller
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, g0] { work(g0); });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
    }
    void work(uint64_t seen) {
        for (;;) {
            const std::function<void(uint32_t)> *fn;
            uint32_t n;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (quit_) return;
                fn = fn_;
                n = n_;
            }
            if (fn)
                for (uint32_t i; (i = next_.fetch_add(1)) < n;) (*fn)(i);
            {
                std::lock_guard<std::mutex> l(m_);
                if (--active_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex cfg_, m_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    const std::function<void(uint32_t)> *fn_ = nullptr;
    std::atomic<uint32_t> next_{0};
    uint32_t n_ = 0;
    int active_ = 0, override_ = 0;
    uint64_t gen_ = 0;
    bool quit_ = false;
};

HostPool &pool() {
    static HostPool *p = new HostPool();   // never destroyed: worker threads must not be joined at exit
    return *p;
}

// ---- expansion of one entry list into one dense row ------------------------------------------------------
// Every 64-byte line of the row is written exactly once with non-temporal stores: runs of lines without an
// entry are streamed as zeros straight from a register; a line that holds entries is composed in a small
// buffer and streamed out from there.  The copy-out of a composed line is DELAYED by a few lines (a ring of
// line buffers): a 16-byte load right behind the 2/4/8-byte stores that composed the line cannot be forwarded
// from the store buffer and waits for them to drain (measured: ~90 cycles per entry, half the DRAM rate on
// uint16 rows, where nearly every line of a k = 9 row of a 9 kb genome holds an entry); a few lines later the
// stores have long retired.  (Composing whole 2 KB chunks in L1 instead cost a second pass over every byte and
// was slower on float64 rows, where three lines out of four are plain zeros.)
constexpr int EXPAND_RING = 4;

template <typename OUT>
inline OUT convert_host(uint32_t c, double sum, int out_kind) {
    if (out_kind == KAM_OUT_FREQ_F64) return (OUT)((double)c / sum);          // numpy true division (backend.py:49)
    if (out_kind == KAM_OUT_FREQ_F32) return (OUT)(float)((double)c / sum);
    return (OUT)c;
}

// The body of the expansion, compiled three times: for SSE2 (any x86-64), AVX2 and AVX-512 (one 64-byte store
// per line), chosen once at run time from the CPU's capabilities.  ZERO_LINES(p, n): n zero lines to p,
// non-temporal; CLEAR_LINE(t): zero a line buffer; COPY_LINE(dst, src): line buffer -> memory, non-temporal.
#define KAM_EXPAND_ROW_BODY                                                                                        \
    unsigned long long s = 0;                                                                                      \
    uint32_t maxc = 0;                                                                                             \
    for (uint32_t e = 0; e < nnz; ++e) {                                                                           \
        const uint32_t c = ent[e] & 0xffu;                                                                         \
        s += c;                                                                                                    \
        maxc = c > maxc ? c : maxc;                                                                                \
    }                                                                                                              \
    OUT lut[256];                                                                                                  \
    for (uint32_t c = 0; c <= maxc; ++c) lut[c] = convert_host<OUT>(c, (double)s, out_kind);                       \
    const OUT zero = lut[0]; /* an empty sequence in a frequencies file is 0/0 = NaN everywhere, like numpy */    \
    const size_t row_bytes = bins * sizeof(OUT);                                                                   \
    const bool plain_zero = out_kind == KAM_OUT_COUNTS || s != 0;                                                  \
    if (((uintptr_t)row & 63) != 0 || row_bytes % 64 != 0 || !plain_zero) { /* odd placement / tiny / NaN rows */ \
        for (size_t i = 0; i < bins; ++i) row[i] = zero;                                                           \
        for (uint32_t e = 0; e < nnz; ++e) row[ent[e] >> 8] = lut[ent[e] & 0xffu];                                 \
        return;                                                                                                    \
    }                                                                                                              \
    constexpr uint32_t EPL = 64 / sizeof(OUT); /* elements per 64-byte line */                                     \
    constexpr uint32_t LSH = EPL == 32 ? 5 : (EPL == 16 ? 4 : 3);                                                  \
    char *p = (char *)row;                                                                                         \
    const size_t nlines = bins / EPL;                                                                              \
    alignas(64) OUT ring[EXPAND_RING][EPL];                                                                        \
    char *pend[EXPAND_RING];                                                                                       \
    for (int i = 0; i < EXPAND_RING; ++i) pend[i] = nullptr;                                                       \
    int cur = 0;                                                                                                   \
    size_t line = 0;                                                                                               \
    uint32_t e = 0;                                                                                                \
    while (e < nnz) {                                                                                              \
        uint32_t bin = ent[e] >> 8;                                                                                \
        const size_t l = bin >> LSH;                                                                               \
        if (l > line) ZERO_LINES(p + line * 64, l - line);                                                         \
        if (pend[cur]) COPY_LINE(pend[cur], (const char *)ring[cur]); /* composed EXPAND_RING lines ago */         \
        OUT *t = ring[cur];                                                                                        \
        CLEAR_LINE((char *)t);                                                                                     \
        do {                                                                                                       \
            t[bin & (EPL - 1)] = lut[ent[e] & 0xffu];                                                              \
            ++e;                                                                                                   \
            if (e == nnz) break;                                                                                   \
            bin = ent[e] >> 8;                                                                                     \
        } while ((bin >> LSH) == l);                                                                               \
        pend[cur] = p + l * 64;                                                                                    \
        cur = (cur + 1) % EXPAND_RING;                                                                             \
        line = l + 1;                                                                                              \
    }                                                                                                              \
    for (int i = 0; i < EXPAND_RING; ++i, cur = (cur + 1) % EXPAND_RING)                                           \
        if (pend[cur]) COPY_LINE(pend[cur], (const char *)ring[cur]);                                              \
    if (nlines > line) ZERO_LINES(p + line * 64, nlines - line);

// ---- SSE2
#define ZERO_LINES(q, n)                                           \
    do {                                                           \
        char *zq = (q);                                            \
        const __m128i z = _mm_setzero_si128();                     \
        for (size_t zi = 0; zi < (size_t)(n); ++zi, zq += 64) {    \
            _mm_stream_si128((__m128i *)zq, z);                    \
            _mm_stream_si128((__m128i *)(zq + 16), z);             \
            _mm_stream_si128((__m128i *)(zq + 32), z);             \
            _mm_stream_si128((__m128i *)(zq + 48), z);             \
        }                                                          \
    } while (0)
#define CLEAR_LINE(t)                                              \
    do {                                                           \
        const __m128i z = _mm_setzero_si128();                     \
        _mm_store_si128((__m128i *)(t), z);                        \
        _mm_store_si128((__m128i *)((t) + 16), z);                 \
        _mm_store_si128((__m128i *)((t) + 32), z);                 \
        _mm_store_si128((__m128i *)((t) + 48), z);                 \
    } while (0)
#define COPY_LINE(dst, src)                                                                \
    do {                                                                                   \
        const __m128i a = _mm_load_si128((const __m128i *)(src));                          \
        const __m128i b = _mm_load_si128((const __m128i *)((src) + 16));                   \
        const __m128i c = _mm_load_si128((const __m128i *)((src) + 32));                   \
        const __m128i d = _mm_load_si128((const __m128i *)((src) + 48));                   \
        _mm_stream_si128((__m128i *)(dst), a);                                             \
        _mm_stream_si128((__m128i *)((dst) + 16), b);                                      \
        _mm_stream_si128((__m128i *)((dst) + 32), c);                                      \
        _mm_stream_si128((__m128i *)((dst) + 48), d);                                      \
    } while (0)
template <typename OUT>
void expand_row_sse2(OUT *row, size_t bins, const uint32_t *ent, uint32_t nnz, int out_kind) {
    KAM_EXPAND_ROW_BODY
}
#undef ZERO_LINES
#undef CLEAR_LINE
#undef COPY_LINE

// ---- AVX2
#define ZERO_LINES(q, n)                                           \
    do {                                                           \
        char *zq = (q);                                            \
        const __m256i z = _mm256_setzero_si256();                  \
        for (size_t zi = 0; zi < (size_t)(n); ++zi, zq += 64) {    \
            _mm256_stream_si256((__m256i *)zq, z);                 \
            _mm256_stream_si256((__m256i *)(zq + 32), z);          \
        }                                                          \
    } while (0)
#define CLEAR_LINE(t)                                              \
    do {                                                           \
        const __m256i z = _mm256_setzero_si256();                  \
        _mm256_store_si256((__m256i *)(t), z);                     \
        _mm256_store_si256((__m256i *)((t) + 32), z);              \
    } while (0)
#define COPY_LINE(dst, src)                                                                \
    do {                                                                                   \
        const __m256i a = _mm256_load_si256((const __m256i *)(src));                       \
        const __m256i b = _mm256_load_si256((const __m256i *)((src) + 32));                \
        _mm256_stream_si256((__m256i *)(dst), a);                                          \
        _mm256_stream_si256((__m256i *)((dst) + 32), b);                                   \
    } while (0)
template <typename OUT>
__attribute__((target("avx2"))) void expand_row_avx2(OUT *row, size_t bins, const uint32_t *ent, uint32_t nnz,
                                                     int out_kind) {
    KAM_EXPAND_ROW_BODY
}
#undef ZERO_LINES
#undef CLEAR_LINE
#undef COPY_LINE

// ---- AVX-512: one instruction per line
#define ZERO_LINES(q, n)                                           \
    do {                                                           \
        char *zq = (q);                                            \
        const __m512i z = _mm512_setzero_si512();                  \
        for (size_t zi = 0; zi < (size_t)(n); ++zi, zq += 64) _mm512_stream_si512((__m512i *)zq, z); \
    } while (0)
#define CLEAR_LINE(t) _mm512_store_si512((__m512i *)(t), _mm512_setzero_si512())
#define COPY_LINE(dst, src) _mm512_stream_si512((__m512i *)(dst), _mm512_load_si512((const __m512i *)(src)))
template <typename OUT>
__attribute__((target("avx512f"))) void expand_row_avx512(OUT *row, size_t bins, const uint32_t *ent, uint32_t nnz,
                                                          int out_kind) {
    KAM_EXPAND_ROW_BODY
}
#undef ZERO_LINES
#undef CLEAR_LINE
#undef COPY_LINE

// ---- count rows on AVX-512 with VPEXPAND: a line of the row is its occupancy mask and its values in order ---
// A row of counts is written in 64-byte lines; the list is sorted, so the values of a line are CONSECUTIVE list
// entries and the line itself is `vpexpandw(mask, values)` (uint16, AVX-512 VBMI2) or `vpexpandd` (uint32):
// pass 1 ORs every entry's bit into its line's mask (a 32 KB block of masks at a time, L1-resident), pass 2 walks
// the lines of the block -- popcount of the mask = how many entries the line consumes, one masked 64-byte load of
// those entries, `& 0xff`, narrow, expand, ONE non-temporal store -- with the same instruction sequence for empty
// and occupied lines (an empty line expands to zeros), so nothing depends on a branch that flips at random.
// ~8 instructions per line + ~5 per entry against ~25 cycles per entry of the scalar composition above, which
// leaves the rate to the memory system (uint16 rows were at 103 GB/s on 16 cores where float64 rows, with a
// quarter of the entries per byte, reach 181).  Only `mode: counts` (values are the counts themselves).
constexpr size_t EXPAND_BLOCK_LINES = 8192;
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi2,popcnt"))) void expand_counts_u16_avx512(
    uint16_t *row, size_t bins, const uint32_t *ent, uint32_t nnz) {
    alignas(64) uint32_t masks[EXPAND_BLOCK_LINES];
    const size_t nlines = bins / 32;
    const __m512i ff = _mm512_set1_epi32(0xff);
    uint32_t e = 0;
    for (size_t l0 = 0; l0 < nlines; l0 += EXPAND_BLOCK_LINES) {
        const size_t nl = nlines - l0 < EXPAND_BLOCK_LINES ? nlines - l0 : EXPAND_BLOCK_LINES;
        memset(masks, 0, nl * sizeof(uint32_t));
        const uint32_t bin0 = (uint32_t)(l0 * 32), bin1 = (uint32_t)((l0 + nl) * 32);
        uint32_t e1 = e;
        for (; e1 < nnz; ++e1) {
            const uint32_t bin = ent[e1] >> 8;
            if (bin >= bin1) break;
            masks[(bin - bin0) >> 5] |= 1u << (bin & 31u);
        }
        char *p = (char *)(row + (size_t)bin0);
        for (size_t l = 0; l < nl; ++l, p += 64) {
            const uint32_t m = masks[l];
            const uint32_t cnt = (uint32_t)__builtin_popcount(m);
            const __mmask16 k0 = (__mmask16)((1u << (cnt < 16u ? cnt : 16u)) - 1u);
            __m512i v = _mm512_and_si512(_mm512_maskz_loadu_epi32(k0, ent + e), ff);
            __m512i vals = _mm512_zextsi256_si512(_mm512_cvtepi32_epi16(v));
            if (__builtin_expect(cnt > 16u, 0)) {
                const __mmask16 k1 = (__mmask16)((1u << (cnt - 16u)) - 1u);
                v = _mm512_and_si512(_mm512_maskz_loadu_epi32(k1, ent + e + 16), ff);
                vals = _mm512_inserti64x4(vals, _mm512_cvtepi32_epi16(v), 1);
            }
            _mm512_stream_si512((__m512i *)p, _mm512_maskz_expand_epi16((__mmask32)m, vals));
            e += cnt;
        }
    }
}

__attribute__((target("avx512f,avx512bw,avx512vl,popcnt"))) void expand_counts_u32_avx512(uint32_t *row, size_t bins,
                                                                                          const uint32_t *ent,
                                                                                          uint32_t nnz) {
    alignas(64) uint16_t masks[EXPAND_BLOCK_LINES];
    const size_t nlines = bins / 16;
    const __m512i ff = _mm512_set1_epi32(0xff);
    uint32_t e = 0;
    for (size_t l0 = 0; l0 < nlines; l0 += EXPAND_BLOCK_LINES) {
        const size_t nl = nlines - l0 < EXPAND_BLOCK_LINES ? nlines - l0 : EXPAND_BLOCK_LINES;
        memset(masks, 0, nl * sizeof(uint16_t));
        const uint32_t bin0 = (uint32_t)(l0 * 16), bin1 = (uint32_t)((l0 + nl) * 16);
        uint32_t e1 = e;
        for (; e1 < nnz; ++e1) {
            const uint32_t bin = ent[e1] >> 8;
            if (bin >= bin1) break;
            masks[(bin - bin0) >> 4] |= (uint16_t)(1u << (bin & 15u));
        }
        char *p = (char *)(row + (size_t)bin0);
        for (size_t l = 0; l < nl; ++l, p += 64) {
            const uint32_t m = masks[l];
            const uint32_t cnt = (uint32_t)__builtin_popcount(m);
            const __mmask16 k0 = (__mmask16)((1u << cnt) - 1u);
            const __m512i v = _mm512_and_si512(_mm512_maskz_loadu_epi32(k0, ent + e), ff);
            _mm512_stream_si512((__m512i *)p, _mm512_maskz_expand_epi32((__mmask16)m, v));
            e += cnt;
        }
    }
}

// float32 frequencies the same way: the quotients count / sum of the (few) distinct counts sit in two registers and
// `vpermt2ps` looks them up, 16 entries at a time.  Rows whose largest count passes 31, and empty sequences (0/0:
// NaN everywhere), take the general body.
__attribute__((target("avx512f,avx512bw,avx512vl,popcnt"))) bool expand_f32_avx512(float *row, size_t bins,
                                                                                   const uint32_t *ent, uint32_t nnz) {
    unsigned long long sum = 0;
    uint32_t maxc = 0;
    for (uint32_t i = 0; i < nnz; ++i) {
        const uint32_t c = ent[i] & 0xffu;
        sum += c;
        maxc = c > maxc ? c : maxc;
    }
    if (sum == 0 || maxc > 31) return false;
    alignas(64) float lut[32];
    for (uint32_t c = 0; c < 32; ++c) lut[c] = convert_host<float>(c <= maxc ? c : 0, (double)sum, KAM_OUT_FREQ_F32);
    const __m512 lut_lo = _mm512_load_ps(lut), lut_hi = _mm512_load_ps(lut + 16);
    alignas(64) uint16_t masks[EXPAND_BLOCK_LINES];
    const size_t nlines = bins / 16;
    const __m512i ff = _mm512_set1_epi32(0xff);
    uint32_t e = 0;
    for (size_t l0 = 0; l0 < nlines; l0 += EXPAND_BLOCK_LINES) {
        const size_t nl = nlines - l0 < EXPAND_BLOCK_LINES ? nlines - l0 : EXPAND_BLOCK_LINES;
        memset(masks, 0, nl * sizeof(uint16_t));
        const uint32_t bin0 = (uint32_t)(l0 * 16), bin1 = (uint32_t)((l0 + nl) * 16);
        for (uint32_t e1 = e; e1 < nnz; ++e1) {
            const uint32_t bin = ent[e1] >> 8;
            if (bin >= bin1) break;
            masks[(bin - bin0) >> 4] |= (uint16_t)(1u << (bin & 15u));
        }
        char *p = (char *)(row + (size_t)bin0);
        for (size_t l = 0; l < nl; ++l, p += 64) {
            const uint32_t m = masks[l];
            const uint32_t cnt = (uint32_t)__builtin_popcount(m);
            const __mmask16 k0 = (__mmask16)((1u << cnt) - 1u);
            const __m512i idx = _mm512_and_si512(_mm512_maskz_loadu_epi32(k0, ent + e), ff);
            const __m512 vals = _mm512_permutex2var_ps(lut_lo, idx, lut_hi);
            _mm512_stream_ps((float *)p, _mm512_maskz_expand_ps((__mmask16)m, vals));
            e += cnt;
        }
    }
    return true;
}

// float64 frequencies: 8 elements per line, 8-bit masks, the quotients of counts up to 15 in two registers
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,popcnt"))) bool expand_f64_avx512(double *row, size_t bins,
                                                                                            const uint32_t *ent,
                                                                                            uint32_t nnz) {
    unsigned long long sum = 0;
    uint32_t maxc = 0;
    for (uint32_t i = 0; i < nnz; ++i) {
        const uint32_t c = ent[i] & 0xffu;
        sum += c;
        maxc = c > maxc ? c : maxc;
    }
    if (sum == 0 || maxc > 15) return false;
    alignas(64) double lut[16];
    for (uint32_t c = 0; c < 16; ++c) lut[c] = convert_host<double>(c <= maxc ? c : 0, (double)sum, KAM_OUT_FREQ_F64);
    const __m512d lut_lo = _mm512_load_pd(lut), lut_hi = _mm512_load_pd(lut + 8);
    alignas(64) uint8_t masks[EXPAND_BLOCK_LINES];
    const size_t nlines = bins / 8;
    const __m256i ff = _mm256_set1_epi32(0xff);
    uint32_t e = 0;
    for (size_t l0 = 0; l0 < nlines; l0 += EXPAND_BLOCK_LINES) {
        const size_t nl = nlines - l0 < EXPAND_BLOCK_LINES ? nlines - l0 : EXPAND_BLOCK_LINES;
        memset(masks, 0, nl);
        const uint32_t bin0 = (uint32_t)(l0 * 8), bin1 = (uint32_t)((l0 + nl) * 8);
        for (uint32_t e1 = e; e1 < nnz; ++e1) {
            const uint32_t bin = ent[e1] >> 8;
            if (bin >= bin1) break;
            masks[(bin - bin0) >> 3] |= (uint8_t)(1u << (bin & 7u));
        }
        char *p = (char *)(row + (size_t)bin0);
        for (size_t l = 0; l < nl; ++l, p += 64) {
            const uint32_t m = masks[l];
            const uint32_t cnt = (uint32_t)__builtin_popcount(m);
            const __mmask8 k0 = (__mmask8)((1u << cnt) - 1u);
            const __m512i idx = _mm512_cvtepu32_epi64(_mm256_and_si256(_mm256_maskz_loadu_epi32(k0, ent + e), ff));
            const __m512d vals = _mm512_permutex2var_pd(lut_lo, idx, lut_hi);
            _mm512_stream_pd((double *)p, _mm512_maskz_expand_pd((__mmask8)m, vals));
            e += cnt;
        }
    }
    return true;
}

// 1 when the VPEXPAND path may be used for count rows (KAM_HOST_ISA below avx512 switches it off, like the bodies above)
int expand_counts_isa(bool need_vbmi2);

int g_expand_isa = -1;   // 0 SSE2, 1 AVX2, 2 AVX-512; KAM_HOST_ISA=sse2|avx2|avx512 in the environment overrides
int expand_isa() {
    if (g_expand_isa < 0) {
        int isa = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
        if (const char *e = getenv("KAM_HOST_ISA")) {
            const int want = !strcmp(e, "sse2") ? 0 : (!strcmp(e, "avx2") ? 1 : (!strcmp(e, "avx512") ? 2 : isa));
            if (want < isa) isa = want;
        }
        g_expand_isa = isa;
    }
    return g_expand_isa;
}

template <typename OUT>
void expand_row(OUT *row, size_t bins, const uint32_t *ent, uint32_t nnz, int out_kind) {
    switch (expand_isa()) {
        case 2: expand_row_avx512<OUT>(row, bins, ent, nnz, out_kind); break;
        case 1: expand_row_avx2<OUT>(row, bins, ent, nnz, out_kind); break;
        default: expand_row_sse2<OUT>(row, bins, ent, nnz, out_kind);
    }
}

int expand_counts_isa(bool need_vbmi2) {
    static const int have = [] {
        const bool base = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                          __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("popcnt");
        int v = base ? 1 : 0;
        if (base && __builtin_cpu_supports("avx512vbmi2")) v = 2;
        if (const char *e = getenv("KAM_HOST_EXPAND")) {   // "scalar": the line-composition bodies only (tests)
            if (!strcmp(e, "scalar")) v = 0;
        }
        return v;
    }();
    return expand_isa() == 2 && have >= (need_vbmi2 ? 2 : 1);
}

void expand_dispatch(void *row, size_t bins, const uint32_t *ent, uint32_t nnz, int count_bits, int out_kind) {
    if (out_kind == KAM_OUT_COUNTS && ((uintptr_t)row & 63) == 0 && bins % 64 == 0) {
        if (count_bits == 16 && expand_counts_isa(true)) return expand_counts_u16_avx512((uint16_t *)row, bins, ent, nnz);
        if (count_bits == 32 && expand_counts_isa(false)) return expand_counts_u32_avx512((uint32_t *)row, bins, ent, nnz);
    }
    if (out_kind == KAM_OUT_FREQ_F32 && ((uintptr_t)row & 63) == 0 && bins % 64 == 0 && expand_counts_isa(false) &&
        expand_f32_avx512((float *)row, bins, ent, nnz))
        return;
    if (out_kind == KAM_OUT_FREQ_F64 && ((uintptr_t)row & 63) == 0 && bins % 64 == 0 && expand_counts_isa(false) &&
        __builtin_cpu_supports("avx512dq") && expand_f64_avx512((double *)row, bins, ent, nnz))
        return;
    if (out_kind == KAM_OUT_FREQ_F64) expand_row<double>((double *)row, bins, ent, nnz, out_kind);
    else if (out_kind == KAM_OUT_FREQ_F32) expand_row<float>((float *)row, bins, ent, nnz, out_kind);
    else if (count_bits == 16) expand_row<uint16_t>((uint16_t *)row, bins, ent, nnz, out_kind);
    else expand_row<uint32_t>((uint32_t *)row, bins, ent, nnz, out_kind);
}

// ---- device / pinned buffers, pipeline slots --------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KAM_OK;
        release();
        size_t want = kam_align_up(bytes + bytes / 8, 1 << 20);
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            kam_set_error("cudaMalloc(%zu) failed", want);
            return KAM_E_NOMEM;
        }
        cap = want;
        return KAM_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KAM_OK;
        release();
        size_t want = kam_align_up(bytes + bytes / 8, 1 << 20);
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            kam_set_error("cudaMallocHost(%zu) failed", want);
            return KAM_E_NOMEM;
        }
        cap = want;
        return KAM_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;   // D2H of the chunk last issued on this slot
    DevBuf chars, starts, planes, base, head, pack_ws, kmer_ws, out, entries, nnz;
    PinBuf h_starts, h_chars, h_out, h_entries, h_nnz;
    bool busy = false;            // a chunk is in flight
    // what the in-flight chunk still owes the caller once `done` has fired
    bool sparse = false;
    uint32_t first = 0, n = 0;
    uint64_t nchars = 0;
    void *pend_dst = nullptr;     // dense, pageable caller buffer: copy out of h_out
    size_t pend_bytes = 0;
};

struct HostCtx {
    bool ready = false;
    Slot slot[NSLOT];
};

std::mutex g_call;                 // kam_kmers_host calls are serialised (one pipeline per device)
HostCtx g_ctx[MAX_DEVICES];
int g_sparse_mode = 1;             // 0 = never, 1 = when the vectors are sparse enough (default), 2 = whenever k allows

int ctx_get(HostCtx **out) {
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) {
        kam_set_error("device index %d out of range", dev);
        return KAM_E_INVALID;
    }
    HostCtx &c = g_ctx[dev];
    if (!c.ready) {
        for (auto &s : c.slot) {
            if (!s.st) KAM_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            if (!s.done) KAM_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        c.ready = true;
    }
    *out = &c;
    return KAM_OK;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

struct Job {
    const uint8_t *h_chars;
    const uint64_t *h_char_start;
    int k, count_bits, out_kind;
    void *h_out;            // may be null: the caller only wants the resident counts
    void *d_keep;           // may be null; else the dense count rows also stay on the device
    uint64_t keep_stride;   // elements between rows of d_keep
    size_t bins, row_bytes;
    bool out_pinned, in_pinned;
};

// the chunk's dense count vectors into the caller's device matrix (the `distances` step that follows in a job
// reads them there instead of re-reading the .mm-repr and uploading it again)
int issue_keep(Slot &s, const Job &j) {
    if (!j.d_keep) return KAM_OK;
    int rc = s.kmer_ws.ensure(kam_kmer_workspace_bytes(s.n, j.k));
    if (rc) return rc;
    char *dst = (char *)j.d_keep + (size_t)s.first * j.keep_stride * (j.count_bits / 8);
    return kam_kmer_count((const uint64_t *)s.planes.p, (const uint64_t *)s.base.p, (const uint32_t *)s.head.p, s.n,
                          s.nchars, j.k, j.count_bits, KAM_OUT_COUNTS, dst, j.keep_stride, s.kmer_ws.p, s.kmer_ws.cap, s.st);
}

// the chunk's vectors, densely, straight into the caller's rows (also the redo of a declined sparse chunk)
int issue_dense(Slot &s, const Job &j) {
    int rc;
    s.sparse = false;
    if (!j.h_out) {
        KAM_CUDA(cudaEventRecord(s.done, s.st));
        return KAM_OK;
    }
    if ((rc = s.kmer_ws.ensure(kam_kmer_workspace_bytes(s.n, j.k))) || (rc = s.out.ensure((size_t)s.n * j.row_bytes)))
        return rc;
    rc = kam_kmer_count((const uint64_t *)s.planes.p, (const uint64_t *)s.base.p, (const uint32_t *)s.head.p, s.n,
                        s.nchars, j.k, j.count_bits, j.out_kind, s.out.p, j.bins, s.kmer_ws.p, s.kmer_ws.cap, s.st);
    if (rc) return rc;
    char *dst = (char *)j.h_out + (size_t)s.first * j.row_bytes;
    if (j.out_pinned) {
        KAM_CUDA(cudaMemcpyAsync(dst, s.out.p, (size_t)s.n * j.row_bytes, cudaMemcpyDeviceToHost, s.st));
    } else {
        if ((rc = s.h_out.ensure((size_t)s.n * j.row_bytes))) return rc;
        KAM_CUDA(cudaMemcpyAsync(s.h_out.p, s.out.p, (size_t)s.n * j.row_bytes, cudaMemcpyDeviceToHost, s.st));
        s.pend_dst = dst;
        s.pend_bytes = (size_t)s.n * j.row_bytes;
    }
    s.sparse = false;
    KAM_CUDA(cudaEventRecord(s.done, s.st));
    return KAM_OK;
}

int slot_drain(Slot &s, const Job &j) {
    if (!s.busy) return KAM_OK;
    KAM_CUDA(cudaEventSynchronize(s.done));
    if (s.sparse) {
        const uint32_t *nnz = (const uint32_t *)s.h_nnz.p;
        bool declined = false;
        for (uint32_t i = 0; i < s.n; ++i) declined |= nnz[i] == KAM_SPARSE_DECLINED;
        if (declined) {                       // rare: a sequence too long for the tiles or a count above 255
            int rc = issue_dense(s, j);
            if (rc) return rc;
            KAM_CUDA(cudaEventSynchronize(s.done));
        } else {
            const uint32_t *ent = (const uint32_t *)s.h_entries.p;
            const uint64_t *hs = (const uint64_t *)s.h_starts.p;
            char *dst = (char *)j.h_out + (size_t)s.first * j.row_bytes;
            pool().parallel_for(s.n, [&](uint32_t i) {
                expand_dispatch(dst + (size_t)i * j.row_bytes, j.bins, ent + hs[i], nnz[i], j.count_bits, j.out_kind);
                _mm_sfence();
            });
        }
    }
    if (s.pend_dst) {
        memcpy(s.pend_dst, s.h_out.p, s.pend_bytes);
        s.pend_dst = nullptr;
    }
    s.busy = false;
    return KAM_OK;
}

// after an error: nothing of this call may still be writing into the caller's buffers when it returns
void abandon(HostCtx &c) {
    for (auto &s : c.slot) {
        if (s.st) cudaStreamSynchronize(s.st);
        s.busy = false;
        s.pend_dst = nullptr;
    }
    cudaGetLastError();
}

int kmers_host_locked(HostCtx &ctx, const Job &j, uint32_t n_seqs) {
    int rc;
    const bool sparse_k = j.k >= KAM_SPARSE_MIN_K && j.k <= KAM_SPARSE_MAX_K && g_sparse_mode != 0;
    const uint32_t dense_rows = (uint32_t)std::max<size_t>(1, OUT_CHUNK_BYTES / j.row_bytes);
    uint32_t first = 0;
    int turn = 0;
    while (first < n_seqs) {
        // a run of sequences is expanded on the host when its entry lists (4 B per character at most) are at
        // most a quarter of the dense bytes; decided per chunk from the first sequences' lengths
        const uint64_t c0 = j.h_char_start[first];
        bool sparse = false;
        if (sparse_k && j.h_out) {
            const uint32_t probe = std::min<uint32_t>(n_seqs - first, SPARSE_MAX_ROWS);
            const uint64_t pc = j.h_char_start[first + probe] - c0;
            sparse = g_sparse_mode == 2 || pc * 4 * 4 <= (uint64_t)probe * j.row_bytes;
        }
        const uint32_t max_rows = sparse ? SPARSE_MAX_ROWS : dense_rows;
        const uint64_t max_chars = sparse ? SPARSE_CHAR_CHUNK : CHAR_CHUNK_BYTES;
        uint32_t last = first;
        while (last < n_seqs && last - first < max_rows &&
               (last == first || j.h_char_start[last + 1] - c0 <= max_chars))
            ++last;
        const uint32_t n = last - first;
        const uint64_t nchars = j.h_char_start[last] - c0;

        Slot &s = ctx.slot[turn];
        if ((rc = slot_drain(s, j))) return rc;
        s.first = first;
        s.n = n;
        s.nchars = nchars;
        s.sparse = sparse;

        if ((rc = s.chars.ensure(nchars + 32)) || (rc = s.starts.ensure((size_t)(n + 1) * 8)) ||
            (rc = s.planes.ensure(kam_planes_words(nchars) * 8)) || (rc = s.base.ensure((size_t)(n + 1) * 8)) ||
            (rc = s.head.ensure((size_t)n * 4)) || (rc = s.pack_ws.ensure(kam_pack_workspace_bytes(nchars, n))) ||
            (rc = s.h_starts.ensure((size_t)(n + 1) * 8)))
            return rc;

        uint64_t *hs = (uint64_t *)s.h_starts.p;
        for (uint32_t i = 0; i <= n; ++i) hs[i] = j.h_char_start[first + i] - c0;
        KAM_CUDA(cudaMemcpyAsync(s.starts.p, hs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s.st));
        s.busy = true;
        if (nchars) {
            const void *src = j.h_chars + c0;
            if (!j.in_pinned) {
                if ((rc = s.h_chars.ensure(nchars))) return rc;
                memcpy(s.h_chars.p, src, nchars);
                src = s.h_chars.p;
            }
            KAM_CUDA(cudaMemcpyAsync(s.chars.p, src, nchars, cudaMemcpyHostToDevice, s.st));
        }
        rc = kam_pack_ascii((const uint8_t *)s.chars.p, (const uint64_t *)s.starts.p, n, nchars, (uint64_t *)s.planes.p,
                            (uint64_t *)s.base.p, (uint32_t *)s.head.p, s.pack_ws.p, s.pack_ws.cap, s.st);
        if (rc) return rc;
        if ((rc = issue_keep(s, j))) return rc;
        if (sparse) {
            const size_t eb = (size_t)(nchars ? nchars : 1) * 4;
            if ((rc = s.kmer_ws.ensure(kam_kmer_workspace_bytes(n, j.k))) || (rc = s.entries.ensure(eb)) ||
                (rc = s.nnz.ensure((size_t)n * 4)) || (rc = s.h_entries.ensure(eb)) || (rc = s.h_nnz.ensure((size_t)n * 4)))
                return rc;
            rc = kam_kmer_count_sparse((const uint64_t *)s.planes.p, (const uint64_t *)s.base.p, (const uint32_t *)s.head.p,
                                       n, j.k, (const uint64_t *)s.starts.p, (uint32_t *)s.entries.p, (uint32_t *)s.nnz.p,
                                       s.kmer_ws.p, s.kmer_ws.cap, s.st);
            if (rc) return rc;
            KAM_CUDA(cudaMemcpyAsync(s.h_nnz.p, s.nnz.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s.st));
            if (nchars)
                KAM_CUDA(cudaMemcpyAsync(s.h_entries.p, s.entries.p, (size_t)nchars * 4, cudaMemcpyDeviceToHost, s.st));
            KAM_CUDA(cudaEventRecord(s.done, s.st));
        } else if ((rc = issue_dense(s, j))) {
            return rc;
        }
        first = last;
        turn = (turn + 1) % NSLOT;
    }
    for (int i = 0; i < NSLOT; ++i) {   // oldest first
        if ((rc = slot_drain(ctx.slot[turn], j))) return rc;
        turn = (turn + 1) % NSLOT;
    }
    return KAM_OK;
}

}  // namespace

extern "C" {

int kam_kmers_host_keep(const uint8_t *h_chars, const uint64_t *h_char_start, uint32_t n_seqs, int k, int count_bits,
                        int out_kind, void *h_out, void *d_counts, uint64_t counts_stride) {
    if (k < 1 || 2 * k > 64) {
        kam_set_error("k is too large to fit in the index");
        return KAM_E_INVALID;
    }
    KAM_REQUIRE(k <= 15, "k > 15 not supported by this build");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind >= KAM_OUT_COUNTS && out_kind <= KAM_OUT_FREQ_F64, "bad out_kind");
    KAM_REQUIRE(h_char_start && (h_out || d_counts), "null pointer");
    KAM_REQUIRE(!d_counts || counts_stride >= ((uint64_t)1 << (2 * k)), "counts_stride smaller than 4^k");
    if (n_seqs == 0) return KAM_OK;
    KAM_REQUIRE(h_chars || h_char_start[n_seqs] == h_char_start[0], "null character buffer");
    std::lock_guard<std::mutex> lock(g_call);
    HostCtx *ctx = nullptr;
    int rc = ctx_get(&ctx);
    if (rc) return rc;
    Job j;
    j.h_chars = h_chars;
    j.h_char_start = h_char_start;
    j.k = k;
    j.count_bits = count_bits;
    j.out_kind = out_kind;
    j.h_out = h_out;
    j.d_keep = d_counts;
    j.keep_stride = counts_stride;
    j.bins = (size_t)1 << (2 * k);
    const size_t esz = out_kind == KAM_OUT_COUNTS ? (size_t)count_bits / 8 : (out_kind == KAM_OUT_FREQ_F32 ? 4 : 8);
    j.row_bytes = j.bins * esz;
    j.out_pinned = h_out && is_pinned(h_out);
    j.in_pinned = h_chars && is_pinned(h_chars);
    rc = kmers_host_locked(*ctx, j, n_seqs);
    if (rc) abandon(*ctx);
    else if (d_counts)
        for (auto &s : ctx->slot)            // the kept rows are complete when the call returns
            if (s.st) cudaStreamSynchronize(s.st);
    return rc;
}

int kam_kmers_host(const uint8_t *h_chars, const uint64_t *h_char_start, uint32_t n_seqs, int k, int count_bits,
                   int out_kind, void *h_out) {
    KAM_REQUIRE(h_out, "null pointer");
    return kam_kmers_host_keep(h_chars, h_char_start, n_seqs, k, count_bits, out_kind, h_out, nullptr, 0);
}

void kam_kmers_host_set_sparse(int mode) { g_sparse_mode = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }
void kam_host_set_threads(int threads) { pool().set_threads(threads); }
int kam_host_threads(void) { return pool().threads(); }

int kam_sparse_expand(const uint32_t *h_entries, const uint64_t *h_entry_start, const uint32_t *h_nnz, uint32_t n_seqs,
                      int k, int count_bits, int out_kind, void *h_out) {
    KAM_REQUIRE(k >= 1 && k <= KAM_SPARSE_MAX_K, "need 1 <= k <= 12");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind >= KAM_OUT_COUNTS && out_kind <= KAM_OUT_FREQ_F64, "bad out_kind");
    KAM_REQUIRE(h_entry_start && h_nnz && h_out, "null pointer");
    if (n_seqs == 0) return KAM_OK;
    const size_t bins = (size_t)1 << (2 * k);
    const size_t esz = out_kind == KAM_OUT_COUNTS ? (size_t)count_bits / 8 : (out_kind == KAM_OUT_FREQ_F32 ? 4 : 8);
    // every list is checked by the thread that expands it (ascending bins inside the table, non-zero counts);
    // a bad list leaves its row untouched and fails the call
    std::atomic<int> bad{0};
    pool().parallel_for(n_seqs, [&](uint32_t i) {
        const uint32_t nz = h_nnz[i];
        const uint32_t *ent = h_entries ? h_entries + h_entry_start[i] : nullptr;
        int why = 0;
        if (nz == KAM_SPARSE_DECLINED) why = 1;
        else if (nz && !ent) why = 2;
        else
            for (uint32_t e = 0; e < nz && !why; ++e) {
                if ((ent[e] >> 8) >= bins || (ent[e] & 0xffu) == 0) why = 3;
                else if (e && (ent[e - 1] >> 8) >= (ent[e] >> 8)) why = 4;
            }
        if (why) {
            int expect = 0;
            bad.compare_exchange_strong(expect, why);
            return;
        }
        expand_dispatch((char *)h_out + (size_t)i * bins * esz, bins, ent, nz, count_bits, out_kind);
        _mm_sfence();
    });
    static const char *msg[] = {"", "a declined vector has no entry list", "null entry buffer", "entry out of range",
                                "entries must ascend by bin"};
    KAM_REQUIRE(bad.load() == 0, msg[bad.load()]);
    return KAM_OK;
}

}  // extern "C"
