/*
 * kameris_oracle.c -- CPU restatement of the kameris `kmers` and `distances` hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in kameris_b200/ (the product) may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker or as the timed CPU baseline.
 *
 * PARITY STATUS: the arithmetic of the reference lives in the prebuilt executables
 * kameris/scripts/generation_cgr_* and generation_dists_* (source: stephensolis/kameris-backend,
 * not vendored, not version-pinned; Linux builds absent from the mount, see
 * /root/reference/.MISSING_LARGE_BLOBS:4-7).  This file restates the semantics recovered from
 * the Mach-O/PE builds.  The reference holds no golden vectors for this
 * path.  How each part is pinned:
 *   CGR counting (orc_cgr_count_*, orc_fasta_record0 feeds it): PINNED against
 *       tests/golden/refcode_cgr.json.gz -- 211 outputs of the reference's OWN machine code (the
 *       leaf CGR-fill function of generation_cgr_windows_avx2.exe, VA 0x140016040, executed
 *       in-process by oracle/refcode/refcode.c; generator oracle/refcode/gen_golden.py committed).
 *       The same values reproduce SURVEY.md Appendix B (demo genomes, quirk cases Q1/Q3/Q4).
 *   frequencies: numpy true division as written in backend.py:49 (executed by numpy itself in
 *       oracle.py; the C twin here is checked against it).
 *   manhat / info (orc_manhat_*, orc_info_*): PINNED against tests/golden/refcode_dists.json.gz --
 *       48 cases produced by the reference's OWN machine code: the six per-element-type row lambdas
 *       of generation_dists_windows_avx2.exe (VA 0x140024610 ... 0x140027d50) mapped at their image
 *       base and called in-process by oracle/refcode/refdists.c (generator gen_golden_dists.py).
 *       This corrected the restatement once: for T=float the truncating sum is formed in single
 *       precision (SURVEY.md A.6 describes the double instance only).
 *   euclid / cosine / pearson: not in the reference at all; defined as scipy cdist (oracle.py).
 *
 * Each function cites the reference location it follows:
 *   "cgr@0x..."   = kameris/scripts/generation_cgr_mac_avx2   virtual address
 *   "dists@0x..." = kameris/scripts/generation_dists_mac_avx2 virtual address
 *   "backend.py:N" = kameris/job_steps/backend.py line N
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * A.1  FASTA record extraction.   cgr@0x10000d691-0x10000d83e, cgr@0x10000e0bb-0x10000e145
 *   getline loop; each line boost::trim'med; an empty line or one starting with '>' ends the
 *   current record (kept if non-empty); other lines are appended; after EOF a non-empty
 *   current record is kept.  Only record[0] is used (quirk Q2).
 * Writes record 0 into `out` (capacity >= n) and returns its length, or -1 when the file holds
 * no record (undefined behaviour in the reference; callers must raise).
 * ---------------------------------------------------------------------------------------- */
static int orc_isspace(unsigned char c) {
    /* std::isspace in the classic locale, which boost::algorithm::trim uses */
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

int64_t orc_fasta_record0(const char *buf, size_t n, char *out) {
    size_t pos = 0, len = 0;
    while (pos < n) {
        size_t e = pos;
        while (e < n && buf[e] != '\n') e++;
        size_t a = pos, b = e;
        while (a < b && orc_isspace((unsigned char)buf[a])) a++;
        while (b > a && orc_isspace((unsigned char)buf[b - 1])) b--;
        if (a == b || buf[a] == '>') {
            if (len > 0) return (int64_t)len; /* record[0] complete */
        } else {
            memcpy(out + len, buf + a, b - a);
            len += b - a;
        }
        pos = e + 1;
    }
    return len > 0 ? (int64_t)len : -1;
}

/* ------------------------------------------------------------------------------------------
 * A.2  CGR counting.   cgr@0x10000e136-0x10000e304 (loop), cgr@0x10000ec41-0x10000ec66 (inc)
 *   x = y = 1<<(k-1); per character (raw index i counts ALL characters):
 *     skip unless c in {A,C,G,T} (case-sensitive, Q4; window is NOT reset)
 *     x >>= 1; y >>= 1; G,T set the top bit of x; A,T set the top bit of y
 *     if (i >= k-1) count[(y<<k)|x]++      (raw-index test: Q1; T arithmetic wraps: Q3)
 * Returns 0, or -1 when 2k > 64 ("k is too large to fit in the index", cgr@0x10000e15f).
 * ---------------------------------------------------------------------------------------- */
#define ORC_CGR_BODY(T)                                                              \
    if (k < 1 || 2 * k > 64) return -1;                                              \
    const uint64_t top = 1ull << (k - 1);                                            \
    uint64_t x = top, y = top;                                                       \
    for (size_t i = 0; i < n; i++) {                                                 \
        const char c = rec[i];                                                       \
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') continue;                  \
        x >>= 1;                                                                     \
        y >>= 1;                                                                     \
        if (c == 'G' || c == 'T') x |= top;                                          \
        if (c == 'A' || c == 'T') y |= top;                                          \
        if (i >= (size_t)(k - 1)) count[(y << k) | x] = (T)(count[(y << k) | x] + 1); \
    }                                                                                \
    return 0;

int orc_cgr_count_u16(const char *rec, size_t n, int k, uint16_t *count) { ORC_CGR_BODY(uint16_t) }
int orc_cgr_count_u32(const char *rec, size_t n, int k, uint32_t *count) { ORC_CGR_BODY(uint32_t) }

/* ------------------------------------------------------------------------------------------
 * backend.py:42-56  frequency normalisation: cgr / cgr.sum() with numpy true division.
 * numpy sums uint16/uint32 arrays in uint64 and divides in float64.
 * ---------------------------------------------------------------------------------------- */
void orc_freq_u16(const uint16_t *count, size_t d, double *out) {
    uint64_t s = 0;
    for (size_t i = 0; i < d; i++) s += count[i];
    const double ds = (double)s;
    for (size_t i = 0; i < d; i++) out[i] = (double)count[i] / ds;
}
void orc_freq_u32(const uint32_t *count, size_t d, double *out) {
    uint64_t s = 0;
    for (size_t i = 0; i < d; i++) s += count[i];
    const double ds = (double)s;
    for (size_t i = 0; i < d; i++) out[i] = (double)count[i] / ds;
}

/* ------------------------------------------------------------------------------------------
 * A.6  manhat.   dists@0x10001be4e-0x10001bf62 (integer T), dists@0x10001dcd1-0x10001dd71 (fp T)
 *   uint32 acc = 0; for d: acc += (T)|a_d - b_d|   -- difference taken in 32 bits for
 *   8/16/32-bit T (so u32 differences wrap mod 2^32 before the conditional negate), truncated
 *   to T, added mod 2^32.  For floating T each step is acc = (uint32)(int64)((T)acc + |a-b|) evaluated
 *   in T's own precision (float: vaddss, double: vaddsd) with a truncating conversion (quirk Q5).
 * ---------------------------------------------------------------------------------------- */
uint32_t orc_manhat_u8(const uint8_t *a, const uint8_t *b, size_t d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) {
        int32_t t = (int32_t)a[i] - (int32_t)b[i];
        acc += (uint8_t)(t < 0 ? -t : t);
    }
    return acc;
}
uint32_t orc_manhat_u16(const uint16_t *a, const uint16_t *b, size_t d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) {
        int32_t t = (int32_t)a[i] - (int32_t)b[i];
        acc += (uint16_t)(t < 0 ? -t : t);
    }
    return acc;
}
uint32_t orc_manhat_u32(const uint32_t *a, const uint32_t *b, size_t d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) acc += a[i] > b[i] ? a[i] - b[i] : b[i] - a[i];
    return acc;
}
uint32_t orc_manhat_u64(const uint64_t *a, const uint64_t *b, size_t d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) acc += (uint32_t)(a[i] > b[i] ? a[i] - b[i] : b[i] - a[i]);
    return acc;
}
uint32_t orc_manhat_f32(const float *a, const float *b, size_t d) {
    /* T=float instance (generation_dists_windows_avx2.exe VA 0x140027a20-0x140027ad3): the sum is
     * formed in SINGLE precision -- vcvtsi2ss (float)acc, vsubss, conditional sign flip, vaddss,
     * vcvttss2si to int64, low 32 bits kept. */
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) {
        volatile float t = a[i] - b[i];
        volatile float m = (b[i] > a[i]) ? -t : t;
        volatile float s = (float)acc + m;
        acc = (uint32_t)(int64_t)s;
    }
    return acc;
}
uint32_t orc_manhat_f64(const double *a, const double *b, size_t d) {
    uint32_t acc = 0;
    for (size_t i = 0; i < d; i++) {
        double t = a[i] - b[i];
        acc = (uint32_t)(int64_t)((double)acc + (t < 0 ? -t : t));
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * A.6  info.   dists@0x10001bab5-0x10001bcb0
 *   U = #{d: a_d>0 or b_d>0}, I = #{d: a_d>0 and b_d>0} (u64 counters);
 *   result = U==0 ? 0.0f : (1.0f/(float)U) * (float)(U-I)     single precision, recip-multiply.
 * ---------------------------------------------------------------------------------------- */
static float orc_info_finish(uint64_t U, uint64_t I) {
    if (U == 0) return 0.0f;
    volatile float r = 1.0f / (float)U; /* volatile: keep the reciprocal a rounded float */
    return r * (float)(U - I);
}
#define ORC_INFO_BODY                                     \
    uint64_t U = 0, I = 0;                                \
    for (size_t i = 0; i < d; i++) {                      \
        const int pa = a[i] > 0, pb = b[i] > 0;           \
        U += (uint64_t)(pa | pb);                         \
        I += (uint64_t)(pa & pb);                         \
    }                                                     \
    return orc_info_finish(U, I);
float orc_info_u8(const uint8_t *a, const uint8_t *b, size_t d) { ORC_INFO_BODY }
float orc_info_u16(const uint16_t *a, const uint16_t *b, size_t d) { ORC_INFO_BODY }
float orc_info_u32(const uint32_t *a, const uint32_t *b, size_t d) { ORC_INFO_BODY }
float orc_info_u64(const uint64_t *a, const uint64_t *b, size_t d) { ORC_INFO_BODY }
float orc_info_f32(const float *a, const float *b, size_t d) { ORC_INFO_BODY }
float orc_info_f64(const double *a, const double *b, size_t d) { ORC_INFO_BODY }

/* ------------------------------------------------------------------------------------------
 * A.7  mmg::ParallelExecutor.   cgr@0x10001a6f0 (execute), cgr@0x100021d84-0x100021de8 (worker)
 *   n threads; each: loop { i = atomic_fetch_add(&next, 1); if (i >= n_tasks) break; task(i) }.
 * Used by the batch drivers below, which are what bench.py times as the CPU baseline.
 * ---------------------------------------------------------------------------------------- */
typedef void (*orc_task_fn)(void *ctx, size_t i);
typedef struct {
    orc_task_fn fn;
    void *ctx;
    size_t n_tasks;
    atomic_size_t next;
} orc_pool;

static void *orc_worker(void *p) {
    orc_pool *pool = (orc_pool *)p;
    for (;;) {
        size_t i = atomic_fetch_add(&pool->next, 1);
        if (i >= pool->n_tasks) break;
        pool->fn(pool->ctx, i);
    }
    return NULL;
}

static void orc_execute(orc_task_fn fn, void *ctx, size_t n_tasks, int n_threads) {
    orc_pool pool = {fn, ctx, n_tasks, 0};
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, orc_worker, &pool);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
}

/* run<T>: one task per sequence (cgr@0x100005df0 / cgr@0x10000b3d0); results N x 4^k in RAM. */
typedef struct {
    const char *chars;
    const uint64_t *start; /* n+1 offsets into chars */
    int k, bits;
    void *out; /* n x 4^k of u16/u32, zero-filled here like the value-initialised vector */
} orc_cgr_job;

static void orc_cgr_task(void *c, size_t i) {
    orc_cgr_job *j = (orc_cgr_job *)c;
    const size_t d = (size_t)1 << (2 * j->k);
    const char *rec = j->chars + j->start[i];
    const size_t n = (size_t)(j->start[i + 1] - j->start[i]);
    if (j->bits == 16) {
        uint16_t *o = (uint16_t *)j->out + i * d;
        memset(o, 0, d * sizeof(uint16_t));
        orc_cgr_count_u16(rec, n, j->k, o);
    } else {
        uint32_t *o = (uint32_t *)j->out + i * d;
        memset(o, 0, d * sizeof(uint32_t));
        orc_cgr_count_u32(rec, n, j->k, o);
    }
}

int orc_cgr_batch(const char *chars, const uint64_t *start, size_t n_seqs, int k, int bits,
                  void *out, int n_threads) {
    if (k < 1 || 2 * k > 64 || (bits != 16 && bits != 32)) return -1;
    orc_cgr_job job = {chars, start, k, bits, out};
    orc_execute(orc_cgr_task, &job, n_seqs, n_threads);
    return 0;
}

/* The reference `kmers` step with mode=frequencies: run<T> (above) followed by the numpy pass
 * backend.py:42-56.  Here both run inside the per-sequence task, on all threads (generous to the
 * reference: its numpy pass is single-threaded), writing float32 or float64 frequencies. */
typedef struct {
    const char *chars;
    const uint64_t *start;
    int k, bits, out_f64;
    void *out;
    void *scratch; /* n_threads? no: one count vector per task slot, allocated per task */
} orc_freq_job;

static void orc_freq_task(void *c, size_t i) {
    orc_freq_job *j = (orc_freq_job *)c;
    const size_t d = (size_t)1 << (2 * j->k);
    const char *rec = j->chars + j->start[i];
    const size_t n = (size_t)(j->start[i + 1] - j->start[i]);
    uint64_t s = 0;
    if (j->bits == 16) {
        uint16_t *cnt = (uint16_t *)calloc(d, sizeof(uint16_t));
        orc_cgr_count_u16(rec, n, j->k, cnt);
        for (size_t q = 0; q < d; q++) s += cnt[q];
        const double ds = (double)s;
        if (j->out_f64) { double *o = (double *)j->out + i * d; for (size_t q = 0; q < d; q++) o[q] = (double)cnt[q] / ds; }
        else { float *o = (float *)j->out + i * d; for (size_t q = 0; q < d; q++) o[q] = (float)((double)cnt[q] / ds); }
        free(cnt);
    } else {
        uint32_t *cnt = (uint32_t *)calloc(d, sizeof(uint32_t));
        orc_cgr_count_u32(rec, n, j->k, cnt);
        for (size_t q = 0; q < d; q++) s += cnt[q];
        const double ds = (double)s;
        if (j->out_f64) { double *o = (double *)j->out + i * d; for (size_t q = 0; q < d; q++) o[q] = (double)cnt[q] / ds; }
        else { float *o = (float *)j->out + i * d; for (size_t q = 0; q < d; q++) o[q] = (float)((double)cnt[q] / ds); }
        free(cnt);
    }
}

int orc_cgr_freq_batch(const char *chars, const uint64_t *start, size_t n_seqs, int k, int bits,
                       int out_f64, void *out, int n_threads) {
    if (k < 1 || 2 * k > 64 || (bits != 16 && bits != 32)) return -1;
    orc_freq_job job = {chars, start, k, bits, out_f64, out, NULL};
    orc_execute(orc_freq_task, &job, n_seqs, n_threads);
    return 0;
}

/* generation_dists main: one task per row i computing every j>i (dists@0x10001b8cd-0x10001b942),
 * packed strict upper triangle, index n*i - i(i+3)/2 + j - 1 (dists@0x10001bf7e-0x10001bfc0). */
typedef struct {
    const void *x;
    size_t n, d;
    int elem; /* element_type: 0 u8, 1 u16, 2 u32, 3 u64, 4 f32, 5 f64 (A.4) */
    uint32_t *manhat; /* may be NULL */
    float *info;      /* may be NULL */
} orc_dist_job;

static void orc_dist_task(void *c, size_t i) {
    orc_dist_job *j = (orc_dist_job *)c;
    const size_t n = j->n, d = j->d;
    for (size_t q = i + 1; q < n; q++) {
        const size_t idx = n * i - i * (i + 3) / 2 + q - 1;
#define ORC_ROW(T, SFX)                                                   \
    {                                                                     \
        const T *a = (const T *)j->x + i * d, *b = (const T *)j->x + q * d; \
        if (j->manhat) j->manhat[idx] = orc_manhat_##SFX(a, b, d);        \
        if (j->info) j->info[idx] = orc_info_##SFX(a, b, d);              \
    }
        switch (j->elem) {
            case 0: ORC_ROW(uint8_t, u8) break;
            case 1: ORC_ROW(uint16_t, u16) break;
            case 2: ORC_ROW(uint32_t, u32) break;
            case 3: ORC_ROW(uint64_t, u64) break;
            case 4: ORC_ROW(float, f32) break;
            default: ORC_ROW(double, f64) break;
        }
#undef ORC_ROW
    }
}

int orc_dists_triangle(const void *x, size_t n, size_t d, int elem, uint32_t *manhat, float *info,
                       int n_threads) {
    if (elem < 0 || elem > 5) return -1;
    orc_dist_job job = {x, n, d, elem, manhat, info};
    orc_execute(orc_dist_task, &job, n, n_threads);
    return 0;
}

/* Rectangular variant used for the BASELINE config-4 shape (reads x centroids); same per-pair
 * arithmetic as above, row-major m x c output.  Not a reference entry point (the reference only
 * computes the symmetric triangle). */
typedef struct {
    const void *x, *y;
    size_t m, c, d;
    int elem;
    uint32_t *manhat;
} orc_rect_job;

static void orc_rect_task(void *cx, size_t i) {
    orc_rect_job *j = (orc_rect_job *)cx;
    for (size_t q = 0; q < j->c; q++) {
        uint32_t v;
        switch (j->elem) {
            case 1: v = orc_manhat_u16((const uint16_t *)j->x + i * j->d, (const uint16_t *)j->y + q * j->d, j->d); break;
            case 2: v = orc_manhat_u32((const uint32_t *)j->x + i * j->d, (const uint32_t *)j->y + q * j->d, j->d); break;
            default: v = 0; break;
        }
        j->manhat[i * j->c + q] = v;
    }
}

int orc_manhat_rect(const void *x, const void *y, size_t m, size_t c, size_t d, int elem,
                    uint32_t *manhat, int n_threads) {
    if (elem != 1 && elem != 2) return -1;
    orc_rect_job job = {x, y, m, c, d, elem, manhat};
    orc_execute(orc_rect_task, &job, m, n_threads);
    return 0;
}
