/*
 * refcode.c -- execute the reference's OWN machine code for the CGR fill loop.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/kameris_oracle.c header).  Runs only in the build
 * container, where /root/reference is mounted; its outputs are committed as
 * tests/golden/refcode_cgr.json by oracle/refcode/gen_golden.py.
 *
 * What it does: kameris ships its hot path only as prebuilt executables.  In
 * kameris/scripts/generation_cgr_windows_avx2.exe the per-file CGR fill is an out-of-line LEAF
 * function (no calls, no imports, position independent):
 *     VA 0x140016040 .. 0x1400161a8   (file offset 0x15440, 0x168 bytes)
 *     void fill(std::vector<uint16_t>& count, const std::string& record, unsigned k,
 *               const char* const& order)          -- Windows x64 ABI (rcx, rdx, r8d, r9)
 * called from VA 0x14001564b with order -> "ACGT" (.rdata VA 0x1400a20b1).  Its uint32 twin (the same source
 * instantiated for bits_per_element = 32: `mov rcx,[rcx]` up front, `inc DWORD PTR [rcx+rax*4]` at VA
 * 0x14001b47d) is the leaf at
 *     VA 0x14001b360 .. 0x14001b4c3   (file offset 0x1a760, 0x163 bytes), same signature with
 *     std::vector<uint32_t>&
 * and is mapped the same way for the uint32 golden vectors (refcode_cgr_u32; golden generation only).
 * We map those bytes from the file where it lies (nothing is copied into the repo), build the two
 * MSVC container headers the code dereferences ([rcx] = data pointer of the vector; the string is
 * {ptr/SSO buffer @0, size @0x10, capacity @0x18}), and call it through an ms_abi pointer.
 * x86-64 + BMI2 (shlx) host required.
 *
 * CPU baseline (bench.py cpu_baseline / --impl reference, "kind": "reference"): /root/reference does
 * not exist on the GPU box, so `make refcode` (run by __graft_entry__.build() in the build container)
 * also writes those 0x168 code bytes + the 4 order bytes to oracle/_ref/cgr_fill_u16.bin -- a build
 * output like the .so files next to it (git-ignored, not gpurun-ignored), never a tracked file --
 * and refcode_load_blob() maps it there.  refcode_cgr_u16_batch() then runs the reference's fill on all
 * host threads over a batch of records, with the task scheme of mmg::ParallelExecutor
 * (@0x10001a6f0: T = hardware_concurrency workers, one atomic task counter, one value-initialised
 * vector<uint16_t>(4^k) per record).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sys/mman.h>

#define FILL_FILE_OFF 0x15440l
#define FILL_LEN 0x168
#define FILL32_FILE_OFF 0x1a760l
#define FILL32_LEN 0x163
#define ORDER_FILE_OFF 0xa0ab1l

typedef void(__attribute__((ms_abi)) * fill_fn)(void *vec, void *str, uint32_t k, void *order);

struct msvc_string {
    char *ptr;
    uint64_t pad;
    uint64_t size;
    uint64_t cap;
};
struct msvc_vector {
    void *first, *last, *end;
};

static void *g_code = NULL, *g_code32 = NULL;
static char g_order[8];

static int map_code(const unsigned char *buf) {
    /* sanity: prologue "push rbx; push rsi; push rdi; sub rsp,0x40" and the final jmp */
    static const unsigned char prologue[] = {0x53, 0x56, 0x57, 0x48, 0x83, 0xec, 0x40};
    if (memcmp(buf, prologue, sizeof prologue) != 0 || buf[FILL_LEN - 5] != 0xe9) return -4;
    void *m = mmap(NULL, 4096, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return -5;
    memset(m, 0xcc, 4096);
    memcpy(m, buf, FILL_LEN);
    g_code = m;
    return 0;
}

static int read_pe(const char *pe_path, unsigned char *buf, char *order) {
    FILE *f = fopen(pe_path, "rb");
    if (!f) return -1;
    if (fseek(f, FILL_FILE_OFF, SEEK_SET) || fread(buf, 1, FILL_LEN, f) != FILL_LEN) {
        fclose(f);
        return -2;
    }
    if (fseek(f, ORDER_FILE_OFF, SEEK_SET) || fread(order, 1, 4, f) != 4) {
        fclose(f);
        return -3;
    }
    fclose(f);
    return 0;
}

/* returns 0 on success; negative on failure */
int refcode_load(const char *pe_path) {
    if (g_code) return 0;
    unsigned char buf[FILL_LEN];
    int rc = read_pe(pe_path, buf, g_order);
    return rc ? rc : map_code(buf);
}

/* build step: the function's bytes + the order string -> a file under oracle/_ref/ */
int refcode_extract(const char *pe_path, const char *blob_path) {
    unsigned char buf[FILL_LEN + 4];
    int rc = read_pe(pe_path, buf, (char *)buf + FILL_LEN);
    if (rc) return rc;
    FILE *f = fopen(blob_path, "wb");
    if (!f) return -6;
    rc = fwrite(buf, 1, sizeof buf, f) == sizeof buf ? 0 : -7;
    fclose(f);
    return rc;
}

/* where /root/reference does not exist: map what refcode_extract wrote */
int refcode_load_blob(const char *blob_path) {
    if (g_code) return 0;
    FILE *f = fopen(blob_path, "rb");
    if (!f) return -1;
    unsigned char buf[FILL_LEN + 4];
    const size_t got = fread(buf, 1, sizeof buf, f);
    fclose(f);
    if (got != sizeof buf) return -2;
    memcpy(g_order, buf + FILL_LEN, 4);
    return map_code(buf);
}

const char *refcode_order(void) { return g_order; }

/* count must hold 4^k zero-initialised uint16 (the caller's value-initialised vector). */
int refcode_cgr_u16(const char *rec, uint64_t n, uint32_t k, uint16_t *count) {
    if (!g_code) return -1;
    struct msvc_string s = {(char *)rec, 0, n, n < 16 ? 16 : n};
    struct msvc_vector v = {count, count + ((uint64_t)1 << (2 * k)), count + ((uint64_t)1 << (2 * k))};
    char *order = g_order;
    ((fill_fn)g_code)(&v, &s, k, &order);
    return 0;
}

/* ---- the reference's fill over a batch, on all host threads (bench.py cpu_baseline) ---------------- */
struct batch {
    const char *chars;
    const uint64_t *start;
    uint64_t n;
    uint32_t k;
    uint16_t *out;
    uint64_t next;   /* the executor's atomic task counter */
};

static void *batch_worker(void *arg) {
    struct batch *b = (struct batch *)arg;
    const uint64_t bins = (uint64_t)1 << (2 * b->k);
    for (;;) {
        const uint64_t i = __atomic_fetch_add(&b->next, 1, __ATOMIC_RELAXED);
        if (i >= b->n) break;
        uint16_t *row = b->out + i * bins;
        memset(row, 0, bins * sizeof(uint16_t));   /* std::vector<uint16_t>(4^k): value-initialised per record */
        refcode_cgr_u16(b->chars + b->start[i], b->start[i + 1] - b->start[i], b->k, row);
    }
    return NULL;
}

/* out: n x 4^k uint16.  Returns 0, or negative when the code is not loaded. */
int refcode_cgr_u16_batch(const char *chars, const uint64_t *start, uint64_t n, uint32_t k, uint16_t *out,
                          int threads) {
    if (!g_code) return -1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    struct batch b = {chars, start, n, k, out, 0};
    pthread_t th[256];
    int made = 0;
    for (; made < threads; made++)
        if (pthread_create(&th[made], NULL, batch_worker, &b)) break;
    if (made == 0) batch_worker(&b);
    for (int t = 0; t < made; t++) pthread_join(th[t], NULL);
    return 0;
}


/* ---- the uint32 twin (golden generation in the build container only) ------------------------------ */
int refcode_load_u32(const char *pe_path) {
    if (g_code32) return 0;
    static const unsigned char prologue[] = {0x53, 0x56, 0x57, 0x48, 0x83, 0xec, 0x40};
    unsigned char buf[FILL32_LEN];
    FILE *f = fopen(pe_path, "rb");
    if (!f) return -1;
    if (fseek(f, FILL32_FILE_OFF, SEEK_SET) || fread(buf, 1, FILL32_LEN, f) != FILL32_LEN) {
        fclose(f);
        return -2;
    }
    if (fseek(f, ORDER_FILE_OFF, SEEK_SET) || fread(g_order, 1, 4, f) != 4) {
        fclose(f);
        return -3;
    }
    fclose(f);
    /* prologue, the `inc DWORD PTR [rcx+rax*4]` at +0x11d, the final jmp */
    if (memcmp(buf, prologue, sizeof prologue) != 0 || buf[0x11d] != 0xff || buf[0x11e] != 0x04 || buf[0x11f] != 0x81 ||
        buf[FILL32_LEN - 5] != 0xe9)
        return -4;
    void *m = mmap(NULL, 4096, PROT_READ | PROT_WRITE | PROT_EXEC, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m == MAP_FAILED) return -5;
    memset(m, 0xcc, 4096);
    memcpy(m, buf, FILL32_LEN);
    g_code32 = m;
    return 0;
}

/* count must hold 4^k zero-initialised uint32. */
int refcode_cgr_u32(const char *rec, uint64_t n, uint32_t k, uint32_t *count) {
    if (!g_code32) return -1;
    struct msvc_string s = {(char *)rec, 0, n, n < 16 ? 16 : n};
    struct msvc_vector v = {count, count + ((uint64_t)1 << (2 * k)), count + ((uint64_t)1 << (2 * k))};
    char *order = g_order;
    ((fill_fn)g_code32)(&v, &s, k, &order);
    return 0;
}
