// host_kmers.cu -- kam_kmers_host: the whole `kmers` step on HOST buffers.
//
// Replaces `generation_cgr cgr <dir> <out> <k> <bits>` + the numpy frequency pass as driven by
// kameris/job_steps/backend.py:34-56 (file reading/writing stays with the caller): ASCII records
// in host memory -> count / frequency vectors in host memory.  The batch is cut into chunks of
// sequences; each chunk goes H2D -> pack -> k-mer count -> D2H on one of NSLOT streams with its
// own device buffers, so the (dominant) device-to-host copy of chunk c overlaps the kernels and
// the host-to-device copy of chunk c+1.  Pinned caller buffers are copied directly; pageable
// ones go through internal pinned staging.
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int NSLOT = 3;
constexpr size_t OUT_CHUNK_BYTES = 128ull << 20;
constexpr size_t CHAR_CHUNK_BYTES = 256ull << 20;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KAM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = kam_align_up(bytes + bytes / 8, 1 << 20);
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            kam_set_error("cudaMalloc(%zu) failed", want);
            return KAM_E_NOMEM;
        }
        cap = want;
        return KAM_OK;
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KAM_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = kam_align_up(bytes + bytes / 8, 1 << 20);
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            kam_set_error("cudaMallocHost(%zu) failed", want);
            return KAM_E_NOMEM;
        }
        cap = want;
        return KAM_OK;
    }
};

struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;   // D2H of the chunk last issued on this slot
    DevBuf chars, starts, planes, base, head, pack_ws, kmer_ws, out;
    PinBuf h_starts, h_chars, h_out;
    // pending copy-out from pinned staging to a pageable caller buffer
    void *pend_dst = nullptr;
    size_t pend_bytes = 0;
};

struct HostCtx {
    int device = -1;
    Slot slot[NSLOT];
};

HostCtx g_ctx;

int ctx_init() {
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    if (g_ctx.device == dev && g_ctx.slot[0].st) return KAM_OK;
    for (auto &s : g_ctx.slot) {
        if (!s.st) KAM_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        if (!s.done) KAM_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    g_ctx.device = dev;
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

int slot_drain(Slot &s) {
    KAM_CUDA(cudaEventSynchronize(s.done));
    if (s.pend_dst) {
        memcpy(s.pend_dst, s.h_out.p, s.pend_bytes);
        s.pend_dst = nullptr;
    }
    return KAM_OK;
}

}  // namespace

extern "C" int kam_kmers_host(const uint8_t *h_chars, const uint64_t *h_char_start, uint32_t n_seqs, int k,
                              int count_bits, int out_kind, void *h_out) {
    if (k < 1 || 2 * k > 64) {
        kam_set_error("k is too large to fit in the index");
        return KAM_E_INVALID;
    }
    KAM_REQUIRE(k <= 15, "k > 15 not supported by this build");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind >= KAM_OUT_COUNTS && out_kind <= KAM_OUT_FREQ_F64, "bad out_kind");
    KAM_REQUIRE(h_char_start && h_out, "null pointer");
    if (n_seqs == 0) return KAM_OK;
    int rc = ctx_init();
    if (rc) return rc;

    const size_t bins = (size_t)1 << (2 * k);
    const size_t esz = out_kind == KAM_OUT_COUNTS ? (size_t)count_bits / 8 : (out_kind == KAM_OUT_FREQ_F32 ? 4 : 8);
    const size_t row_bytes = bins * esz;
    const bool out_pinned = is_pinned(h_out);
    const bool in_pinned = h_chars && is_pinned(h_chars);

    uint32_t max_rows = (uint32_t)std::max<size_t>(1, OUT_CHUNK_BYTES / row_bytes);
    uint32_t first = 0;
    int turn = 0;
    bool used[NSLOT] = {false, false, false};
    while (first < n_seqs) {
        // chunk [first, last): bounded by rows and by characters
        uint32_t last = first;
        const uint64_t c0 = h_char_start[first];
        while (last < n_seqs && last - first < max_rows &&
               (last == first || h_char_start[last + 1] - c0 <= CHAR_CHUNK_BYTES))
            ++last;
        const uint32_t n = last - first;
        const uint64_t nchars = h_char_start[last] - c0;

        Slot &s = g_ctx.slot[turn];
        if (used[turn] && (rc = slot_drain(s))) return rc;
        used[turn] = true;

        if ((rc = s.chars.ensure(nchars + 32)) || (rc = s.starts.ensure((size_t)(n + 1) * 8)) ||
            (rc = s.planes.ensure(kam_planes_words(nchars) * 8)) || (rc = s.base.ensure((size_t)(n + 1) * 8)) ||
            (rc = s.head.ensure((size_t)n * 4)) || (rc = s.pack_ws.ensure(kam_pack_workspace_bytes(nchars, n))) ||
            (rc = s.kmer_ws.ensure(kam_kmer_workspace_bytes(n, k))) || (rc = s.out.ensure((size_t)n * row_bytes)) ||
            (rc = s.h_starts.ensure((size_t)(n + 1) * 8)))
            return rc;

        uint64_t *hs = (uint64_t *)s.h_starts.p;
        for (uint32_t i = 0; i <= n; ++i) hs[i] = h_char_start[first + i] - c0;
        KAM_CUDA(cudaMemcpyAsync(s.starts.p, hs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s.st));
        if (nchars) {
            const void *src = h_chars + c0;
            if (!in_pinned) {
                if ((rc = s.h_chars.ensure(nchars))) return rc;
                memcpy(s.h_chars.p, src, nchars);
                src = s.h_chars.p;
            }
            KAM_CUDA(cudaMemcpyAsync(s.chars.p, src, nchars, cudaMemcpyHostToDevice, s.st));
        }
        rc = kam_pack_ascii((const uint8_t *)s.chars.p, (const uint64_t *)s.starts.p, n, nchars, (uint64_t *)s.planes.p,
                            (uint64_t *)s.base.p, (uint32_t *)s.head.p, s.pack_ws.p, s.pack_ws.cap, s.st);
        if (rc) return rc;
        rc = kam_kmer_count((const uint64_t *)s.planes.p, (const uint64_t *)s.base.p, (const uint32_t *)s.head.p, n,
                            nchars, k, count_bits, out_kind, s.out.p, bins, s.kmer_ws.p, s.kmer_ws.cap, s.st);
        if (rc) return rc;
        char *dst = (char *)h_out + (size_t)first * row_bytes;
        if (out_pinned) {
            KAM_CUDA(cudaMemcpyAsync(dst, s.out.p, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, s.st));
        } else {
            if ((rc = s.h_out.ensure((size_t)n * row_bytes))) return rc;
            KAM_CUDA(cudaMemcpyAsync(s.h_out.p, s.out.p, (size_t)n * row_bytes, cudaMemcpyDeviceToHost, s.st));
            s.pend_dst = dst;
            s.pend_bytes = (size_t)n * row_bytes;
        }
        KAM_CUDA(cudaEventRecord(s.done, s.st));
        first = last;
        turn = (turn + 1) % NSLOT;
    }
    for (int i = 0; i < NSLOT; ++i)
        if (used[i] && (rc = slot_drain(g_ctx.slot[i]))) return rc;
    return KAM_OK;
}
