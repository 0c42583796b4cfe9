// host_dists.cu -- kam_dists_host / kam_dists_device: the whole `distances` step with HOST results.
//
// Replaces `generation_dists <in> <prefix> <names>` as driven by kameris/job_steps/backend.py:59-66
// (file reading/writing stays with the caller): the n x d matrix of a .mm-repr -> packed
// strict-upper-triangle distance arrays in host memory (the .mm-dist payloads).  kam_dists_host takes the
// matrix from host memory (what the step reads from the file); kam_dists_device takes it where the `kmers`
// step of the same job left it on the device (kam_kmers_host_keep), so a kmers -> distances job neither
// re-reads the .mm-repr nor uploads it again.  The device buffers are kept between calls (they used to be
// allocated and freed inside every call: ~100 ms for the first calls of a process).
#include <mutex>

#include "common.cuh"

namespace {
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KAM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = kam_align_up(bytes + bytes / 16 + 256, 1 << 20);
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
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

struct DistCtx {
    cudaStream_t st = nullptr;
    Buf x, tri[3], ws;
};
constexpr int MAX_DEVICES = 64;
DistCtx g_ctx[MAX_DEVICES];
std::mutex g_mu;

int ctx_get(DistCtx **out) {
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) {
        kam_set_error("device index %d out of range", dev);
        return KAM_E_INVALID;
    }
    DistCtx &c = g_ctx[dev];
    if (!c.st) KAM_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    *out = &c;
    return KAM_OK;
}

int check_args(int elem, uint64_t d, unsigned metrics, const void *h_manhat, const void *h_info, const void *h_euclid,
               const void *h_cosine, const void *h_pearson) {
    KAM_REQUIRE(elem >= KAM_U8 && elem <= KAM_F64, "bad element type");
    KAM_REQUIRE(metrics != 0 && (metrics & ~(KAM_M_MANHAT | KAM_M_INFO | KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON)) == 0,
                "bad metrics mask");
    KAM_REQUIRE(!(metrics & KAM_M_MANHAT) || h_manhat, "null manhat output");
    KAM_REQUIRE(!(metrics & KAM_M_INFO) || h_info, "null info output");
    KAM_REQUIRE(!(metrics & KAM_M_EUCLID) || h_euclid, "null euclid output");
    KAM_REQUIRE(!(metrics & KAM_M_COSINE) || h_cosine, "null cosine output");
    KAM_REQUIRE(!(metrics & KAM_M_PEARSON) || h_pearson, "null pearson output");
    KAM_REQUIRE(d >= 1, "Vectors must have the same length");   // generation_dists @0x10002809c
    return KAM_OK;
}

// the matrix is on the device (row stride ld elements); results go to the host
int dists_locked(DistCtx &c, const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, unsigned metrics,
                 int euclid_on_freq, uint32_t *h_manhat, float *h_info, float *h_euclid, float *h_cosine,
                 float *h_pearson) {
    const size_t pairs = (size_t)n * (n - 1) / 2;
    cudaStream_t st = c.st;
    int rc;
    if ((rc = c.tri[0].ensure(pairs * 4))) return rc;
    if (metrics & KAM_M_MANHAT) {
        // room for the tensor path up to its own limit (column maxima summing to 12 d) where the device has it to
        // spare, else the standard size (4 d; heavier data then takes the CUDA-core kernel)
        size_t wb = kam_manhattan_workspace_bytes(n, d);
        {
            const size_t want = kam_manhattan_workspace_bytes_levels(n, d, 12);
            size_t free_b = 0, total_b = 0;
            if (want > c.ws.cap && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && want < (free_b + c.ws.cap) / 2) wb = want;
            else if (want <= c.ws.cap) wb = want;
            cudaGetLastError();
        }
        if ((rc = c.ws.ensure(wb))) return rc;
        if ((rc = kam_manhattan_tri(d_x, elem, n, d, ld, 0, n, (uint32_t *)c.tri[0].p, c.ws.p, c.ws.cap, st))) return rc;
        KAM_CUDA(cudaMemcpyAsync(h_manhat, c.tri[0].p, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
    }
    if (metrics & KAM_M_INFO) {
        const size_t wb = kam_info_workspace_bytes(n, d);
        if ((rc = c.ws.ensure(wb))) return rc;
        if ((rc = kam_info_tri(d_x, elem, n, d, ld, 0, n, (float *)c.tri[0].p, c.ws.p, c.ws.cap, st))) return rc;
        KAM_CUDA(cudaMemcpyAsync(h_info, c.tri[0].p, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
    }
    const unsigned gm = metrics & (KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON);
    if (gm) {
        if ((rc = c.ws.ensure(kam_gram_workspace_bytes(n, d)))) return rc;
        float *pe = (float *)c.tri[0].p, *pc = nullptr, *pp = nullptr;
        if (gm & KAM_M_COSINE) {
            if ((rc = c.tri[1].ensure(pairs * 4))) return rc;
            pc = (float *)c.tri[1].p;
        }
        if (gm & KAM_M_PEARSON) {
            if ((rc = c.tri[2].ensure(pairs * 4))) return rc;
            pp = (float *)c.tri[2].p;
        }
        // euclid is the distance between the rows as stored, like scipy euclidean on the rows of the
        // file: count vectors (euclid_on_freq = 0), or the frequency vectors of a `frequencies` file
        // (floating-point rows: kam_gram_tri recovers the counts and sets euclid_on_freq itself)
        rc = kam_gram_tri(d_x, elem, n, d, ld, 0, n, gm, euclid_on_freq, pe, pc, pp, c.ws.p, c.ws.cap, st);
        if (rc == KAM_E_WORKSPACE) {   // outside the tensor-exact envelopes: the limb / integer paths need more room
            if ((rc = c.ws.ensure(kam_gram_workspace_bytes_int(n, d)))) return rc;
            rc = kam_gram_tri(d_x, elem, n, d, ld, 0, n, gm, euclid_on_freq, pe, pc, pp, c.ws.p, c.ws.cap, st);
        }
        if (rc) return rc;
        if (gm & KAM_M_EUCLID) KAM_CUDA(cudaMemcpyAsync(h_euclid, pe, pairs * 4, cudaMemcpyDeviceToHost, st));
        if (gm & KAM_M_COSINE) KAM_CUDA(cudaMemcpyAsync(h_cosine, pc, pairs * 4, cudaMemcpyDeviceToHost, st));
        if (gm & KAM_M_PEARSON) KAM_CUDA(cudaMemcpyAsync(h_pearson, pp, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
    }
    return KAM_OK;
}
}  // namespace

extern "C" {

int kam_dists_host(const void *h_x, int elem, uint32_t n, uint64_t d, unsigned metrics, uint32_t *h_manhat, float *h_info,
                   float *h_euclid, float *h_cosine, float *h_pearson) {
    KAM_REQUIRE(h_x, "null input");
    int rc = check_args(elem, d, metrics, h_manhat, h_info, h_euclid, h_cosine, h_pearson);
    if (rc) return rc;
    if (n < 2) return KAM_OK;
    static const size_t esz[6] = {1, 2, 4, 8, 4, 8};
    std::lock_guard<std::mutex> lock(g_mu);
    DistCtx *c = nullptr;
    if ((rc = ctx_get(&c))) return rc;
    const size_t bytes = (size_t)n * d * esz[elem];
    if ((rc = c->x.ensure(bytes))) return rc;
    KAM_CUDA(cudaMemcpyAsync(c->x.p, h_x, bytes, cudaMemcpyHostToDevice, c->st));
    rc = dists_locked(*c, c->x.p, elem, n, d, d, metrics, 0, h_manhat, h_info, h_euclid, h_cosine, h_pearson);
    if (rc) {
        cudaStreamSynchronize(c->st);   // nothing of this call may still be copying when it returns
        cudaGetLastError();
    }
    return rc;
}

int kam_dists_device(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, unsigned metrics, int euclid_on_freq,
                     uint32_t *h_manhat, float *h_info, float *h_euclid, float *h_cosine, float *h_pearson) {
    KAM_REQUIRE(d_x, "null input");
    int rc = check_args(elem, d, metrics, h_manhat, h_info, h_euclid, h_cosine, h_pearson);
    if (rc) return rc;
    KAM_REQUIRE(ld >= d, "row stride smaller than the row");
    if (n < 2) return KAM_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    DistCtx *c = nullptr;
    if ((rc = ctx_get(&c))) return rc;
    rc = dists_locked(*c, d_x, elem, n, d, ld, metrics, euclid_on_freq, h_manhat, h_info, h_euclid, h_cosine, h_pearson);
    if (rc) {
        cudaStreamSynchronize(c->st);
        cudaGetLastError();
    }
    return rc;
}

void kam_host_release(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &c : g_ctx) {
        if (c.st) cudaStreamSynchronize(c.st);
        c.x.release();
        c.ws.release();
        for (auto &t : c.tri) t.release();
    }
}

}  // extern "C"
