// host_dists.cu -- kam_dists_host: the whole `distances` step on HOST buffers.
//
// Replaces `generation_dists <in> <prefix> <names>` as driven by kameris/job_steps/backend.py:59-66
// (file reading/writing stays with the caller): the n x d matrix of a .mm-repr in host memory ->
// packed strict-upper-triangle distance arrays in host memory (the .mm-dist payloads).
#include "common.cuh"

namespace {
struct Buf {
    void *p = nullptr;
    ~Buf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
            cudaGetLastError();
            kam_set_error("cudaMalloc(%zu) failed", bytes);
            p = nullptr;
            return KAM_E_NOMEM;
        }
        return KAM_OK;
    }
};
}  // namespace

extern "C" int kam_dists_host(const void *h_x, int elem, uint32_t n, uint64_t d, unsigned metrics, uint32_t *h_manhat,
                              float *h_info, float *h_euclid, float *h_cosine, float *h_pearson) {
    KAM_REQUIRE(h_x, "null input");
    KAM_REQUIRE(elem >= KAM_U8 && elem <= KAM_F64, "bad element type");
    KAM_REQUIRE(metrics != 0 && (metrics & ~(KAM_M_MANHAT | KAM_M_INFO | KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON)) == 0,
                "bad metrics mask");
    KAM_REQUIRE(!(metrics & KAM_M_MANHAT) || h_manhat, "null manhat output");
    KAM_REQUIRE(!(metrics & KAM_M_INFO) || h_info, "null info output");
    KAM_REQUIRE(!(metrics & KAM_M_EUCLID) || h_euclid, "null euclid output");
    KAM_REQUIRE(!(metrics & KAM_M_COSINE) || h_cosine, "null cosine output");
    KAM_REQUIRE(!(metrics & KAM_M_PEARSON) || h_pearson, "null pearson output");
    KAM_REQUIRE(d >= 1, "Vectors must have the same length");   // generation_dists @0x10002809c
    if (n < 2) return KAM_OK;
    static const size_t esz[6] = {1, 2, 4, 8, 4, 8};
    const size_t pairs = (size_t)n * (n - 1) / 2;
    cudaStream_t st = nullptr;
    KAM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } guard{st};

    Buf x, tri, ws;
    int rc;
    if ((rc = x.alloc((size_t)n * d * esz[elem])) || (rc = tri.alloc(pairs * 4))) return rc;
    KAM_CUDA(cudaMemcpyAsync(x.p, h_x, (size_t)n * d * esz[elem], cudaMemcpyHostToDevice, st));

    if (metrics & KAM_M_MANHAT) {
        const size_t wb = kam_manhattan_workspace_bytes(n, d);
        if ((rc = ws.alloc(wb))) return rc;
        if ((rc = kam_manhattan_tri(x.p, elem, n, d, d, 0, n, (uint32_t *)tri.p, ws.p, wb, st))) return rc;
        KAM_CUDA(cudaMemcpyAsync(h_manhat, tri.p, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
        cudaFree(ws.p);
        ws.p = nullptr;
    }
    if (metrics & KAM_M_INFO) {
        const size_t wb = kam_info_workspace_bytes(n, d);
        if ((rc = ws.alloc(wb))) return rc;
        if ((rc = kam_info_tri(x.p, elem, n, d, d, 0, n, (float *)tri.p, ws.p, wb, st))) return rc;
        KAM_CUDA(cudaMemcpyAsync(h_info, tri.p, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
        cudaFree(ws.p);
        ws.p = nullptr;
    }
    const unsigned gm = metrics & (KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON);
    if (gm) {
        Buf t2, t3;
        size_t wb = kam_gram_workspace_bytes(n, d);
        if ((rc = ws.alloc(wb))) return rc;
        float *pe = (float *)tri.p, *pc = nullptr, *pp = nullptr;
        if (gm & KAM_M_COSINE) {
            if ((rc = t2.alloc(pairs * 4))) return rc;
            pc = (float *)t2.p;
        }
        if (gm & KAM_M_PEARSON) {
            if ((rc = t3.alloc(pairs * 4))) return rc;
            pp = (float *)t3.p;
        }
        // euclid is the distance between the rows as stored, like scipy euclidean on the rows of the
        // file: count vectors (euclid_on_freq = 0), or the frequency vectors of a `frequencies` file
        // (floating-point rows: kam_gram_tri recovers the counts and sets euclid_on_freq itself)
        rc = kam_gram_tri(x.p, elem, n, d, d, 0, n, gm, 0, pe, pc, pp, ws.p, wb, st);
        if (rc == KAM_E_WORKSPACE) {   // outside the tensor-exact envelope: the integer path needs more room
            cudaFree(ws.p);
            ws.p = nullptr;
            wb = kam_gram_workspace_bytes_int(n, d);
            if ((rc = ws.alloc(wb))) return rc;
            rc = kam_gram_tri(x.p, elem, n, d, d, 0, n, gm, 0, pe, pc, pp, ws.p, wb, st);
        }
        if (rc) return rc;
        if (gm & KAM_M_EUCLID) KAM_CUDA(cudaMemcpyAsync(h_euclid, pe, pairs * 4, cudaMemcpyDeviceToHost, st));
        if (gm & KAM_M_COSINE) KAM_CUDA(cudaMemcpyAsync(h_cosine, pc, pairs * 4, cudaMemcpyDeviceToHost, st));
        if (gm & KAM_M_PEARSON) KAM_CUDA(cudaMemcpyAsync(h_pearson, pp, pairs * 4, cudaMemcpyDeviceToHost, st));
        KAM_CUDA(cudaStreamSynchronize(st));
    }
    return KAM_OK;
}
