// peer.cu -- device memory that other processes of the same box can read over NVLink (CUDA IPC), and
// copy-engine transfers from it.  The multi-GPU `distances` step (kameris_b200/parallel.py: one process per
// GPU) moves shards of tensor-core operand rows between ranks while the Gram kernel of the previous block
// runs.  Doing that with NCCL send/recv puts NCCL's kernels on SMs: next to a persistent cluster-of-2 kernel
// that owns every SM's shared memory they either wait or -- landing one per TPC -- keep whole TPCs away from
// the clusters (measured: the overlapped block ran 33-100 % longer).  A pull by the copy engines from the
// owner's exported buffer uses no SM at all.
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace {
struct Opened {
    unsigned char handle[KAM_PEER_HANDLE_BYTES];
    void *mapped = nullptr;
    int dev = -1;
};
constexpr int MAX_OPEN = 64;
Opened g_open[MAX_OPEN];
std::mutex g_mu;
}  // namespace

static_assert(sizeof(cudaIpcMemHandle_t) == KAM_PEER_HANDLE_BYTES, "handle size");

extern "C" {

int kam_peer_alloc(size_t bytes, void **d_ptr) {
    KAM_REQUIRE(d_ptr, "null pointer");
    *d_ptr = nullptr;
    if (cudaMalloc(d_ptr, bytes ? bytes : 256) != cudaSuccess) {
        cudaGetLastError();
        kam_set_error("cudaMalloc(%zu) failed", bytes);
        return KAM_E_NOMEM;
    }
    return KAM_OK;
}

int kam_peer_free(void *d_ptr) {
    if (d_ptr) KAM_CUDA(cudaFree(d_ptr));
    return KAM_OK;
}

int kam_peer_export(const void *d_ptr, unsigned char *handle) {
    KAM_REQUIRE(d_ptr && handle, "null pointer");
    cudaIpcMemHandle_t h;
    KAM_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return KAM_OK;
}

int kam_peer_open(const unsigned char *handle, void **d_mapped) {
    KAM_REQUIRE(handle && d_mapped, "null pointer");
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_mu);
    Opened *free_slot = nullptr;
    for (auto &o : g_open) {
        if (o.mapped && o.dev == dev && memcmp(o.handle, handle, KAM_PEER_HANDLE_BYTES) == 0) {
            *d_mapped = o.mapped;
            return KAM_OK;
        }
        if (!o.mapped && !free_slot) free_slot = &o;
    }
    if (!free_slot) {
        kam_set_error("kam_peer_open: too many open peer buffers (close some)");
        return KAM_E_NOMEM;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    KAM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    memcpy(free_slot->handle, handle, KAM_PEER_HANDLE_BYTES);
    free_slot->mapped = p;
    free_slot->dev = dev;
    *d_mapped = p;
    return KAM_OK;
}

int kam_peer_close(void *d_mapped) {
    if (!d_mapped) return KAM_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &o : g_open)
        if (o.mapped == d_mapped) {
            o.mapped = nullptr;
            KAM_CUDA(cudaIpcCloseMemHandle(d_mapped));
            return KAM_OK;
        }
    kam_set_error("kam_peer_close: not a pointer of kam_peer_open");
    return KAM_E_INVALID;
}

int kam_peer_copy(void *d_dst, const void *d_src, size_t bytes, void *stream) {
    KAM_REQUIRE(bytes == 0 || (d_dst && d_src), "null pointer");
    if (bytes) KAM_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return KAM_OK;
}

}  // extern "C"
