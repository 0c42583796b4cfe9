// peer.cu -- device memory that other processes of the same box can read over NVLink (CUDA IPC), and
// copy-engine transfers from it.  The multi-GPU `distances` step (kameris_b200/parallel.py: one process per
// GPU) moves shards of tensor-core operand rows between ranks while the Gram kernel of the previous block
// runs.  Doing that with NCCL send/recv puts NCCL's kernels on SMs: next to a persistent cluster-of-2 kernel
// that owns every SM's shared memory they either wait or -- landing one per TPC -- keep whole TPCs away from
// the clusters (measured: the overlapped block ran 33-100 % longer).  A pull by the copy engines from the
// owner's exported buffer uses no SM at all.
#include <cuda.h>
#include <stdlib.h>
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

// cuStreamWriteValue32 through the runtime's driver entry point (no link-time libcuda dependency): a 4-byte
// write the stream performs in order, without a kernel and without a copy-engine descriptor.
typedef CUresult (*write_value32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
write_value32_fn write_value32() {
    static write_value32_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (write_value32_fn)p;
    }();
    return fn;
}
uint32_t *g_one = nullptr;   // pinned word holding 1: the source of the flag copies when the stream write is not to be had
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

int kam_peer_signal(uint32_t *d_flag, void *stream) {
    KAM_REQUIRE(d_flag, "null pointer");
    const char *mode = getenv("KAM_PEER_FLAG");   // "copy": a 4-byte host-to-device copy instead of the stream write
    const bool by_copy = mode && strcmp(mode, "copy") == 0;
    if (write_value32_fn fn = by_copy ? nullptr : write_value32()) {
        const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)d_flag, 1u, 0u /* default: ordered after prior work */);
        if (r == CUDA_SUCCESS) return KAM_OK;
    }
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (!g_one) {
            KAM_CUDA(cudaMallocHost(&g_one, sizeof(uint32_t)));
            *g_one = 1u;
        }
    }
    KAM_CUDA(cudaMemcpyAsync(d_flag, g_one, sizeof(uint32_t), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return KAM_OK;
}

int kam_peer_pull(const kam_peer_piece *pieces, uint32_t n_pieces, uint32_t *d_flags, void *const *streams,
                  uint32_t n_streams) {
    KAM_REQUIRE(n_pieces == 0 || pieces, "null pointer");
    KAM_REQUIRE(streams && n_streams >= 1, "kam_peer_pull needs at least one stream");
    for (uint32_t i = 0; i < n_pieces; ++i) {
        const kam_peer_piece &p = pieces[i];
        KAM_REQUIRE(i == 0 || pieces[i - 1].chunk <= p.chunk || pieces[i - 1].chunk == KAM_PEER_NO_FLAG,
                    "kam_peer_pull: pieces must come in chunk order");
        cudaStream_t st = (cudaStream_t)streams[p.chunk == KAM_PEER_NO_FLAG ? 0 : p.chunk % n_streams];
        if (p.bytes) {
            KAM_REQUIRE(p.d_dst && p.d_src, "null pointer in a piece");
            KAM_CUDA(cudaMemcpyAsync(p.d_dst, p.d_src, p.bytes, cudaMemcpyDefault, st));
        }
        const bool last_of_chunk = i + 1 == n_pieces || pieces[i + 1].chunk != p.chunk;
        if (p.chunk != KAM_PEER_NO_FLAG && last_of_chunk) {
            KAM_REQUIRE(d_flags, "kam_peer_pull: flagged pieces need d_flags");
            int rc = kam_peer_signal(d_flags + p.chunk, st);
            if (rc) return rc;
        }
    }
    return KAM_OK;
}

}  // extern "C"
