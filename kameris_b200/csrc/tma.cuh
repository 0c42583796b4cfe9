// tma.cuh -- minimal TMA (cp.async.bulk.tensor) + mbarrier helpers, inline PTX for sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library still loads on a box without a driver).
typedef CUresult (*kam_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static inline kam_encode_tiled_fn kam_get_encode_tiled() {
    static kam_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (kam_encode_tiled_fn)p;
    }
    return fn;
}

// 2-D row-major tensor [outer][inner] of `elem_bytes`-wide elements, box {box_inner, box_outer}.
static inline int kam_make_tmap_2d(CUtensorMap *tm, CUtensorMapDataType dt, uint32_t elem_bytes, const void *base,
                                   uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner,
                                   uint32_t box_outer, CUtensorMapSwizzle swz) {
    kam_encode_tiled_fn enc = kam_get_encode_tiled();
    if (!enc) {
        kam_set_error("cuTensorMapEncodeTiled not available from the driver");
        return KAM_E_CUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)elem_bytes;
    CUresult r = enc(tm, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        kam_set_error("cuTensorMapEncodeTiled failed (CUresult %d): inner=%llu outer=%llu stride=%llu box=%ux%u",
                      (int)r, (unsigned long long)inner, (unsigned long long)outer,
                      (unsigned long long)row_stride_bytes, box_inner, box_outer);
        return KAM_E_CUDA;
    }
    return KAM_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a wrong transaction count must not hang the GPU box -- trap instead.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int32_t c_inner, int32_t c_outer,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar))
        : "memory");
}
// multicast load: the box lands at the same CTA-relative smem offset, and signals the mbarrier at
// the same CTA-relative offset, in every CTA of the cluster selected by cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *tm, int32_t c_inner, int32_t c_outer,
                                               uint64_t *bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
#endif
