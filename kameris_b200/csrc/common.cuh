// common.cuh -- shared helpers for libkameris_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kameris_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "kameris_b200 targets sm_100a (B200) only"
#endif

void kam_set_error(const char *fmt, ...);

#define KAM_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            kam_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return KAM_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

#define KAM_REQUIRE(cond, msg)                 \
    do {                                       \
        if (!(cond)) {                         \
            kam_set_error("%s: %s", __func__, msg); \
            return KAM_E_INVALID;              \
        }                                      \
    } while (0)

static inline size_t kam_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
int kam_sm_count();

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// streaming 16-byte store: written once, never re-read by this kernel
__device__ __forceinline__ void st_stream16(void *p, uint4 v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
#endif
