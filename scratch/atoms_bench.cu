// micro-benchmark: shared-memory atomic increment rate (random bins) vs global RED, to decide whether a
// multi-pass shared-memory table can beat the RED.ADD spill path for megabase sequences.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>   // 0: ATOMS u32 no return, 1: ATOMS u32 with return, 2: packed u16 in u32 no return
__global__ void smem_atoms(uint32_t *out, int iters, uint32_t bins_mask) {
    extern __shared__ uint32_t tab[];
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u, acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t b = rng(s);
            if (MODE == 0) atomicAdd(&tab[b & bins_mask], 1u);
            else if (MODE == 1) acc += atomicAdd(&tab[b & bins_mask], 1u);
            else atomicAdd(&tab[(b >> 1) & bins_mask], (b & 1u) ? 0x10000u : 1u);
        }
    }
    __syncthreads();
    uint32_t t = acc;
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) t += tab[i];
    if (t == 0xdeadbeefu) out[0] = t;
}

__global__ void gmem_red(uint32_t *tab, int iters, uint32_t bins_mask, uint32_t tables) {
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    uint32_t *t = tab + (size_t)(blockIdx.x % tables) * (bins_mask + 1);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) atomicAdd(&t[rng(s) & bins_mask], 1u);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *d; cudaMalloc(&d, 256u << 20); cudaMemset(d, 0, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    auto report = [&](const char *name, double ops, float ms) {
        printf("%-44s %8.3f ms  %8.1f G incr/s  %6.2f incr/clk/SM @1.965GHz\n", name, ms, ops / ms / 1e6,
               ops / (ms * 1e-3) / sms / 1.965e9);
    };
    for (int threads : {256, 512, 1024}) {
        for (int kb : {32, 128}) {
            const uint32_t mask = kb * 256 - 1;
            const int ctas_per_sm = kb == 32 ? (threads == 1024 ? 2 : 2048 / threads > 6 ? 6 : 2048 / threads) : 1;
            const int grid = sms * ctas_per_sm;
            const double ops = (double)grid * threads * iters * 8;
            float ms;
            char nm[128];
#define RUN(MODE, label)                                                                                  \
    cudaFuncSetAttribute(smem_atoms<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);      \
    smem_atoms<MODE><<<grid, threads, kb * 1024>>>(d, 10, mask);                                          \
    cudaEventRecord(e0);                                                                                  \
    smem_atoms<MODE><<<grid, threads, kb * 1024>>>(d, iters, mask);                                       \
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);                     \
    snprintf(nm, sizeof nm, "%s thr=%d tab=%dKB cta/sm=%d", label, threads, kb, ctas_per_sm); report(nm, ops, ms);
            RUN(0, "ATOMS noret")
            RUN(1, "ATOMS ret")
            RUN(2, "ATOMS packed16 noret")
        }
    }
    for (uint32_t tables : {1u, 16u, 48u}) {
        const uint32_t mask = (1u << 18) - 1;   // 4^9 bins, 1 MB per table
        const int grid = sms * 8, threads = 256;
        gmem_red<<<grid, threads>>>(d, 10, mask, tables);
        cudaEventRecord(e0);
        gmem_red<<<grid, threads>>>(d, iters, mask, tables);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        char nm[128]; snprintf(nm, sizeof nm, "RED.global 1MB tables=%u", tables);
        report(nm, (double)grid * threads * iters * 8, ms);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
