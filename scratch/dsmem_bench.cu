// micro-benchmark: red.shared::cluster.add.u32 to random bins of random CTAs of the cluster (distributed table)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void red_cluster(uint32_t addr, uint32_t v) {
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// MODE 0: every increment to a uniformly random CTA of the cluster; MODE 1: always the own CTA (through the
// cluster address); MODE 2: plain local atomicAdd (reference)
template <int MODE>
__global__ void dsmem_atoms(uint32_t *out, int iters, uint32_t bins_mask, uint32_t csize) {
    extern __shared__ uint32_t tab[];
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) tab[i] = 0;
    cluster_sync();
    uint32_t myrank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(myrank));
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t b = rng(s);
            const uint32_t off = base + ((b & bins_mask) << 2);
            if (MODE == 2) atomicAdd(&tab[b & bins_mask], 1u);
            else red_cluster(mapa(off, MODE == 0 ? ((b >> 20) & (csize - 1)) : myrank), 1u);
        }
    }
    cluster_sync();
    uint32_t t = 0;
    for (uint32_t i = threadIdx.x; i <= bins_mask; i += blockDim.x) t += tab[i];
    if (t == 0xdeadbeefu) out[0] = t;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *d; cudaMalloc(&d, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 500, threads = 1024, kb = 128;
    const uint32_t mask = kb * 256 - 1;
    for (int csize : {1, 2, 4, 8}) {
        for (int mode = 0; mode < 3; ++mode) {
            void (*kern)(uint32_t *, int, uint32_t, uint32_t) =
                mode == 0 ? dsmem_atoms<0> : mode == 1 ? dsmem_atoms<1> : dsmem_atoms<2>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
            cudaLaunchConfig_t cfg = {};
            int grid = sms / csize * csize;
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = kb * 1024;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nclus = 0;
            cudaOccupancyMaxActiveClusters(&nclus, kern, &cfg);
            if (nclus * csize < grid) { grid = nclus * csize; cfg.gridDim = dim3(grid); }
            cudaLaunchKernelEx(&cfg, kern, d, 5, mask, (uint32_t)csize);
            cudaEventRecord(e0);
            cudaLaunchKernelEx(&cfg, kern, d, iters, mask, (uint32_t)csize);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double)grid * threads * iters * 8;
            printf("cluster=%d mode=%s ctas=%d: %8.3f ms %8.1f G incr/s %6.2f incr/clk/SM(active) err=%s\n", csize,
                   mode == 0 ? "random-cta" : mode == 1 ? "own-via-cluster" : "local-atomicAdd", grid, ms, ops / ms / 1e6,
                   ops / (ms * 1e-3) / grid / 1.965e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
