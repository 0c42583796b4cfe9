// micro-benchmark: FP64 FMA issue rate per SM (independent chains), to know the bound of the mds block product
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {256, 512, 1024}) {
        const int grid = p.multiProcessorCount * (2048 / threads), iters = 20000;
        dfma<<<grid, threads>>>(d, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dfma<<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = (double)grid * threads * iters * 8;
        printf("threads=%d: %.3f ms, %.2f TFLOP/s FP64, %.1f DFMA/clk/SM @1.965 GHz\n", threads, ms, 2 * fmas / ms / 1e9,
               fmas / (ms * 1e-3) / p.multiProcessorCount / 1.965e9);
    }
    return 0;
}
