// pairop_bench.cu -- issue rate of the inner operations of pair_tile_kernel, per SM and clock, measured the
// way the kernel uses them (register operands, 64 independent accumulators per thread like its 8x8 block):
//   VABSDIFF4.U8.ACC  (__vsadu4 + add)        4 pair-elements per lane-op   -> manhat, counts <= 255
//   VABSDIFF.U32 + IADD (__usad)              1 pair-element per lane-op    -> manhat, exact mod 2^32
//   LOP3 + POPC x2 (info)                     32 pair-elements per 4 lane-ops
//   FADD |a-b| x2 (float L1)                  1 pair-element per 2 lane-ops
// These are the denominators of the manhattan / info / float-L1 rooflines in bench.py (round-1 used the
// nominal 128 lane-ops per clock and SM for all of them).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scratch/pairop_bench scratch/pairop_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[8], b[8], acc[64], acc2[64];
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        b[i] = seed * (threadIdx.x + 7) + i * 0x85ebca6bu;
    }
    for (int i = 0; i < 64; ++i) acc[i] = acc2[i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (OP == 0) acc[r * 8 + c] += __vsadu4(a[r], b[c]);
                else if (OP == 1) acc[r * 8 + c] = __usad(a[r], b[c], acc[r * 8 + c]);
                else if (OP == 2) {
                    acc[r * 8 + c] += __popc(a[r] | b[c]);
                    acc2[r * 8 + c] += __popc(a[r] ^ b[c]);
                } else {
                    float f = __uint_as_float(acc[r * 8 + c]);
                    f += fabsf(__uint_as_float(a[r]) - __uint_as_float(b[c]));
                    acc[r * 8 + c] = __float_as_uint(f);
                }
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // new operands every trip (what the shared-memory loads do in the kernel)
            a[i] = a[i] * 1664525u + 1013904223u;
            b[i] ^= a[i] >> 3;
        }
    }
    uint32_t t = 0;
    for (int i = 0; i < 64; ++i) t += acc[i] + acc2[i];
    if (t == 0x12345678u) out[0] = t;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t *d;
    cudaMalloc(&d, 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4000, ctas = p.multiProcessorCount * 8;
    const char *names[4] = {"VABSDIFF4.U8.ACC (manhat bytes)", "VABSDIFF.U32 (manhat exact)", "LOP3+POPC x2 (info)",
                            "FADD |a-b| (float L1)"};
    const double elems_per_op[4] = {4, 1, 32, 1};
    auto run = [&](int op, int it) {
        if (op == 0) bench<0><<<ctas, 256>>>(d, it, 12345u);
        if (op == 1) bench<1><<<ctas, 256>>>(d, it, 12345u);
        if (op == 2) bench<2><<<ctas, 256>>>(d, it, 12345u);
        if (op == 3) bench<3><<<ctas, 256>>>(d, it, 12345u);
    };
    for (int op = 0; op < 4; ++op) {
        run(op, 10);
        cudaEventRecord(e0);
        run(op, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)ctas * 256 * iters * 64;          // pair-operations (one per accumulator update)
        const double per_clk_sm = ops / (ms * 1e-3) / p.multiProcessorCount / (clk_khz * 1e3);
        printf("%-34s %8.3f ms  %7.2f pair-ops/clk/SM  %7.1f pair-elements/clk/SM  %8.2f T pair-elements/s\n", names[op],
               ms, per_clk_sm, per_clk_sm * elems_per_op[op], ops * elems_per_op[op] / (ms * 1e-3) / 1e12);
    }
    printf("clock %d MHz, %d SMs, err=%s\n", clk_khz / 1000, p.multiProcessorCount, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
