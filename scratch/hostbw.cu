// hostbw.cu -- what the host side of kam_kmers_host can sustain on the GPU box: non-temporal fills of
// pinned memory on T threads, the D2H copy engine alone, and both at once (they share host DRAM).
//   nvcc -O3 -std=c++17 -Xcompiler -pthread,-mavx2 -o scratch/hostbw scratch/hostbw.cu
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum Mode { NT_SSE = 0, NT_AVX = 1, MEMSET = 2, SPARSE = 3 };

// SPARSE: a 2 MB vector of doubles with ~9000 non-zeros at sorted positions, written line by line with NT stores
static void fill(char *p, size_t bytes, int mode, const uint32_t *pos, int npos) {
    if (mode == MEMSET) {
        memset(p, 0, bytes);
    } else if (mode == NT_SSE) {
        const __m128i z = _mm_setzero_si128();
        for (size_t i = 0; i < bytes; i += 64) {
            _mm_stream_si128((__m128i *)(p + i), z);
            _mm_stream_si128((__m128i *)(p + i + 16), z);
            _mm_stream_si128((__m128i *)(p + i + 32), z);
            _mm_stream_si128((__m128i *)(p + i + 48), z);
        }
    } else if (mode == NT_AVX) {
        const __m256i z = _mm256_setzero_si256();
        for (size_t i = 0; i < bytes; i += 64) {
            _mm256_stream_si256((__m256i *)(p + i), z);
            _mm256_stream_si256((__m256i *)(p + i + 32), z);
        }
    } else {
        const size_t vec = 2u << 20;
        const __m256i z = _mm256_setzero_si256();
        for (size_t v = 0; v + vec <= bytes; v += vec) {
            char *q = p + v;
            size_t line = 0;   // next line to write
            int e = 0;
            while (e < npos) {
                const size_t l = ((size_t)pos[e] * 8) >> 6;
                for (; line < l; ++line) {
                    _mm256_stream_si256((__m256i *)(q + line * 64), z);
                    _mm256_stream_si256((__m256i *)(q + line * 64 + 32), z);
                }
                alignas(64) double tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                while (e < npos && (((size_t)pos[e] * 8) >> 6) == l) {
                    tmp[pos[e] & 7] = 1.0 / 8992.0;
                    ++e;
                }
                _mm256_stream_si256((__m256i *)(q + l * 64), *(__m256i *)&tmp[0]);
                _mm256_stream_si256((__m256i *)(q + l * 64 + 32), *(__m256i *)&tmp[4]);
                line = l + 1;
            }
            for (; line < vec / 64; ++line) {
                _mm256_stream_si256((__m256i *)(q + line * 64), z);
                _mm256_stream_si256((__m256i *)(q + line * 64 + 32), z);
            }
        }
    }
    _mm_sfence();
}

static double run_fill(char *buf, size_t bytes, int T, int mode, const uint32_t *pos, int npos) {
    std::vector<std::thread> th;
    const size_t chunk = 2u << 20;
    std::atomic<size_t> next{0};
    const double t0 = now();
    for (int t = 0; t < T; ++t)
        th.emplace_back([&] {
            for (;;) {
                const size_t o = next.fetch_add(chunk);
                if (o >= bytes) break;
                fill(buf + o, chunk, mode, pos, npos);
            }
        });
    for (auto &x : th) x.join();
    return bytes / (now() - t0) / 1e9;
}

int main(int argc, char **argv) {
    const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 8.0) * (1u << 30);
    char *h = nullptr, *h2 = nullptr, *d = nullptr;
    if (cudaMallocHost(&h, bytes) != cudaSuccess || cudaMallocHost(&h2, bytes) != cudaSuccess ||
        cudaMalloc(&d, bytes) != cudaSuccess) {
        printf("alloc failed\n");
        return 1;
    }
    memset(h, 1, bytes);
    memset(h2, 1, bytes);
    cudaMemset(d, 0, bytes);
    std::vector<uint32_t> pos;
    uint64_t s = 88172645463325252ull;
    std::vector<char> hit(262144, 0);
    for (int i = 0; i < 8992; ++i) {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        hit[s % 262144] = 1;
    }
    for (uint32_t i = 0; i < 262144; ++i)
        if (hit[i]) pos.push_back(i);
    const int hw = (int)std::thread::hardware_concurrency();
    printf("hardware_concurrency %d, buffer %.1f GB, nnz %zu\n", hw, bytes / 1e9, pos.size());
    const char *names[4] = {"nt_sse", "nt_avx2", "memset", "sparse_expand"};
    for (int mode = 0; mode < 4; ++mode)
        for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
            if (T > hw && T != 1) continue;
            double best = 0;
            for (int r = 0; r < 2; ++r) {
                const double g = run_fill(h, bytes, T, mode, pos.data(), (int)pos.size());
                if (g > best) best = g;
            }
            printf("%-14s T=%2d  %.1f GB/s\n", names[mode], T, best);
            fflush(stdout);
        }
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int r = 0; r < 2; ++r) {
        const double t0 = now();
        cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        printf("d2h alone: %.1f GB/s\n", bytes / (now() - t0) / 1e9);
    }
    for (int T : {4, 8, 12, 16, 24, 32}) {
        if (T > hw) continue;
        const double t0 = now();
        cudaMemcpyAsync(h2, d, bytes, cudaMemcpyDeviceToHost, st);
        const double g = run_fill(h, bytes, T, SPARSE, pos.data(), (int)pos.size());
        const double t1 = now();
        cudaStreamSynchronize(st);
        const double t2 = now();
        printf("concurrent T=%2d: sparse_expand %.1f GB/s (%.3f s), d2h %.1f GB/s (%.3f s)\n", T, g, t1 - t0,
               bytes / (t2 - t0) / 1e9, t2 - t0);
    }
    // H2D alongside (inputs are tiny, but for completeness)
    {
        const double t0 = now();
        cudaMemcpyAsync(d, h2, bytes, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        printf("h2d alone: %.1f GB/s\n", bytes / (now() - t0) / 1e9);
    }
    return 0;
}
