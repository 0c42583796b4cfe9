// cgr_pool.cu -- two small device-to-device steps around the k-mer tables:
//
//   kam_cgr_pool   order-(k+1) CGR counts -> order-k counts by 2x2 sum-pooling of the image plus the one
//                  k-mer that has no (k+1)-mer ending at the same base (follows from the
//                  emit rule of generation_cgr @0x10000e2a8: base j emits iff j >= lead(k), and
//                  lead(k+1) = lead(k) + [raw character k-1 is a base]).  Works on wrapped uint16 counts
//                  too: pooling commutes with arithmetic mod 2^16.
//   kam_normalize  counts -> frequencies, cgr / cgr.sum() (kameris/job_steps/backend.py:49), for count
//                  matrices that are already on the device (kam_kmer_count fuses this for fresh tables).
//
// Both are one read + one write of their tables: HBM-bound, 16-byte accesses where alignment allows.
#include <type_traits>

#include "common.cuh"

namespace {

// bin of the k-mer (full or partial) that ends at base j of the sequence -- the literal recurrence
__device__ uint32_t bin_of_base(const uint2 *__restrict__ planes, uint64_t p0, uint32_t j, int k) {
    uint32_t x = 1u << (k - 1), y = x;
    for (uint32_t t = 0; t <= j; ++t) {
        const uint64_t p = p0 + t;
        const uint2 w = planes[p >> 5];
        const uint32_t b = (uint32_t)(p & 31);
        x = (x >> 1) | (((w.x >> b) & 1u) << (k - 1));
        y = (y >> 1) | (((w.y >> b) & 1u) << (k - 1));
    }
    return (y << k) | x;
}

template <typename T>
__global__ void __launch_bounds__(256) cgr_pool_kernel(const uint2 *__restrict__ planes,
                                                       const uint64_t *__restrict__ base_start,
                                                       const uint32_t *__restrict__ head_mask, int k,
                                                       const T *__restrict__ in, uint64_t in_stride,
                                                       T *__restrict__ out, uint64_t out_stride) {
    __shared__ uint32_t s_fix;   // bin that gets the extra +1, or 0xffffffff
    const uint32_t s = blockIdx.y;
    const uint32_t side = 1u << k, bins = side * side;
    if (threadIdx.x == 0) {
        const uint64_t p0 = base_start[s], n = base_start[s + 1] - p0;
        const uint32_t hm = head_mask[s];
        const uint32_t lead = __popc(hm & ((k > 1 ? (1u << (k - 1)) : 1u) - 1u));
        const bool raw_valid = (hm >> (k - 1)) & 1u;     // raw character k-1 is a base
        s_fix = (raw_valid && lead < n) ? bin_of_base(planes, p0, lead, k) : 0xffffffffu;
    }
    __syncthreads();
    const uint32_t fix = s_fix;
    const T *src = in + (uint64_t)s * in_stride;
    T *dst = out + (uint64_t)s * out_stride;
    for (uint32_t b = blockIdx.x * 256 + threadIdx.x; b < bins; b += gridDim.x * 256) {
        const uint32_t y = b >> k, x = b & (side - 1);
        const uint64_t r0 = ((uint64_t)(2 * y) << (k + 1)) + 2 * x, r1 = r0 + (2ull << k);
        uint32_t v = (uint32_t)src[r0] + (uint32_t)src[r0 + 1] + (uint32_t)src[r1] + (uint32_t)src[r1 + 1];
        if (b == fix) v += 1u;
        dst[b] = (T)v;   // wraps mod 2^bits like the reference's counters
    }
}

template <typename T, typename OUT>
__global__ void __launch_bounds__(256) normalize_kernel(const T *__restrict__ in, uint64_t in_stride, uint64_t d,
                                                        OUT *__restrict__ out, uint64_t out_stride) {
    __shared__ unsigned long long red[8];
    __shared__ double s_sum;
    const T *src = in + (uint64_t)blockIdx.x * in_stride;
    OUT *dst = out + (uint64_t)blockIdx.x * out_stride;
    unsigned long long part = 0;
    for (uint64_t i = threadIdx.x; i < d; i += 256) part += src[i];
    part = warp_sum64(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        s_sum = (double)t;    // numpy sums unsigned integers in uint64 and divides in float64
    }
    __syncthreads();
    const double sum = s_sum;
    for (uint64_t i = threadIdx.x; i < d; i += 256) dst[i] = (OUT)((double)src[i] / sum);
}

}  // namespace

extern "C" {

int kam_cgr_pool(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask, uint32_t n_seqs,
                 int k, int count_bits, const void *d_in, uint64_t in_stride, void *d_out, uint64_t out_stride,
                 void *stream) {
    KAM_REQUIRE(d_planes && d_base_start && d_head_mask && d_in && d_out, "null pointer");
    KAM_REQUIRE(k >= 1 && k <= 14, "k out of range (the input is the order k+1 table)");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(in_stride >= (1ull << (2 * k + 2)) && out_stride >= (1ull << (2 * k)), "stride smaller than the table");
    if (n_seqs == 0) return KAM_OK;
    const uint32_t bins = 1u << (2 * k);
    uint32_t bx = (bins + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid(bx, n_seqs);
    cudaStream_t st = (cudaStream_t)stream;
    if (count_bits == 16)
        cgr_pool_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint2 *)d_planes, d_base_start, d_head_mask, k,
                                                        (const uint16_t *)d_in, in_stride, (uint16_t *)d_out, out_stride);
    else
        cgr_pool_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint2 *)d_planes, d_base_start, d_head_mask, k,
                                                        (const uint32_t *)d_in, in_stride, (uint32_t *)d_out, out_stride);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

int kam_normalize(const void *d_counts, int count_bits, uint32_t n, uint64_t d, uint64_t in_stride, int out_kind,
                  void *d_out, uint64_t out_stride, void *stream) {
    KAM_REQUIRE(d_counts && d_out, "null pointer");
    KAM_REQUIRE(count_bits == 16 || count_bits == 32, "count_bits must be 16 or 32");
    KAM_REQUIRE(out_kind == KAM_OUT_FREQ_F32 || out_kind == KAM_OUT_FREQ_F64, "out_kind must be a frequency type");
    KAM_REQUIRE(in_stride >= d && out_stride >= d, "stride smaller than d");
    if (n == 0 || d == 0) return KAM_OK;
    cudaStream_t st = (cudaStream_t)stream;
#define KAM_NORM(T, OUT) normalize_kernel<T, OUT><<<n, 256, 0, st>>>((const T *)d_counts, in_stride, d, (OUT *)d_out, out_stride)
    if (count_bits == 16) {
        if (out_kind == KAM_OUT_FREQ_F32) KAM_NORM(uint16_t, float);
        else KAM_NORM(uint16_t, double);
    } else {
        if (out_kind == KAM_OUT_FREQ_F32) KAM_NORM(uint32_t, float);
        else KAM_NORM(uint32_t, double);
    }
#undef KAM_NORM
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

}  // extern "C"
