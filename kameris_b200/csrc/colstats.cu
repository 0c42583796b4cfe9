// colstats.cu -- the feature scaling in front of the classifiers.
//
// kameris/job_steps/classify.py:29-50 (build_pipeline) puts `StandardScaler(with_mean=False)` and a
// TruncatedSVD in front of every classifier; the scaler's arithmetic lives in scikit-learn
// (sklearn/utils/extmath.py `_incremental_mean_and_var`, first batch; sklearn/preprocessing/_data.py
// StandardScaler.partial_fit / transform):
//     n_c   = number of non-NaN entries of column c
//     T_c   = sum_i x_ic / n_c                                  (mean_)
//     var_c = (sum_i (x_ic - T_c)^2 - (sum_i (x_ic - T_c))^2 / n_c) / n_c      (var_, float64)
//     transform: x_ic / scale_c                                 (a division, float64)
// Here: the feature matrix (count vectors uint16/uint32 or frequency vectors float32/float64, n rows of d
// columns, row-major) is swept twice, one thread per column (a warp reads 32 consecutive elements of a
// row), rows cut into slabs so that the grid fills the machine; the per-slab partial sums are combined in
// slab order by a second small kernel, so the result does not depend on scheduling.  HBM-bound:
// 2 x n x d x sizeof(element) bytes for the statistics, n x d x (sizeof(element) + 8) for the transform.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr int CS_THREADS = 256;

template <typename T>
__device__ __forceinline__ bool is_nan_elem(T v) {
    if constexpr (std::is_floating_point<T>::value) return v != v;
    else return false;
}

// VEC adjacent columns per thread, fetched with one VEC*sizeof(T)-byte load (uint16: 2, uint8: 4), so that a
// warp reads 128 bytes of a row per instruction
template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) ColPack {
    T v[VEC];
};

// pass 1: per-slab column sums and non-NaN counts
template <typename T, int VEC>
__global__ void __launch_bounds__(CS_THREADS) col_sum_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                             uint32_t rows_per_slab, double *__restrict__ psum,
                                                             uint32_t *__restrict__ pcnt) {
    const uint64_t c = ((uint64_t)blockIdx.x * CS_THREADS + threadIdx.x) * VEC;
    if (c >= d) return;
    const uint32_t r0 = blockIdx.y * rows_per_slab;
    const uint32_t r1 = min(n, r0 + rows_per_slab);
    double s[VEC];
    uint32_t cnt[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[j] = 0.0, cnt[j] = 0;
    const T *p = x + (uint64_t)r0 * ld + c;
#pragma unroll 8
    for (uint32_t r = r0; r < r1; ++r, p += ld) {
        const ColPack<T, VEC> w = *reinterpret_cast<const ColPack<T, VEC> *>(p);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (!is_nan_elem<T>(w.v[j])) {
                s[j] += (double)w.v[j];
                ++cnt[j];
            }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        psum[(uint64_t)blockIdx.y * d + c + j] = s[j];
        pcnt[(uint64_t)blockIdx.y * d + c + j] = cnt[j];
    }
}

__global__ void col_mean_kernel(const double *__restrict__ psum, const uint32_t *__restrict__ pcnt, uint32_t slabs,
                                uint64_t d, double *__restrict__ mean, long long *__restrict__ count) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double s = 0.0;
    long long cnt = 0;
    for (uint32_t b = 0; b < slabs; ++b) {
        s += psum[(uint64_t)b * d + c];
        cnt += pcnt[(uint64_t)b * d + c];
    }
    mean[c] = s / (double)cnt;     // 0/0 = NaN for a column of NaN only, like numpy
    count[c] = cnt;
}

// pass 2: per-slab sums of (x - T) and (x - T)^2
template <typename T, int VEC>
__global__ void __launch_bounds__(CS_THREADS) col_dev_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                             uint32_t rows_per_slab, const double *__restrict__ mean,
                                                             double *__restrict__ pcorr, double *__restrict__ psq) {
    const uint64_t c = ((uint64_t)blockIdx.x * CS_THREADS + threadIdx.x) * VEC;
    if (c >= d) return;
    const uint32_t r0 = blockIdx.y * rows_per_slab;
    const uint32_t r1 = min(n, r0 + rows_per_slab);
    double m[VEC], corr[VEC], sq[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) m[j] = mean[c + j], corr[j] = 0.0, sq[j] = 0.0;
    const T *p = x + (uint64_t)r0 * ld + c;
#pragma unroll 8
    for (uint32_t r = r0; r < r1; ++r, p += ld) {
        const ColPack<T, VEC> w = *reinterpret_cast<const ColPack<T, VEC> *>(p);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (!is_nan_elem<T>(w.v[j])) {
                const double t = (double)w.v[j] - m[j];
                corr[j] += t;
                sq[j] = __dadd_rn(sq[j], __dmul_rn(t, t));   // temp **= 2 is a rounded product: no fused multiply-add
            }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        pcorr[(uint64_t)blockIdx.y * d + c + j] = corr[j];
        psq[(uint64_t)blockIdx.y * d + c + j] = sq[j];
    }
}

__global__ void col_var_kernel(const double *__restrict__ pcorr, const double *__restrict__ psq, uint32_t slabs, uint64_t d,
                               const long long *__restrict__ count, double *__restrict__ var) {
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    double corr = 0.0, sq = 0.0;
    for (uint32_t b = 0; b < slabs; ++b) {
        corr += pcorr[(uint64_t)b * d + c];
        sq += psq[(uint64_t)b * d + c];
    }
    const double nn = (double)count[c];
    var[c] = (sq - corr * corr / nn) / nn;
}

// transform: out = x / scale, float64.  Four rows per trip: the loads are issued before the (long) divisions
template <typename T>
__global__ void __launch_bounds__(CS_THREADS) col_scale_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                               const double *__restrict__ scale, double *__restrict__ out,
                                                               uint64_t out_ld) {
    const uint64_t c = (uint64_t)blockIdx.x * CS_THREADS + threadIdx.x;
    if (c >= d) return;
    const double s = scale[c];
    // x / s, correctly rounded, without the general division sequence (the kernel was issue-bound on it): with
    // rc = RN(1/s), q = RN(x rc), the exact remainder e = x - q s (one FMA) and q' = RN(q + e rc) is the
    // correctly rounded quotient (Markstein's final step; what the hardware-assisted sequence ends with too).
    // Only for a normal divisor and finite operands well inside the exponent range -- anything else divides.
    const double rc = __drcp_rn(s);
    const bool fast = s > 1e-100 && s < 1e100;
    const uint32_t gy = gridDim.y;
    for (uint32_t r = blockIdx.y; r < n; r += 4 * gy) {
        T v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = r + i * gy < n ? __ldg(x + (uint64_t)(r + i * gy) * ld + c) : (T)0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (r + i * gy >= n) continue;
            const double a = (double)v[i];
            double q;
            if (fast && (a == 0.0 || (fabs(a) > 1e-100 && fabs(a) < 1e100))) {
                q = __dmul_rn(a, rc);
                const double e = __fma_rn(-q, s, a);
                q = __fma_rn(e, rc, q);
            } else {
                q = a / s;          // NaN, infinities, extreme magnitudes
            }
            __stcs(out + (uint64_t)(r + i * gy) * out_ld + c, q);
        }
    }
}

inline uint32_t slab_count(uint32_t n, uint64_t d) {
    const uint64_t col_blocks = (d + CS_THREADS - 1) / CS_THREADS;   // (an upper bound when columns are vectorised)
    uint64_t s = ((uint64_t)kam_sm_count() * 8 + col_blocks - 1) / col_blocks;   // ~8 CTAs per SM in flight
    if (s > n) s = n;
    if (s > 4096) s = 4096;
    if (s < 1) s = 1;
    return (uint32_t)s;
}

template <typename T, int VEC>
int run_stats_v(const void *x, uint32_t n, uint64_t d, uint64_t ld, double *mean, double *var, long long *count, void *ws,
                cudaStream_t st) {
    const uint32_t slabs = slab_count(n, d);
    const uint32_t rps = (n + slabs - 1) / slabs;
    const uint32_t used = (n + rps - 1) / rps;          // slabs that hold rows
    double *pa = (double *)ws;
    double *pb = pa + (size_t)slabs * d;
    uint32_t *pc = (uint32_t *)(pb + (size_t)slabs * d);
    const uint64_t threads = (d + VEC - 1) / VEC;
    const dim3 grid((unsigned)((threads + CS_THREADS - 1) / CS_THREADS), used);
    const unsigned cb = (unsigned)((d + 255) / 256);
    col_sum_kernel<T, VEC><<<grid, CS_THREADS, 0, st>>>((const T *)x, n, d, ld, rps, pa, pc);
    col_mean_kernel<<<cb, 256, 0, st>>>(pa, pc, used, d, mean, count);
    col_dev_kernel<T, VEC><<<grid, CS_THREADS, 0, st>>>((const T *)x, n, d, ld, rps, mean, pa, pb);
    col_var_kernel<<<cb, 256, 0, st>>>(pa, pb, used, d, count, var);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

template <typename T>
int run_stats(const void *x, uint32_t n, uint64_t d, uint64_t ld, double *mean, double *var, long long *count, void *ws,
              cudaStream_t st) {
    constexpr int VEC = sizeof(T) < 4 ? 4 / (int)sizeof(T) : 1;
    if constexpr (VEC > 1) {
        if (d % VEC == 0 && ld % VEC == 0 && ((uintptr_t)x % (VEC * sizeof(T))) == 0)
            return run_stats_v<T, VEC>(x, n, d, ld, mean, var, count, ws, st);
    }
    return run_stats_v<T, 1>(x, n, d, ld, mean, var, count, ws, st);
}

template <typename T>
int run_scale(const void *x, uint32_t n, uint64_t d, uint64_t ld, const double *scale, double *out, uint64_t out_ld,
              cudaStream_t st) {
    uint32_t gy = (uint32_t)kam_sm_count() * 8;
    if (gy > n) gy = n;
    const dim3 grid((unsigned)((d + CS_THREADS - 1) / CS_THREADS), gy);
    col_scale_kernel<T><<<grid, CS_THREADS, 0, st>>>((const T *)x, n, d, ld, scale, out, out_ld);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

}  // namespace

extern "C" {

size_t kam_colstats_workspace_bytes(uint32_t n, uint64_t d) {
    const size_t slabs = slab_count(n ? n : 1, d ? d : 1);
    return slabs * (size_t)d * (8 + 8 + 4) + 64;
}

int kam_col_stats(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, double *d_mean, double *d_var,
                  int64_t *d_count, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_mean && d_var && d_count && d_ws, "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32 || elem == KAM_F32 || elem == KAM_F64,
                "element type must be uint8/16/32 or float32/64");
    KAM_REQUIRE(n >= 1 && d >= 1 && ld >= d, "bad n / d / ld");
    KAM_REQUIRE(((d + CS_THREADS - 1) / CS_THREADS) <= 0x7fffffffull, "too many columns");
    if (ws_bytes < kam_colstats_workspace_bytes(n, d)) {
        kam_set_error("kam_col_stats: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long *cnt = (long long *)d_count;
    switch (elem) {
        case KAM_U8: return run_stats<uint8_t>(d_x, n, d, ld, d_mean, d_var, cnt, d_ws, st);
        case KAM_U16: return run_stats<uint16_t>(d_x, n, d, ld, d_mean, d_var, cnt, d_ws, st);
        case KAM_U32: return run_stats<uint32_t>(d_x, n, d, ld, d_mean, d_var, cnt, d_ws, st);
        case KAM_F32: return run_stats<float>(d_x, n, d, ld, d_mean, d_var, cnt, d_ws, st);
        default: return run_stats<double>(d_x, n, d, ld, d_mean, d_var, cnt, d_ws, st);
    }
}

int kam_col_scale(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const double *d_scale, double *d_out,
                  uint64_t out_ld, void *stream) {
    KAM_REQUIRE(d_x && d_scale && d_out, "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32 || elem == KAM_F32 || elem == KAM_F64,
                "element type must be uint8/16/32 or float32/64");
    KAM_REQUIRE(d >= 1 && ld >= d && out_ld >= d, "bad d / ld");
    if (n == 0) return KAM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (elem) {
        case KAM_U8: return run_scale<uint8_t>(d_x, n, d, ld, d_scale, d_out, out_ld, st);
        case KAM_U16: return run_scale<uint16_t>(d_x, n, d, ld, d_scale, d_out, out_ld, st);
        case KAM_U32: return run_scale<uint32_t>(d_x, n, d, ld, d_scale, d_out, out_ld, st);
        case KAM_F32: return run_scale<float>(d_x, n, d, ld, d_scale, d_out, out_ld, st);
        default: return run_scale<double>(d_x, n, d, ld, d_scale, d_out, out_ld, st);
    }
}

}  // extern "C"
