// freq_gram.cu -- euclid / cosine / pearson over FLOATING-POINT rows (what a `mode: frequencies`
// .mm-repr holds: kameris/job_steps/backend.py:45-56 rewrites every count vector as cgr / cgr.sum(),
// float64), so that the `distances` step (backend.py:59-66) takes both kinds of file.
//
// 1. Integer recovery (gram_prepass_float_kernel).  A frequency row is f_i = fl(c_i / S) with integer
//    counts c_i and S = sum c_i, one correctly rounded division per element.  The pre-pass finds the
//    smallest positive element of the row, proposes S = rint(c_min / f_min) for c_min = 1, 2, ... 8
//    (at the orders where a Gram pass is expensive -- 4^k bins >> sequence length -- the smallest
//    positive count of a row is 1), and ACCEPTS a proposal only when every element passes
//        c_i = rint(f_i * S),   fl(c_i / S) == f_i  bit for bit,   and  sum c_i == S.
//    An accepted row IS c / S to the last bit, so the exact integer Gram matrix of the recovered counts
//    (tcgen05 uint8 / fp16 path of gram.cu) gives the metrics of the stored rows, with euclid taken
//    between the frequency vectors.  The same sweep writes the tensor-core operand row, the row's S and
//    sum of squares and the envelope flags -- the recovered counts are never materialised.  Rows of
//    NaN only (an empty sequence: 0/0) become zero rows, whose distances come out NaN like scipy's.
// 2. Anything else (a row that is not a frequency vector, counts outside the tensor envelopes) runs
//    gram_f64_kernel: a plain float64 Gram on the CUDA cores (64 x 64 pair tiles, 4 x 4 accumulators per
//    thread, DFMA-bound) with the same float64 epilogue, between the rows as stored.
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "dist_internal.cuh"

namespace {

constexpr int FP_THREADS = 256;
constexpr int N_CMIN = 8;          // smallest count proposed for the row's smallest positive element

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16 consecutive elements of a row starting at e (zeros past d)
template <typename T>
__device__ __forceinline__ void load16(const T *__restrict__ xr, uint64_t e, uint64_t d, int vec_ok, T (&v)[16]) {
    if (e >= d) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (T)0;
    } else if (vec_ok) {   // d % 16 == 0 and 16-byte aligned rows
        constexpr int NV = (int)sizeof(T);   // uint4 loads per 16 elements
        uint4 w[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) w[i] = __ldg(reinterpret_cast<const uint4 *>(xr + e) + i);
        const T *t = reinterpret_cast<const T *>(w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = t[i];
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = e + i < d ? xr[e + i] : (T)0;
    }
}

// c = the integer whose quotient by S is f, if there is one
template <typename T>
__device__ __forceinline__ bool recover_count(T f, double S, unsigned &c) {
    c = 0;
    if (f == (T)0) return true;
    const double r = rint((double)f * S);
    if (!(r >= 1.0 && r <= 4294967295.0)) return false;   // negative, infinite, NaN, or not a count
    c = (unsigned)r;
    return (T)(r / S) == f;   // the division numpy (float64) / kmer_count.cu (float32) performed
}

template <typename T, typename OP>
__global__ void __launch_bounds__(FP_THREADS) gram_prepass_float_kernel(const T *__restrict__ x, uint32_t n, uint64_t d,
                                                                        uint64_t ld, uint64_t dpad, OP *__restrict__ h,
                                                                        unsigned long long *__restrict__ S,
                                                                        unsigned long long *__restrict__ G,
                                                                        unsigned long long *__restrict__ flags,
                                                                        int vec_ok) {
    __shared__ double red_d[FP_THREADS / 32];
    __shared__ unsigned long long red_u[4][FP_THREADS / 32];
    __shared__ double s_fmin;
    __shared__ int s_mode;                       // 0 = not a frequency row, 1 = zero row, 2 = try proposals
    __shared__ unsigned long long s_tot[4];      // s, g, mx, bad of one proposal
    const uint32_t row = blockIdx.x;
    const T *xr = x + (uint64_t)row * ld;
    OP *hr = h + (uint64_t)row * dpad;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool F32 = std::is_same<T, float>::value;

    // smallest positive element of the first `limit` elements, and what else is there
    auto scan = [&](uint64_t limit) {
        double fm = INFINITY;
        unsigned long long npos = 0, nnan = 0, nbad = 0;
        for (uint64_t e = (uint64_t)threadIdx.x * 16; e < limit; e += FP_THREADS * 16) {
            T v[16];
            load16<T>(xr, e, d, vec_ok, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double f = (double)v[i];
                if (f != f) ++nnan;
                else if (f < 0.0 || f == INFINITY) ++nbad;
                else if (f > 0.0) {
                    ++npos;
                    fm = fmin(fm, f);
                }
            }
        }
        fm = warp_min_d(fm);
        npos = warp_sum64(npos), nnan = warp_sum64(nnan), nbad = warp_sum64(nbad);
        if (lane == 0) red_d[warp] = fm, red_u[0][warp] = npos, red_u[1][warp] = nnan, red_u[2][warp] = nbad;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = INFINITY;
            unsigned long long p = 0, q = 0, b = 0;
            for (int w = 0; w < FP_THREADS / 32; ++w)
                m = fmin(m, red_d[w]), p += red_u[0][w], q += red_u[1][w], b += red_u[2][w];
            s_fmin = m;
            // NaN in every bin: the 0/0 of an empty sequence (the zero fill past d is not counted)
            s_mode = (q == limit) ? 1 : (b == 0 && q == 0 && p > 0) ? 2 : 0;
        }
        __syncthreads();
    };

    unsigned long long fs = 0, fg = 0, fmx = 0;
    // proposals S = rint(c_min / fm), c_min = 1..N_CMIN: verify every element, write the operand row
    auto try_proposals = [&](double fm) -> bool {
        constexpr int N_DELTA = F32 ? 3 : 1;   // a float32 quotient pins S only to within 2^-24
        for (int prop = 0; prop < N_CMIN * N_DELTA; ++prop) {
            const int cmin = prop / N_DELTA + 1, dl = prop % N_DELTA;
            const double Sd = rint((double)cmin / fm) + (dl == 0 ? 0.0 : dl == 1 ? -1.0 : 1.0);
            if (!(Sd >= 1.0 && Sd < 4503599627370496.0)) continue;   // uniform: same value in every thread
            unsigned long long s = 0, g = 0;
            unsigned mx = 0, bad = 0;
            for (uint64_t e = (uint64_t)threadIdx.x * 16; e < dpad; e += FP_THREADS * 16) {
                T v[16];
                load16<T>(xr, e, d, vec_ok, v);
                unsigned c[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (!recover_count<T>(v[i], Sd, c[i])) bad = 1;
                    s += c[i];
                    g += (unsigned long long)c[i] * c[i];
                    mx = max(mx, c[i]);
                }
                store_operand16<OP>(hr + e, c);
            }
            s = warp_sum64(s), g = warp_sum64(g);
            unsigned long long mb = ((unsigned long long)bad << 32) | mx;   // bad flag and largest count in one word
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long t = __shfl_xor_sync(0xffffffffu, mb, o);
                mb = ((mb | t) & 0xffffffff00000000ull) | max(mb & 0xffffffffull, t & 0xffffffffull);
            }
            __syncthreads();   // the previous proposal's totals have been read
            if (lane == 0) red_u[0][warp] = s, red_u[1][warp] = g, red_u[2][warp] = mb;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long ts = 0, tg = 0, tm = 0, tb = 0;
                for (int w = 0; w < FP_THREADS / 32; ++w) {
                    ts += red_u[0][w], tg += red_u[1][w];
                    tb |= red_u[2][w] >> 32;
                    tm = max(tm, red_u[2][w] & 0xffffffffull);
                }
                s_tot[0] = ts, s_tot[1] = tg, s_tot[2] = tm;
                s_tot[3] = (tb != 0 || (double)ts != Sd) ? 1ull : 0ull;
            }
            __syncthreads();
            if (s_tot[3] == 0) {
                fs = s_tot[0], fg = s_tot[1], fmx = s_tot[2];
                return true;
            }
        }
        return false;
    };

    bool done = false;
    double tried = 0.0;
    // Long rows: the smallest positive element of a short prefix is nearly always a count of 1 already, so
    // the proposals are formed after reading 1/16 of the row or less and the row is swept ONCE (verification
    // does not depend on where the proposal came from); only if they fail is the whole row scanned.
    constexpr uint64_t PREFIX = 16384;
    if (d >= 16 * PREFIX) {
        scan(PREFIX);
        if (s_mode == 2) {
            tried = s_fmin;
            done = try_proposals(tried);
        }
    }
    if (!done) {
        scan(d);
        const int mode = s_mode;
        const double fm = s_fmin;
        if (mode == 1) {
            for (uint64_t e = (uint64_t)threadIdx.x * 16; e < dpad; e += FP_THREADS * 16) {
                unsigned c[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = 0;
                store_operand16<OP>(hr + e, c);
            }
            done = true;
        } else if (mode == 2 && fm != tried) {
            done = try_proposals(fm);
        }
    }
    if (threadIdx.x == 0) {
        if (done) {
            S[row] = fs;
            G[row] = fg;
            atomicMax(&flags[0], fmx);
            atomicMax(&flags[1], fg);
        } else {
            S[row] = 0;
            G[row] = 0;
            atomicAdd(&flags[2], 1ull);   // rows that are not frequency vectors: the caller takes the float64 kernel
        }
    }
}

// ---- float64 Gram on the CUDA cores ---------------------------------------------------------------
// per-row constants from the rows as stored: S = sum x, G = sum x^2 (float64)
template <typename T>
__global__ void __launch_bounds__(FP_THREADS) rowk_float_kernel(const T *__restrict__ x, uint32_t n, uint64_t d,
                                                                uint64_t ld, RowK *__restrict__ rk) {
    __shared__ double red[2][FP_THREADS / 32];
    const uint32_t row = blockIdx.x;
    const T *xr = x + (uint64_t)row * ld;
    double s = 0.0, g = 0.0;
    for (uint64_t e = threadIdx.x; e < d; e += FP_THREADS) {
        const double f = (double)xr[e];
        s += f;
        g = fma(f, f, g);
    }
    s = warp_sum_d(s), g = warp_sum_d(g);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s, red[1][threadIdx.x >> 5] = g;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tg = 0.0;
        for (int w = 0; w < FP_THREADS / 32; ++w) ts += red[0][w], tg += red[1][w];
        const double D = (double)d;
        RowK r;
        r.S = ts;
        r.mu = ts / D;
        r.rn = 1.0 / sqrt(tg);
        r.rp = 1.0 / sqrt(tg - ts * ts / D);
        r.a = tg;
        r.q = 1.0;
        rk[row] = r;
    }
}

constexpr int FT = 64, FK = 16, FPAD = 2;

__device__ __forceinline__ long long tri_base_f(long long n, long long i) { return n * i - i * (i + 3) / 2 - 1; }

template <typename T>
__global__ void __launch_bounds__(256, 3) gram_f64_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                       uint32_t row_begin, uint32_t row_end,
                                                       const RowK *__restrict__ rk, GramOut go) {
    const uint32_t bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;                                   // strict upper triangle only
    const uint32_t r0 = bi * FT, c0 = bj * FT;
    if (r0 >= row_end || r0 + FT <= row_begin) return;     // another rank's rows
    __shared__ __align__(16) double As[2][FK][FT + FPAD];
    __shared__ __align__(16) double Bs[2][FK][FT + FPAD];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    // staging: thread (lk, lr) loads element k0 + lk of rows lr, lr + 16, lr + 32, lr + 48 -- a half-warp reads
    // 16 consecutive elements of one row
    const int lk = threadIdx.x & 15, lr = threadIdx.x >> 4;
    T ra[4], rb[4];
    const T *pa = x + (uint64_t)(r0 + lr) * ld + lk, *pb = x + (uint64_t)(c0 + lr) * ld + lk;
    const uint64_t ld16 = 16 * ld;
    uint32_t ok = 0;   // bit i: row lr + 16 i of the A tile exists; bit 4 + i: of the B tile
#pragma unroll
    for (int i = 0; i < 4; ++i) ok |= (r0 + lr + 16 * i < n ? 1u << i : 0u) | (c0 + lr + 16 * i < n ? 16u << i : 0u);
    auto gload = [&](uint64_t k0) {
        const bool kin = k0 + lk < d;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ra[i] = (kin && (ok >> i & 1u)) ? __ldg(pa + k0 + i * ld16) : (T)0;
            rb[i] = (kin && (ok >> (4 + i) & 1u)) ? __ldg(pb + k0 + i * ld16) : (T)0;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[buf][lk][lr + 16 * i] = (double)ra[i];
            Bs[buf][lk][lr + 16 * i] = (double)rb[i];
        }
    };
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    // thread (ty, tx) owns rows ty*4 .. ty*4+3 and columns tx*2, tx*2+1, 32+tx*2, 32+tx*2+1: the 16 lanes of a
    // half-warp read 16 consecutive 16-byte pieces of a B row (no bank conflicts), the A values are broadcasts
    const uint64_t nk = (d + FK - 1) / FK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (uint64_t kt = 0; kt < nk; ++kt) {
        const int buf = (int)(kt & 1);
        if (kt + 1 < nk) gload((kt + 1) * FK);
#pragma unroll
        for (int k = 0; k < FK; ++k) {
            const double2 a01 = *reinterpret_cast<const double2 *>(&As[buf][k][ty * 4]);
            const double2 a23 = *reinterpret_cast<const double2 *>(&As[buf][k][ty * 4 + 2]);
            const double2 b01 = *reinterpret_cast<const double2 *>(&Bs[buf][k][tx * 2]);
            const double2 b23 = *reinterpret_cast<const double2 *>(&Bs[buf][k][32 + tx * 2]);
            const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) sstore(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t gi = r0 + ty * 4 + i;
        if (gi >= n || gi < row_begin || gi >= row_end) continue;
        const long long base = tri_base_f(n, gi);
        const RowK ri = rk[gi];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t gj = c0 + (j < 2 ? tx * 2 + j : 32 + tx * 2 + (j - 2));
            if (gj >= n || gj <= gi) continue;
            gram_store(go, base + gj, acc[i][j], ri, rk[gj]);
        }
    }
}

template <typename T>
int run_float_prepass(const void *x, uint32_t n, uint64_t d, uint64_t ld, uint64_t dpad, bool i8, void *h,
                      unsigned long long *S, unsigned long long *G, unsigned long long *flags, cudaStream_t st) {
    const int vec_ok = (d % 16 == 0) && ((ld * sizeof(T)) % 16 == 0) && (((uintptr_t)x & 15) == 0);
    if (i8)
        gram_prepass_float_kernel<T, uint8_t><<<n, FP_THREADS, 0, st>>>((const T *)x, n, d, ld, dpad, (uint8_t *)h, S, G,
                                                                         flags, vec_ok);
    else
        gram_prepass_float_kernel<T, __half><<<n, FP_THREADS, 0, st>>>((const T *)x, n, d, ld, dpad, (__half *)h, S, G,
                                                                        flags, vec_ok);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

template <typename T>
int run_float_gram(const void *x, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin, uint32_t row_end, RowK *rk,
                   GramOut go, cudaStream_t st) {
    rowk_float_kernel<T><<<n, FP_THREADS, 0, st>>>((const T *)x, n, d, ld, rk);
    KAM_CUDA(cudaGetLastError());
    const uint32_t tc = (n + FT - 1) / FT;
    gram_f64_kernel<T><<<dim3(tc, tc), 256, 0, st>>>((const T *)x, n, d, ld, row_begin, row_end, rk, go);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

}  // namespace

int kam_float_prepass(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint64_t dpad, bool i8, void *h,
                      unsigned long long *S, unsigned long long *G, unsigned long long *flags, cudaStream_t st) {
    return elem == KAM_F32 ? run_float_prepass<float>(d_x, n, d, ld, dpad, i8, h, S, G, flags, st)
                           : run_float_prepass<double>(d_x, n, d, ld, dpad, i8, h, S, G, flags, st);
}

int kam_float_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                       uint32_t row_end, RowK *rk, GramOut go, cudaStream_t st) {
    if (n > 65535u * FT) {
        kam_set_error("kam_gram_tri: too many rows for the float64 path");
        return KAM_E_INVALID;
    }
    return elem == KAM_F32 ? run_float_gram<float>(d_x, n, d, ld, row_begin, row_end, rk, go, st)
                           : run_float_gram<double>(d_x, n, d, ld, row_begin, row_end, rk, go, st);
}
