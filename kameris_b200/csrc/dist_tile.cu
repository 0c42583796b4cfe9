// dist_tile.cu -- pairwise distances that are NOT contractions: manhattan (reference `manhat`,
// uint32-exact), Jaccard-on-presence (reference `info`, float32) and float32 L1 to centroids.
//
// Replaces the row lambdas of generation_dists (generation_dists_mac_avx2 @0x10001b8c0; manhat
// body @0x10001bd87-0x10001bfc8, info body @0x10001bab5-0x10001bcb0), one task per row i over
// all j>i, and its packed-triangle store (index n*i - i(i+3)/2 + j - 1, @0x10001bf7e).
//
// Design: a pre-pass rewrites the operand rows into a k-major word matrix Xt[kw][row]
// (uint32 words: one element, four uint8 elements, 32 presence bits, or one float), so a
// (rows x k-chunk) tile is a plain 2-D box that TMA (cp.async.bulk.tensor) drops into shared
// memory behind an mbarrier, 4 stages deep.  Each CTA owns a TM x 128 tile of pairs; each thread
// an RM x 8 register block and reads its a/b fragments with conflict-free LDS.128.  The inner op
// is one VABSDIFF.U32 (|a-b|+acc) per pair-element, one VABSDIFF4.U8.ACC per four when every
// count fits a byte, LOP3+POPC for Jaccard, 2 FADD for float L1.  Bound: CUDA-core issue rate
// (not HBM, not tensor: |a-b| is not a contraction).
#include <type_traits>

#include "common.cuh"
#include "dist_internal.cuh"
#include "tma.cuh"

namespace {

enum { OP_SAD32 = 0, OP_SAD8X4 = 1, OP_JACC = 2, OP_L1F32 = 3, OP_DOT = 4 };

constexpr int DT_THREADS = 256;
constexpr int TN = 128;
constexpr int KC = 16;
constexpr int STAGES = 4;

template <int OP>
struct OpT {
    static constexpr int RM = (OP == OP_JACC || OP == OP_DOT || OP == OP_L1F32) ? 4 : 8;
    // float L1 keeps a per-stage partial (acc[0]) and the running total (acc[1]): two-level summation
    // holds the float32 result within ~1e-6 relative over thousands of terms
    static constexpr int NACC = (OP == OP_JACC || OP == OP_L1F32) ? 2 : 1;
    using acc_t = typename std::conditional<OP == OP_L1F32, float,
                  typename std::conditional<OP == OP_DOT, unsigned long long, uint32_t>::type>::type;
    using out_t = typename std::conditional<OP == OP_JACC || OP == OP_L1F32, float, uint32_t>::type;
};

template <int OP>
__device__ __forceinline__ void pair_op(typename OpT<OP>::acc_t *acc, uint32_t a, uint32_t b) {
    if constexpr (OP == OP_SAD32) acc[0] = __usad(a, b, acc[0]);
    else if constexpr (OP == OP_SAD8X4) {
        // the accumulating form, spelled out: with `acc += __vsadu4(a, b)` the compiler issued VABSDIFF4 into a
        // zero accumulator and folded pairs of results into the sum with IADD3 -- 1.5 ALU-pipe instructions per
        // pair-op instead of 1 (SASS of round 1; the kernel sat at 0.70 of the measured VABSDIFF4 rate)
        asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[0]) : "r"(a), "r"(b));
    }
    else if constexpr (OP == OP_JACC) {
        acc[0] += __popc(a | b);   // union
        acc[1] += __popc(a ^ b);   // union - intersection
    } else if constexpr (OP == OP_DOT) acc[0] += (unsigned long long)a * b;   // IMAD.WIDE.U32: exact integer Gram
    else acc[0] += fabsf(__uint_as_float(a) - __uint_as_float(b));
}

// packed-triangle offset of row i: index(i, j) = tri_base(i) + j
__device__ __forceinline__ long long tri_base(long long n, long long i) { return n * i - i * (i + 3) / 2 - 1; }

template <int OP, bool TRI>
__global__ void __launch_bounds__(DT_THREADS) pair_tile_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               uint32_t kw_total, uint32_t n_rows, uint32_t n_cols,
                                                               uint32_t row_begin, uint32_t row_end,
                                                               uint32_t tile_row0,
                                                               typename OpT<OP>::out_t *__restrict__ out,
                                                               const RowK *__restrict__ rowk, GramOut go) {
    constexpr int RM = OpT<OP>::RM, NACC = OpT<OP>::NACC, TM = 16 * RM;
    using acc_t = typename OpT<OP>::acc_t;
    extern __shared__ unsigned char smem_raw[];
    // TMA destinations: align the carve-up to 1024 B (the launch adds 1024 B of slack)
    unsigned char *smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint32_t *sA = reinterpret_cast<uint32_t *>(smem_al);    // [STAGES][KC][TM]
    uint32_t *sB = sA + STAGES * KC * TM;                    // [STAGES][KC][TN]
    __shared__ __align__(8) uint64_t full[STAGES];

    const uint32_t bi = tile_row0 + blockIdx.y, bj = blockIdx.x;
    const uint32_t r0 = bi * TM, c0 = bj * TN;
    if (TRI && c0 + TN - 1 <= r0) return;                    // no pair i<j inside this tile
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const uint32_t n_it = (kw_total + KC - 1) / KC;
    constexpr uint32_t STAGE_BYTES = (uint32_t)(KC * (TM + TN) * 4);

    if (tid == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (uint32_t it = 0; it < (uint32_t)STAGES && it < n_it; ++it) {
            mbar_arrive_expect_tx(&full[it], STAGE_BYTES);
            tma_load_2d(sA + it * KC * TM, &tmA, (int32_t)r0, (int32_t)(it * KC), &full[it]);
            tma_load_2d(sB + it * KC * TN, &tmB, (int32_t)c0, (int32_t)(it * KC), &full[it]);
        }
    }

    acc_t acc[RM][8][NACC];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[i][j][q] = 0;

    for (uint32_t it = 0; it < n_it; ++it) {
        const uint32_t s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1u);
        const uint32_t *a_base = sA + s * KC * TM, *b_base = sB + s * KC * TN;
#pragma unroll 4
        for (int kk = 0; kk < KC; ++kk) {
            uint32_t a[RM], b[8];
            const uint4 a0 = *reinterpret_cast<const uint4 *>(a_base + kk * TM + ty * 4);
            a[0] = a0.x, a[1] = a0.y, a[2] = a0.z, a[3] = a0.w;
            if constexpr (RM == 8) {
                const uint4 a1 = *reinterpret_cast<const uint4 *>(a_base + kk * TM + 64 + ty * 4);
                a[4] = a1.x, a[5] = a1.y, a[6] = a1.z, a[7] = a1.w;
            }
            const uint4 b0 = *reinterpret_cast<const uint4 *>(b_base + kk * TN + tx * 4);
            const uint4 b1 = *reinterpret_cast<const uint4 *>(b_base + kk * TN + 64 + tx * 4);
            b[0] = b0.x, b[1] = b0.y, b[2] = b0.z, b[3] = b0.w;
            b[4] = b1.x, b[5] = b1.y, b[6] = b1.z, b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) pair_op<OP>(acc[i][j], a[i], b[j]);
        }
        if constexpr (OP == OP_L1F32) {
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc[i][j][NACC - 1] += acc[i][j][0];
                    acc[i][j][0] = 0;
                }
        }
        __syncthreads();   // every thread is done with stage s
        if (tid == 0 && it + STAGES < n_it) {
            const uint32_t nx = it + STAGES;
            mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
            tma_load_2d(sA + s * KC * TM, &tmA, (int32_t)r0, (int32_t)(nx * KC), &full[s]);
            tma_load_2d(sB + s * KC * TN, &tmB, (int32_t)c0, (int32_t)(nx * KC), &full[s]);
        }
    }

#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const uint32_t r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= n_rows || r < row_begin || r >= row_end) continue;
        const long long base = TRI ? tri_base(n_rows, r) : (long long)r * n_cols;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t c = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c >= n_cols || (TRI && c <= r)) continue;
            if constexpr (OP == OP_DOT) {
                gram_store(go, base + c, (double)acc[i][j][0], rowk[r], rowk[c]);
            } else if constexpr (OP == OP_JACC) {
                // generation_dists @0x10001bc13: U==0 ? 0 : (1.0f/(float)U) * (float)(U-I), single precision
                const uint32_t U = acc[i][j][0], X = acc[i][j][1];
                out[base + c] = U == 0 ? 0.0f : __fmul_rn(__fdiv_rn(1.0f, (float)U), (float)X);
            } else {
                out[base + c] = acc[i][j][NACC - 1];
            }
        }
    }
}

// ---- operand pre-pass: row-major x[n][d] (type T) -> k-major words Xt[kw][np] ---------------------
template <typename T, int OP>
__device__ __forceinline__ uint32_t make_word(const T *__restrict__ row, uint64_t kw, uint64_t d) {
    if constexpr (OP == OP_SAD32) {
        return kw < d ? (uint32_t)row[kw] : 0u;
    } else if constexpr (OP == OP_SAD8X4) {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint64_t e = kw * 4 + b;
            if (e < d) w |= ((uint32_t)row[e] & 0xffu) << (8 * b);
        }
        return w;
    } else if constexpr (OP == OP_JACC) {
        uint32_t w = 0;
        for (int b = 0; b < 32; ++b) {
            const uint64_t e = kw * 32 + b;
            if (e < d && row[e] > (T)0) w |= 1u << b;
        }
        return w;
    } else if constexpr (OP == OP_DOT) {
        return kw < d ? (uint32_t)row[kw] : 0u;
    } else {
        return kw < d ? __float_as_uint((float)row[kw]) : 0u;
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(256) to_kmajor_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                        uint32_t kw_total, uint32_t np, uint32_t *__restrict__ xt) {
    __shared__ uint32_t tile[32][33];
    const uint32_t k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (uint32_t ry = threadIdx.y; ry < 32; ry += 8) {
        const uint32_t row = r0 + ry, kw = k0 + threadIdx.x;
        tile[ry][threadIdx.x] = (row < n && kw < kw_total) ? make_word<T, OP>(x + (uint64_t)row * ld, kw, d) : 0u;
    }
    __syncthreads();
    for (uint32_t ky = threadIdx.y; ky < 32; ky += 8) {
        const uint32_t kw = k0 + ky, col = r0 + threadIdx.x;
        if (kw < kw_total && col < np) xt[(uint64_t)kw * np + col] = tile[threadIdx.x][ky];
    }
}

template <typename T>
__global__ void max_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    const uint64_t total = (uint64_t)n * d;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = (uint32_t)x[(i / d) * ld + (i % d)];
        m = v > m ? v : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Faithful serial form for element types the tile kernel does not cover (u64, and the truncating
// float accumulate of quirk Q5): one thread per pair, sequential over d like the reference.
template <typename T>
__global__ void manhat_serial_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                                     uint32_t row_end, uint32_t *__restrict__ tri) {
    // rows in gridDim.x (up to 2^31 - 1 blocks); the column blocks of a row, at most n / 128, in gridDim.y
    const uint32_t i = row_begin + blockIdx.x;
    if (i >= row_end) return;
    const T *a = x + (uint64_t)i * ld;
    for (uint32_t j = blockIdx.y * blockDim.x + threadIdx.x; j < n; j += gridDim.y * blockDim.x) {
    if (j <= i) continue;
    const T *b = x + (uint64_t)j * ld;
    uint32_t acc = 0;
    for (uint64_t e = 0; e < d; ++e) {
        if constexpr (std::is_same<T, float>::value) {
            // single-precision sum, truncating conversion through int64 (generation_dists T=float,
            // windows_avx2 VA 0x140027a20: vcvtsi2ss / vsubss / vaddss / vcvttss2si)
            const float t = __fsub_rn(a[e], b[e]);
            const float m = b[e] > a[e] ? -t : t;
            acc = (uint32_t)__float2ll_rz(__fadd_rn(__uint2float_rn(acc), m));
        } else if constexpr (std::is_same<T, double>::value) {
            const double t = __dsub_rn(a[e], b[e]);
            const double m = b[e] > a[e] ? -t : t;
            acc = (uint32_t)__double2ll_rz(__dadd_rn((double)acc, m));
        } else {
            acc += (uint32_t)(a[e] > b[e] ? a[e] - b[e] : b[e] - a[e]);
        }
    }
    tri[tri_base(n, i) + j] = acc;
    }
}

inline uint32_t pad4(uint32_t n) { return (n + 3u) & ~3u; }

inline uint32_t kw_of(int op, uint64_t d) {
    if (op == OP_SAD8X4) return (uint32_t)((d + 3) / 4);
    if (op == OP_JACC) return (uint32_t)((d + 31) / 32);
    return (uint32_t)d;
}

template <typename T, int OP>
int convert_rows(const void *x, uint32_t n, uint64_t d, uint64_t ld, uint32_t *xt, cudaStream_t st) {
    const uint32_t kw = kw_of(OP, d), np = pad4(n);
    dim3 grid((kw + 31) / 32, (np + 31) / 32), block(32, 8);
    to_kmajor_kernel<T, OP><<<grid, block, 0, st>>>((const T *)x, n, d, ld, kw, np, xt);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

template <int OP>
int convert_dispatch(const void *x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *xt, cudaStream_t st) {
    switch (elem) {
        case KAM_U8: return convert_rows<uint8_t, OP>(x, n, d, ld, xt, st);
        case KAM_U16: return convert_rows<uint16_t, OP>(x, n, d, ld, xt, st);
        case KAM_U32: return convert_rows<uint32_t, OP>(x, n, d, ld, xt, st);
        case KAM_U64:
            if (OP == OP_JACC) return convert_rows<uint64_t, OP>(x, n, d, ld, xt, st);
            break;
        case KAM_F32:
            if (OP == OP_JACC || OP == OP_L1F32) return convert_rows<float, OP>(x, n, d, ld, xt, st);
            break;
        case KAM_F64:
            if (OP == OP_JACC) return convert_rows<double, OP>(x, n, d, ld, xt, st);
            break;
    }
    kam_set_error("element type %d not supported for this metric", elem);
    return KAM_E_UNSUPPORTED;
}

template <int OP, bool TRI>
int launch_pairs(const uint32_t *at, uint32_t n_rows, const uint32_t *bt, uint32_t n_cols, uint32_t kw,
                 uint32_t row_begin, uint32_t row_end, void *out, cudaStream_t st, const RowK *rowk = nullptr,
                 GramOut go = GramOut{nullptr, nullptr, nullptr}) {
    constexpr int TM = 16 * OpT<OP>::RM;
    CUtensorMap tmA, tmB;
    const uint32_t npa = pad4(n_rows), npb = pad4(n_cols);
    int rc = kam_make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, at, npa, kw, (uint64_t)npa * 4, TM, KC,
                              CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = kam_make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, bt, npb, kw, (uint64_t)npb * 4, TN, KC,
                          CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    const size_t smem = (size_t)STAGES * KC * (TM + TN) * 4 + 1024;
    auto kern = pair_tile_kernel<OP, TRI>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (row_end > n_rows) row_end = n_rows;
    if (row_begin >= row_end) return KAM_OK;
    const uint32_t t0 = row_begin / TM, t1 = (row_end - 1) / TM + 1;
    dim3 grid((n_cols + TN - 1) / TN, t1 - t0);
    kern<<<grid, DT_THREADS, smem, st>>>(tmA, tmB, kw, n_rows, n_cols, row_begin, row_end, t0,
                                         (typename OpT<OP>::out_t *)out, rowk, go);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

size_t xt_bytes(uint32_t rows, uint32_t kw) { return kam_align_up((size_t)kw * pad4(rows) * 4, 1024); }

// max element of x (device) -> host; synchronises `st`
int max_of(const void *x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *d_flag, uint32_t *h_max,
           cudaStream_t st) {
    KAM_CUDA(cudaMemsetAsync(d_flag, 0, 4, st));
    const int blocks = kam_sm_count() * 8;
    if (elem == KAM_U8) max_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t *)x, n, d, ld, d_flag);
    else if (elem == KAM_U16) max_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t *)x, n, d, ld, d_flag);
    else max_kernel<uint32_t><<<blocks, 256, 0, st>>>((const uint32_t *)x, n, d, ld, d_flag);
    KAM_CUDA(cudaGetLastError());
    KAM_CUDA(cudaMemcpyAsync(h_max, d_flag, 4, cudaMemcpyDeviceToHost, st));
    KAM_CUDA(cudaStreamSynchronize(st));
    return KAM_OK;
}

// manhat / info as tensor-core contractions over thermometer-coded rows (gram.cu: kam_thermo_*): 0 = never,
// 1 = when the shape fills tiles of 256 x 256 pairs and the largest count is small (default), 2 = whenever the
// operands can be encoded (tests: the small golden cases of the reference's machine code run through it too)
int g_tensor_mode = 1;
// the tensor path pays sum-of-column-maxima MACs per pair where the byte kernel pays d VABSDIFF4 lane-operations:
// it wins up to ~20 bytes per column on average; beyond THERMO_MAX_AVG (and always above a count of 255) the
// CUDA-core kernels run
constexpr uint32_t THERMO_MAX_AVG = 12;

inline bool thermo_shape_ok(uint32_t rows, uint32_t cols) {
    return g_tensor_mode == 2 || (g_tensor_mode == 1 && rows >= 512 && cols >= 128);
}
inline bool thermo_elem_ok(int elem) { return elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32; }

}  // namespace

// exact integer Gram metrics for inputs outside the fp16/fp32-exact envelope of gram.cu
size_t kam_int_gram_xt_bytes(uint32_t n, uint64_t d) { return xt_bytes(n, kw_of(OP_DOT, d)); }

int kam_int_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                     uint32_t row_end, const RowK *rk, GramOut go, uint32_t *xt, cudaStream_t st) {
    int rc = convert_dispatch<OP_DOT>(d_x, elem, n, d, ld, xt, st);
    if (rc) return rc;
    return launch_pairs<OP_DOT, true>(xt, n, xt, n, kw_of(OP_DOT, d), row_begin, row_end, nullptr, st, rk, go);
}

extern "C" {

static size_t thermo_need(uint32_t rows, uint64_t d, uint64_t avg_levels) {   // flag + column arrays + operand rows
    return 1024 + kam_thermo_cols_bytes(d) + kam_thermo_bytes(rows, kam_thermo_kpad(avg_levels * d));
}

size_t kam_manhattan_workspace_bytes(uint32_t rows_total, uint64_t d) {
    // worst case: one uint32 word per element, plus padding for two operands and the flag word
    // ... and never less than the tensor path needs while the column maxima sum to 4 d (thermometer rows of 4 d
    // bytes); a caller that wants it for heavier data passes kam_manhattan_workspace_bytes_levels(rows, d, average)
    const size_t classic = xt_bytes(rows_total, (uint32_t)d) + 2 * 4096 + 1024;
    const size_t tensor = thermo_need(rows_total, d, 4);
    return classic > tensor ? classic : tensor;
}

size_t kam_manhattan_workspace_bytes_levels(uint32_t rows_total, uint64_t d, uint32_t avg_levels) {
    const uint32_t levels = avg_levels < 1 ? 1 : avg_levels > THERMO_MAX_AVG ? THERMO_MAX_AVG : avg_levels;
    const size_t base = kam_manhattan_workspace_bytes(rows_total, d), tensor = thermo_need(rows_total, d, levels);
    return base > tensor ? base : tensor;
}

void kam_dist_set_tensor(int mode) { g_tensor_mode = mode < 0 ? 0 : mode > 2 ? 2 : mode; }

int kam_manhattan_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                      uint32_t row_end, uint32_t *d_tri, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_tri, "null pointer");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 2 || row_begin >= row_end) return KAM_OK;
    if (row_end > n) row_end = n;
    if (elem == KAM_U64 || elem == KAM_F32 || elem == KAM_F64) {
        uint32_t gy = (n + 127) / 128;
        if (gy > 65535) gy = 65535;
        dim3 grid(row_end - row_begin, gy);
        if (elem == KAM_U64) manhat_serial_kernel<uint64_t><<<grid, 128, 0, st>>>((const uint64_t *)d_x, n, d, ld, row_begin, row_end, d_tri);
        else if (elem == KAM_F32) manhat_serial_kernel<float><<<grid, 128, 0, st>>>((const float *)d_x, n, d, ld, row_begin, row_end, d_tri);
        else manhat_serial_kernel<double><<<grid, 128, 0, st>>>((const double *)d_x, n, d, ld, row_begin, row_end, d_tri);
        KAM_CUDA(cudaGetLastError());
        return KAM_OK;
    }
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32, "bad element type");
    KAM_REQUIRE(d_ws, "null workspace");
    if (ws_bytes < kam_manhattan_workspace_bytes(n, d)) {
        kam_set_error("kam_manhattan_tri: workspace too small");
        return KAM_E_WORKSPACE;
    }
    uint32_t *flag = (uint32_t *)d_ws;
    uint32_t *xt = (uint32_t *)((char *)d_ws + 1024);
    uint32_t mx = 0;
    int rc;
    if (thermo_shape_ok(n, n) && ws_bytes >= thermo_need(n, d, 1)) {
        // small counts: sum |a - b| = S_a + S_b - 2 sum min(a, b) on the tensor cores (thermometer rows, gram.cu).
        // The per-column maxima decide the layout (and give the largest count the other paths need).
        uint32_t *colmax = (uint32_t *)((char *)d_ws + 1024), *off = colmax + d;
        unsigned long long *stats = (unsigned long long *)(off + d + 2);
        unsigned long long hs[2] = {0, 0};
        KAM_CUDA(cudaMemsetAsync(colmax, 0, d * 4, st));
        if ((rc = kam_thermo_colmax(d_x, elem, n, d, ld, colmax, st)) || (rc = kam_thermo_layout(colmax, d, off, stats, hs, st)))
            return rc;
        mx = (uint32_t)(hs[0] > 0xffffffffull ? 0xffffffffull : hs[0]);
        const uint64_t kpad = kam_thermo_kpad(hs[1]);
        const size_t need = 1024 + kam_thermo_cols_bytes(d) + kam_thermo_bytes(n, kpad);
        if (mx <= 255 && hs[1] < (1ull << 31) && need <= ws_bytes && (g_tensor_mode == 2 || hs[1] <= (uint64_t)THERMO_MAX_AVG * d)) {
            ThermoOp op;
            if ((rc = kam_thermo_encode(d_x, elem, n, d, ld, off, kpad, (char *)d_ws + 1024 + kam_thermo_cols_bytes(d), &op, st)))
                return rc;
            return kam_thermo_pairs(op, op, true, row_begin, row_end, d_tri, nullptr, 0, st);
        }
    } else if ((rc = max_of(d_x, elem, n, d, ld, flag, &mx, st))) {
        return rc;
    }
    if (mx <= 255) {   // every count fits a byte: four elements per VABSDIFF4
        if ((rc = convert_dispatch<OP_SAD8X4>(d_x, elem, n, d, ld, xt, st))) return rc;
        return launch_pairs<OP_SAD8X4, true>(xt, n, xt, n, kw_of(OP_SAD8X4, d), row_begin, row_end, d_tri, st);
    }
    if ((rc = convert_dispatch<OP_SAD32>(d_x, elem, n, d, ld, xt, st))) return rc;
    return launch_pairs<OP_SAD32, true>(xt, n, xt, n, kw_of(OP_SAD32, d), row_begin, row_end, d_tri, st);
}

int kam_manhattan_rect(const void *d_x, const void *d_y, int elem, uint32_t m, uint32_t c, uint64_t d, uint64_t ldx,
                       uint64_t ldy, uint32_t *d_out, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_y && d_out && d_ws, "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32, "bad element type");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ldx >= d && ldy >= d, "bad d / ld");
    if (m == 0 || c == 0) return KAM_OK;
    if (ws_bytes < kam_manhattan_workspace_bytes(m, d) + kam_manhattan_workspace_bytes(c, d)) {
        kam_set_error("kam_manhattan_rect: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *flag = (uint32_t *)d_ws;
    uint32_t mx = 0, my = 0;
    int rc;
    if (thermo_shape_ok(m, c) && ws_bytes >= thermo_need(m, d, 1) + kam_thermo_bytes(c, kam_thermo_kpad(d))) {
        uint32_t *colmax = (uint32_t *)((char *)d_ws + 1024), *off = colmax + d;
        unsigned long long *stats = (unsigned long long *)(off + d + 2);
        unsigned long long hs[2] = {0, 0};
        KAM_CUDA(cudaMemsetAsync(colmax, 0, d * 4, st));
        if ((rc = kam_thermo_colmax(d_x, elem, m, d, ldx, colmax, st)) || (rc = kam_thermo_colmax(d_y, elem, c, d, ldy, colmax, st)) ||
            (rc = kam_thermo_layout(colmax, d, off, stats, hs, st)))
            return rc;
        mx = my = (uint32_t)(hs[0] > 0xffffffffull ? 0xffffffffull : hs[0]);
        const uint64_t kpad = kam_thermo_kpad(hs[1]);
        const size_t at_x = 1024 + kam_thermo_cols_bytes(d), at_y = at_x + kam_thermo_bytes(m, kpad);
        if (mx <= 255 && hs[1] < (1ull << 31) && at_y + kam_thermo_bytes(c, kpad) <= ws_bytes &&
            (g_tensor_mode == 2 || hs[1] <= (uint64_t)THERMO_MAX_AVG * d)) {
            ThermoOp ox, oy;
            if ((rc = kam_thermo_encode(d_x, elem, m, d, ldx, off, kpad, (char *)d_ws + at_x, &ox, st)) ||
                (rc = kam_thermo_encode(d_y, elem, c, d, ldy, off, kpad, (char *)d_ws + at_y, &oy, st)))
                return rc;
            return kam_thermo_pairs(ox, oy, false, 0, m, d_out, nullptr, c, st);
        }
    } else if ((rc = max_of(d_x, elem, m, d, ldx, flag, &mx, st)) || (rc = max_of(d_y, elem, c, d, ldy, flag, &my, st))) {
        return rc;
    }
    const bool bytes = mx <= 255 && my <= 255;
    const uint32_t kw = kw_of(bytes ? OP_SAD8X4 : OP_SAD32, d);
    uint32_t *xt = (uint32_t *)((char *)d_ws + 1024);
    uint32_t *yt = (uint32_t *)((char *)xt + xt_bytes(m, kw));
    if (bytes) {
        if ((rc = convert_dispatch<OP_SAD8X4>(d_x, elem, m, d, ldx, xt, st)) ||
            (rc = convert_dispatch<OP_SAD8X4>(d_y, elem, c, d, ldy, yt, st)))
            return rc;
        return launch_pairs<OP_SAD8X4, false>(xt, m, yt, c, kw, 0, m, d_out, st);
    }
    if ((rc = convert_dispatch<OP_SAD32>(d_x, elem, m, d, ldx, xt, st)) ||
        (rc = convert_dispatch<OP_SAD32>(d_y, elem, c, d, ldy, yt, st)))
        return rc;
    return launch_pairs<OP_SAD32, false>(xt, m, yt, c, kw, 0, m, d_out, st);
}

int kam_manhattan_rect_f32(const void *d_x, int elem_x, const float *d_y, uint32_t m, uint32_t c, uint64_t d,
                           uint64_t ldx, uint64_t ldy, float *d_out, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_y && d_out && d_ws, "null pointer");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ldx >= d && ldy >= d, "bad d / ld");
    if (m == 0 || c == 0) return KAM_OK;
    if (ws_bytes < kam_manhattan_workspace_bytes(m, d) + kam_manhattan_workspace_bytes(c, d)) {
        kam_set_error("kam_manhattan_rect_f32: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t kw = kw_of(OP_L1F32, d);
    uint32_t *xt = (uint32_t *)((char *)d_ws + 1024);
    uint32_t *yt = (uint32_t *)((char *)xt + xt_bytes(m, kw));
    int rc;
    if ((rc = convert_dispatch<OP_L1F32>(d_x, elem_x, m, d, ldx, xt, st)) ||
        (rc = convert_dispatch<OP_L1F32>(d_y, KAM_F32, c, d, ldy, yt, st)))
        return rc;
    return launch_pairs<OP_L1F32, false>(xt, m, yt, c, kw, 0, m, d_out, st);
}

static size_t info_classic_bytes(uint32_t n, uint64_t d) { return xt_bytes(n, kw_of(OP_JACC, d)) + 1024; }
size_t kam_info_workspace_bytes(uint32_t n, uint64_t d) {   // presence bits, or presence BYTES for the tensor path
    const size_t classic = info_classic_bytes(n, d), tensor = kam_thermo_bytes(n, kam_thermo_kpad(d)) + 1024;
    return classic > tensor ? classic : tensor;
}

int kam_info_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                 uint32_t row_end, float *d_tri, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_tri && d_ws, "null pointer");
    KAM_REQUIRE(elem >= KAM_U8 && elem <= KAM_F64, "bad element type");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    if (n < 2 || row_begin >= row_end) return KAM_OK;
    if (ws_bytes < info_classic_bytes(n, d)) {
        kam_set_error("kam_info_tri: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (row_end > n) row_end = n;
    // |supp_i & supp_j| is the Gram of the presence vectors: tensor cores (gram.cu: level 1 of the thermometer rows)
    if (thermo_elem_ok(elem) && thermo_shape_ok(n, n) && ws_bytes >= 1024 + kam_thermo_bytes(n, kam_thermo_kpad(d))) {
        ThermoOp op;
        int rc = kam_thermo_presence(d_x, elem, n, d, ld, (char *)d_ws + 1024, &op, st);
        if (rc) return rc;
        return kam_thermo_pairs(op, op, true, row_begin, row_end, nullptr, d_tri, 0, st);
    }
    uint32_t *xt = (uint32_t *)((char *)d_ws + 1024);
    int rc = convert_dispatch<OP_JACC>(d_x, elem, n, d, ld, xt, st);
    if (rc) return rc;
    return launch_pairs<OP_JACC, true>(xt, n, xt, n, kw_of(OP_JACC, d), row_begin, row_end, d_tri, st);
}

}  // extern "C"
