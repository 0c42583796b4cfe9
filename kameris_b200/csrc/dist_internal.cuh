// dist_internal.cuh -- shared between gram.cu (tensor-core Gram) and dist_tile.cu (exact integer
// Gram on the CUDA cores): per-row constants and the float64 epilogue of the Gram metrics.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// per-row constants derived from S = sum of the row and G = sum of squares (exact integers)
struct RowK {
    double rn;   // 1/sqrt(G)                       (cosine)
    double S;    // S
    double mu;   // S/D
    double rp;   // 1/sqrt(G - S^2/D)               (pearson)
    double a;    // G (counts) or G/S^2 (frequencies)   (euclid)
    double q;    // 1 (counts) or 1/S   (frequencies)
};

struct GramOut {
    float *euclid, *cosine, *pearson;   // nullptr = not requested
    int *raw = nullptr;                 // when set: the exact int32 Gram entry itself, dense [row * ld + col]
                                        // (limb-split path: three partial Grams are combined afterwards)
    // thermometer operands (gram.cu: kam_thermo_*): G_ij = sum_e min(c_i[e], c_j[e]) resp. |supp_i & supp_j|;
    // the RowK fields `a` / `q` then hold the BIT PATTERNS of the integers sum(c) / nnz(c), not doubles
    uint32_t *manhat = nullptr;         // sum(c_i) + sum(c_j) - 2 G_ij  (mod 2^32, like the reference's uint32 sum)
    float *info = nullptr;              // U = nnz_i + nnz_j - G_ij: U == 0 ? 0 : (1.0f / (float)U) * (float)(U - G_ij)
};

// geometry of one Gram block: rows of matrix A against rows of matrix B
struct GramGeom {
    uint32_t n_a, n_b;            // rows of A (block rows) and of B (block columns)
    uint32_t row_begin, row_end;  // the A rows this launch writes
    uint32_t tri;                 // 1: A and B are the same matrix, only pairs j > i exist
    unsigned long long ld;        // 0: packed strict upper triangle of n_a (tri only); else dense row stride
    const RowK *rk_a, *rk_b;      // per-row constants of A's and B's rows
    // fused exchange (kam_gram_fused): B's rows arrive over NVLink while the kernel runs.  ready[t / ready_tiles]
    // becomes non-zero once the rows of column tile t (256 rows of B) are in place; nullptr = all there.
    const uint32_t *ready = nullptr;
    uint32_t ready_tiles = 1;
    unsigned long long *wait_probe = nullptr;   // diagnostics: [0] += cycles producers spent waiting, [1] = max of one wait
};

#ifdef __CUDACC__
// g = exact G_ij.  Definitions: scipy.spatial.distance {euclidean, cosine, correlation}.
__device__ __forceinline__ void gram_store(const GramOut &go, long long idx, double g, const RowK &ri, const RowK &cj) {
    if (go.cosine) go.cosine[idx] = (float)(1.0 - g * ri.rn * cj.rn);
    if (go.pearson) go.pearson[idx] = (float)(1.0 - (g - ri.S * cj.mu) * ri.rp * cj.rp);
    if (go.euclid) {
        const double e2 = ri.a + cj.a - 2.0 * g * ri.q * cj.q;
        go.euclid[idx] = sqrtf((float)(e2 < 0.0 ? 0.0 : e2));   // cancellation below zero -> 0; NaN (an empty row) stays NaN
    }
}
#endif

#ifdef __CUDACC__
// 16 counts -> 16 tensor-core operand elements (uint8, saturating, or integer-valued fp16), one or two
// 16-byte stores
template <typename OP>
__device__ __forceinline__ void store_operand16(OP *dst, const unsigned (&c)[16]) {
    if constexpr (sizeof(OP) == 1) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = min(c[4 * i], 255u) | (min(c[4 * i + 1], 255u) << 8) | (min(c[4 * i + 2], 255u) << 16) |
                   (min(c[4 * i + 3], 255u) << 24);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        __half2 h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = __halves2half2(__uint2half_rn(c[2 * i]), __uint2half_rn(c[2 * i + 1]));
        reinterpret_cast<uint4 *>(dst)[0] = *reinterpret_cast<uint4 *>(&h[0]);
        reinterpret_cast<uint4 *>(dst)[1] = *reinterpret_cast<uint4 *>(&h[4]);
    }
}
#endif

// floating-point rows (freq_gram.cu): operand pre-pass with integer recovery (flags[2] counts the rows
// that are not frequency vectors), and the float64 CUDA-core Gram between the rows as stored
int kam_float_prepass(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint64_t dpad, bool i8, void *h,
                      unsigned long long *S, unsigned long long *G, unsigned long long *flags, cudaStream_t st);
int kam_float_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                       uint32_t row_end, RowK *rk, GramOut go, cudaStream_t st);

// manhat / info on the tensor cores (gram.cu): thermometer-coded operand rows of one matrix.  Column e of the
// input owns the K positions [off[e], off[e + 1]) of an operand row; byte t of them is (c[e] > t).  off is the
// running sum of the per-column maxima over BOTH matrices of a product, so a column costs as many MACs as its
// largest count anywhere (a few hot k-mers do not inflate the other 4^k columns); info uses one byte per column.
struct ThermoOp {
    const uint8_t *op = nullptr;   // [n][kpad]
    const RowK *rk = nullptr;      // a = bits of sum(c), q = bits of nnz(c)
    uint32_t n = 0;
    uint64_t d = 0, kpad = 0;      // kpad: row length in bytes (a multiple of 128)
};
size_t kam_thermo_cols_bytes(uint64_t d);                       // colmax[d] + off[d + 1] + 2 statistics words
size_t kam_thermo_bytes(uint32_t n, uint64_t kpad);             // RowK[n] + operand rows
uint64_t kam_thermo_kpad(uint64_t k_total);
// colmax[e] = max(colmax[e], max_r x[r][e])  (the caller zeroes colmax)
int kam_thermo_colmax(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *colmax, cudaStream_t st);
// off = exclusive running sum of colmax (d + 1 entries); h_stats = {largest count, sum of the maxima}; synchronises
int kam_thermo_layout(const uint32_t *colmax, uint64_t d, uint32_t *off, unsigned long long *d_stats,
                      unsigned long long (&h_stats)[2], cudaStream_t st);
int kam_thermo_encode(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *off,
                      uint64_t kpad, void *d_ws, ThermoOp *out, cudaStream_t st);
// info: presence bytes, one per column (kpad = d rounded up to 128)
int kam_thermo_presence(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, void *d_ws, ThermoOp *out,
                        cudaStream_t st);
// rows [row_begin, row_end) of A against all rows of B (tri: A and B are the same operand, pairs j > i, packed
// triangle when out_ld == 0); exactly one of d_manhat / d_info
int kam_thermo_pairs(const ThermoOp &A, const ThermoOp &B, bool tri, uint32_t row_begin, uint32_t row_end,
                     uint32_t *d_manhat, float *d_info, uint64_t out_ld, cudaStream_t st);

size_t kam_int_gram_xt_bytes(uint32_t n, uint64_t d);
int kam_int_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin,
                     uint32_t row_end, const RowK *rk, GramOut go, uint32_t *xt, cudaStream_t st);
