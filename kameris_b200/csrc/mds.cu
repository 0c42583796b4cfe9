// mds.cu -- the n^2 parts of the `mds` job step (kameris/job_steps/mds.py:10-35: classical
// multidimensional scaling of a distance matrix) as HBM-bound float64 kernels:
//
//   kam_mds_center   packed `.mm-dist` triangle (uint32 manhat, float32 others) ->
//                    B = -0.5 (D^2 - colmean_j - rowmean_i + mean), n x n float64 (mds.py:14-24, same
//                    operation order), in three sweeps: expand+square (each 32x32 tile of the triangle
//                    is read once, coalesced, and written to both (i,j) and, transposed through shared
//                    memory, (j,i)), row sums, centring in place
//   kam_symm_block   Y = B Q for the symmetric B and a block of 16 column vectors: the matrix-block
//                    product of the subspace iteration that replaces scipy's eigsh(B, k=dim) (mds.py:26).
//                    B is streamed from HBM exactly once per call (n^2 x 8 bytes -- the roofline);
//                    the 16-column block is staged through shared memory in 128-row slabs and shared by
//                    the 16 rows of a CTA; each warp owns two rows of B, lanes stride the columns, 8
//                    independent 256-byte row loads are in flight per warp.
#include "common.cuh"
#include "tma.cuh"

namespace {

constexpr int MT = 32;                     // tile edge of the expand kernel
constexpr int SB = 16;                     // columns of the vector block
constexpr int SY_WARPS = 8, SY_ROWS = 4;   // rows of B per warp
constexpr int SY_SLAB = 128;               // rows of Q staged per step (4 chunks of 32 columns of B)

template <typename T>
__global__ void __launch_bounds__(MT * 8) mds_expand_sq_kernel(const T *__restrict__ tri, uint32_t n,
                                                               double *__restrict__ B) {
    __shared__ double tile[MT][MT + 1];
    // blockIdx.x enumerates the tile pairs (bi <= bj) of the upper triangle row by row
    const uint32_t nt = (n + MT - 1) / MT;
    uint32_t bi = 0, rem = blockIdx.x;
    while (rem >= nt - bi) {               // at most nt steps; nt is small next to the tile work
        rem -= nt - bi;
        ++bi;
    }
    const uint32_t bj = bi + rem;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint32_t r = ty; r < MT; r += 8) {
        const uint64_t i = (uint64_t)bi * MT + r, j = (uint64_t)bj * MT + tx;
        double v = 0.0;
        if (i < n && j < n && j > i) {
            const double d = (double)tri[(uint64_t)n * i - i * (i + 3) / 2 + j - 1];
            v = d * d;
        } else if (i < n && j < n && j < i) {                      // only inside diagonal tiles
            const double d = (double)tri[(uint64_t)n * j - j * (j + 3) / 2 + i - 1];
            v = d * d;
        }
        tile[r][tx] = v;
        if (i < n && j < n) B[i * n + j] = v;
    }
    if (bi == bj) return;                                          // diagonal tile: both halves written above
    __syncthreads();
    for (uint32_t r = ty; r < MT; r += 8) {
        const uint64_t j = (uint64_t)bj * MT + r, i = (uint64_t)bi * MT + tx;   // row j of B, column i
        if (i < n && j < n) B[j * n + i] = tile[tx][r];
    }
}

// tot[i] = (sum of row i) / n  (== numpy's column means: B is symmetric).  One warp per row.
__global__ void __launch_bounds__(256) mds_rowmean_kernel(const double *__restrict__ B, uint32_t n,
                                                          double *__restrict__ tot) {
    const uint32_t row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *r = B + (uint64_t)row * n;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t j = lane;
    for (; j + 96 < n; j += 128) {
        s0 += r[j];
        s1 += r[j + 32];
        s2 += r[j + 64];
        s3 += r[j + 96];
    }
    for (; j < n; j += 32) s0 += r[j];
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tot[row] = s / (double)n;
}

// mean[0] = (sum of tot) / n; one CTA
__global__ void __launch_bounds__(1024) mds_mean_kernel(const double *__restrict__ tot, uint32_t n,
                                                        double *__restrict__ mean) {
    __shared__ double red[32];
    double s = 0;
    for (uint32_t i = threadIdx.x; i < n; i += 1024) s += tot[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) mean[0] = s / (double)n;
    }
}

// B_ij = ((B_ij - tot_j) - tot_i + mean) * -0.5, the operation order of mds.py:19-24
__global__ void __launch_bounds__(256) mds_center_kernel(double *__restrict__ B, uint32_t n,
                                                         const double *__restrict__ tot,
                                                         const double *__restrict__ mean) {
    const uint32_t i = blockIdx.x;                       // rows in gridDim.x: n may exceed the 65535 of gridDim.y
    const double ti = tot[i], m = mean[0];
    double *r = B + (uint64_t)i * n;
    for (uint32_t j = blockIdx.y * 256 + threadIdx.x; j < n; j += gridDim.y * 256)
        r[j] = ((r[j] - tot[j]) - ti + m) * -0.5;
}

// Ypart[split][n x 16] = B[:, j range of the split] * Q[j range, :].  CTA = 8 warps x 4 rows of B, lanes
// stride the columns j; grid = (row blocks, column splits) so that the number of work items fills the SMs
// evenly (n = 10 000: 313 row blocks alone would be 2.1 waves).  Software pipeline over slabs of 128
// columns: the 16 row loads of slab s+1 (4 rows x 4 chunks, independent 256-byte warp requests) and the
// thread's share of the next 128 x 16 slab of Q are issued BEFORE the multiply of slab s, so HBM/L2
// latency hides behind 256 DFMA per lane.  Q slabs live in shared memory with a row stride of 18 doubles
// (the 16-byte reads of a quarter-warp fall into distinct banks), double-buffered: one barrier per slab.
constexpr int SY_QLD = SB + 2;
constexpr int SY_QPT = SY_SLAB * SB / (SY_WARPS * 32);   // Q elements staged per thread and slab (8)

__global__ void __launch_bounds__(SY_WARPS * 32, 1) symm_block_kernel(const double *__restrict__ B, uint32_t n,
                                                                      const double *__restrict__ Q,
                                                                      double *__restrict__ Ypart,
                                                                      uint32_t cols_per_split) {
    __shared__ __align__(16) double qs[2][SY_SLAB][SY_QLD];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row0 = (blockIdx.x * SY_WARPS + warp) * SY_ROWS;
    const uint32_t jbeg = blockIdx.y * cols_per_split;
    const uint32_t jend = jbeg + cols_per_split < n ? jbeg + cols_per_split : n;
    const double *rp[SY_ROWS];
#pragma unroll
    for (int r = 0; r < SY_ROWS; ++r) rp[r] = B + (uint64_t)(row0 + r < n ? row0 + r : 0) * n;
    double acc[SY_ROWS][SB];
#pragma unroll
    for (int r = 0; r < SY_ROWS; ++r)
#pragma unroll
        for (int c = 0; c < SB; ++c) acc[r][c] = 0.0;

    double bv[SY_ROWS][SY_SLAB / 32], qreg[SY_QPT];
    auto load_slab = [&](uint32_t j0) {
#pragma unroll
        for (int q = 0; q < SY_SLAB / 32; ++q) {
            const uint32_t j = j0 + q * 32 + lane;
#pragma unroll
            for (int r = 0; r < SY_ROWS; ++r) bv[r][q] = j < jend ? __ldg(rp[r] + j) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < SY_QPT; ++i) {
            const uint32_t e = threadIdx.x + i * SY_WARPS * 32, jj = e / SB;
            qreg[i] = j0 + jj < jend ? __ldg(Q + (uint64_t)(j0 + jj) * SB + e % SB) : 0.0;
        }
    };
    auto store_q = [&](uint32_t buf) {
#pragma unroll
        for (int i = 0; i < SY_QPT; ++i) {
            const uint32_t e = threadIdx.x + i * SY_WARPS * 32;
            qs[buf][e / SB][e % SB] = qreg[i];
        }
    };

    load_slab(jbeg);
    store_q(0);
    __syncthreads();
    uint32_t buf = 0;
    for (uint32_t j0 = jbeg; j0 < jend; j0 += SY_SLAB, buf ^= 1u) {
        double bc[SY_ROWS][SY_SLAB / 32];
#pragma unroll
        for (int r = 0; r < SY_ROWS; ++r)
#pragma unroll
            for (int q = 0; q < SY_SLAB / 32; ++q) bc[r][q] = bv[r][q];
        const bool more = j0 + SY_SLAB < jend;
        if (more) load_slab(j0 + SY_SLAB);                     // in flight during the multiply below
#pragma unroll
        for (int q = 0; q < SY_SLAB / 32; ++q) {
            const double2 *qrow = reinterpret_cast<const double2 *>(&qs[buf][q * 32 + lane][0]);
#pragma unroll
            for (int c2 = 0; c2 < SB / 2; ++c2) {
                const double2 qv = qrow[c2];
#pragma unroll
                for (int r = 0; r < SY_ROWS; ++r) {
                    acc[r][2 * c2] = fma(bc[r][q], qv.x, acc[r][2 * c2]);
                    acc[r][2 * c2 + 1] = fma(bc[r][q], qv.y, acc[r][2 * c2 + 1]);
                }
            }
        }
        if (more) store_q(buf ^ 1u);                           // that buffer was last read one slab ago
        __syncthreads();
    }
    double *Yp = Ypart + (uint64_t)blockIdx.y * n * SB;
#pragma unroll
    for (int r = 0; r < SY_ROWS; ++r) {
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < SB; ++c) {
            double v = acc[r][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((int)lane == c) mine = v;
        }
        if (lane < SB && row0 + r < n) Yp[(uint64_t)(row0 + r) * SB + lane] = mine;
    }
}


// ---- the same product with TMA-fed shared-memory tiles ---------------------------------------------------
// symm_block_kernel keeps its operands in flight in REGISTERS (16 row loads per lane), and its 64
// replicated accumulators per lane leave room for no more: ~32 KB in flight per SM, a third of what
// HBM latency x bandwidth asks for.  Here one elected thread streams B tiles (32 rows x 128 columns,
// 32 KB) and the matching Q slab (128 x 16, 16 KB, SWIZZLE_128B so that the eight 16-byte chunks a
// quarter-warp reads from eight consecutive rows fall into distinct banks) through a 4-stage
// cp.async.bulk.tensor + mbarrier ring: 128 KB of B in flight per SM, independent of the register
// file; out-of-range rows / columns are zero-filled by the TMA unit, so the body has no bounds checks.
// The eight consumer warps keep the lanes-stride-columns mapping (conflict-free, coalesced reads of B).
constexpr int TY_STAGES = 4;
constexpr int TY_ROWS = SY_WARPS * SY_ROWS;                                  // 32 rows of B per CTA
constexpr uint32_t TY_B_BYTES = TY_ROWS * SY_SLAB * 8, TY_Q_BYTES = SY_SLAB * SB * 8;
constexpr uint32_t TY_STAGE_BYTES = TY_B_BYTES + TY_Q_BYTES;                 // 48 KB

__global__ void __launch_bounds__(SY_WARPS * 32 + 32, 1) symm_tma_kernel(const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmQ,
                                                                         uint32_t n, double *__restrict__ Ypart,
                                                                         uint32_t cols_per_split) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[TY_STAGES], empty[TY_STAGES];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t jbeg = blockIdx.y * cols_per_split;
    const uint32_t jend = jbeg + cols_per_split < n ? jbeg + cols_per_split : n;
    const uint32_t n_slabs = jbeg < jend ? (jend - jbeg + SY_SLAB - 1) / SY_SLAB : 0;
    const uint32_t row_base = blockIdx.x * TY_ROWS;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmQ);
        for (int s = 0; s < TY_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], SY_WARPS);   // one arrive per consumer warp
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == SY_WARPS) {
        // ===== producer: one thread =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t sl = 0; sl < n_slabs; ++sl) {
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&full[stage], TY_STAGE_BYTES);
                unsigned char *sb = smem + stage * TY_STAGE_BYTES;
                const int32_t j0 = (int32_t)(jbeg + sl * SY_SLAB);
                tma_load_2d(sb, &tmB, j0, (int32_t)row_base, &full[stage]);          // [32 rows][128 cols]
                tma_load_2d(sb + TY_B_BYTES, &tmQ, 0, j0, &full[stage]);             // [128 rows][16 cols], swizzled
                if (++stage == TY_STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }

    // ===== consumers: warp = (row group g of 8 rows, column half h of 8 columns); lanes stride the
    // columns j.  A lane needs R values of B and C values of Q per column for R x C FMAs, i.e.
    // shared-memory wavefronts per FMA ~ 1/(16R) + 1/(16C): the square 8 x 8 block needs 20 % fewer than 4 x 16.
    constexpr int TR = 8, TC = 8;
    const uint32_t g = warp >> 1, h = warp & 1u;
    double acc[TR][TC];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[r][c] = 0.0;
    uint32_t stage = 0, phase = 0;
    for (uint32_t sl = 0; sl < n_slabs; ++sl) {
        mbar_wait(&full[stage], phase);
        const unsigned char *sb = smem + stage * TY_STAGE_BYTES;
        const double *bs = reinterpret_cast<const double *>(sb) + (g * TR) * SY_SLAB;
        // columns of this slab beyond jend belong to the next split (the TMA box does not stop there)
        const uint32_t valid = jend - (jbeg + sl * SY_SLAB);
#pragma unroll
        for (int q = 0; q < SY_SLAB / 32; ++q) {
            const uint32_t jj = q * 32 + lane;
            double bv[TR];
#pragma unroll
            for (int r = 0; r < TR; ++r) bv[r] = jj < valid ? bs[r * SY_SLAB + jj] : 0.0;
            const unsigned char *qrow = sb + TY_B_BYTES + jj * 128u;
#pragma unroll
            for (int c2 = 0; c2 < TC / 2; ++c2) {
                const uint32_t chunk = h * (TC / 2) + c2;
                const double2 qv = *reinterpret_cast<const double2 *>(qrow + ((chunk ^ (jj & 7u)) << 4));
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    acc[r][2 * c2] = fma(bv[r], qv.x, acc[r][2 * c2]);
                    acc[r][2 * c2 + 1] = fma(bv[r], qv.y, acc[r][2 * c2 + 1]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == TY_STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
    const uint32_t row0 = row_base + g * TR;
    double *Yp = Ypart + (uint64_t)blockIdx.y * n * SB;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < TC; ++c) {
            double v = acc[r][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((int)lane == c) mine = v;
        }
        if (lane < TC && row0 + r < n) Yp[(uint64_t)(row0 + r) * SB + h * TC + lane] = mine;
    }
}

bool g_symm_tma = true;   // test hook: false = register-pipelined kernel for every n

// Y = sum over the column splits (fixed order: deterministic)
__global__ void __launch_bounds__(256) symm_reduce_kernel(const double *__restrict__ Ypart, uint32_t n, uint32_t splits,
                                                          double *__restrict__ Y) {
    const uint64_t e = (uint64_t)blockIdx.x * 256 + threadIdx.x, tot = (uint64_t)n * SB;
    if (e >= tot) return;
    double s = Ypart[e];
    for (uint32_t k = 1; k < splits; ++k) s += Ypart[(uint64_t)k * tot + e];
    Y[e] = s;
}

// number of column splits: fill whole waves of SMs (one CTA per SM), at least 4 slabs per split
inline uint32_t symm_splits(uint32_t n) {
    const uint32_t rb = (n + SY_WARPS * SY_ROWS - 1) / (SY_WARPS * SY_ROWS), sms = (uint32_t)kam_sm_count();
    uint32_t best = 1;
    double best_eff = 0.0;
    for (uint32_t s = 1; s <= 8; ++s) {
        if (s > 1 && n / s < 4 * SY_SLAB) break;
        const double waves = (double)rb * s / sms;
        const double eff = waves / (double)(uint64_t)(waves + 0.999999);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}

}  // namespace

extern "C" {

size_t kam_mds_workspace_bytes(uint32_t n) { return kam_align_up((size_t)n * 8, 256) + 256; }

int kam_mds_center(const void *d_tri, int elem, uint32_t n, double *d_B, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_B && d_ws && (d_tri || n < 2), "null pointer");
    KAM_REQUIRE(elem == KAM_U32 || elem == KAM_F32, "a .mm-dist triangle holds uint32 (manhat) or float32 values");
    KAM_REQUIRE(n >= 1, "empty matrix");
    if (ws_bytes < kam_mds_workspace_bytes(n)) {
        kam_set_error("kam_mds_center: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double *tot = (double *)d_ws;
    double *mean = (double *)((char *)d_ws + kam_align_up((size_t)n * 8, 256));
    const uint64_t nt = (n + MT - 1) / MT, pairs = nt * (nt + 1) / 2;
    KAM_REQUIRE(pairs < (1ull << 31), "matrix too large for one call");
    if (elem == KAM_U32)
        mds_expand_sq_kernel<uint32_t><<<(unsigned)pairs, MT * 8, 0, st>>>((const uint32_t *)d_tri, n, d_B);
    else
        mds_expand_sq_kernel<float><<<(unsigned)pairs, MT * 8, 0, st>>>((const float *)d_tri, n, d_B);
    mds_rowmean_kernel<<<(n + 7) / 8, 256, 0, st>>>(d_B, n, tot);
    mds_mean_kernel<<<1, 1024, 0, st>>>(tot, n, mean);
    unsigned gx = (n + 255) / 256;
    if (gx > 64) gx = 64;
    mds_center_kernel<<<dim3(n, gx), 256, 0, st>>>(d_B, n, tot, mean);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

size_t kam_symm_workspace_bytes(uint32_t n) { return (size_t)symm_splits(n) * n * SB * sizeof(double) + 256; }

int kam_symm_block(const double *d_B, uint32_t n, const double *d_Q, double *d_Y, void *d_ws, size_t ws_bytes,
                   void *stream) {
    KAM_REQUIRE(d_B && d_Q && d_Y && d_ws, "null pointer");
    KAM_REQUIRE(n >= 1, "empty matrix");
    if (ws_bytes < kam_symm_workspace_bytes(n)) {
        kam_set_error("kam_symm_block: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t splits = symm_splits(n);
    const uint32_t cps = ((n + splits - 1) / splits + SY_SLAB - 1) / SY_SLAB * SY_SLAB;   // whole slabs per split
    const unsigned rb = (n + SY_WARPS * SY_ROWS - 1) / (SY_WARPS * SY_ROWS);
    double *part = splits == 1 ? d_Y : (double *)d_ws;
    // TMA needs 16-byte-aligned rows (n even) and bases; everything else takes the register-pipelined kernel
    const bool tma_ok = g_symm_tma && n % 2 == 0 && n >= 64 && ((uintptr_t)d_B & 15) == 0 && ((uintptr_t)d_Q & 15) == 0;
    if (tma_ok) {
        CUtensorMap tmB, tmQ;
        int rc;
        if ((rc = kam_make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, d_B, n, n, (uint64_t)n * 8, SY_SLAB, TY_ROWS,
                                   CU_TENSOR_MAP_SWIZZLE_NONE)) ||
            (rc = kam_make_tmap_2d(&tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, d_Q, SB, n, (uint64_t)SB * 8, SB, SY_SLAB,
                                   CU_TENSOR_MAP_SWIZZLE_128B)))
            return rc;
        const size_t smem = (size_t)TY_STAGES * TY_STAGE_BYTES + 1024;
        KAM_CUDA(cudaFuncSetAttribute(symm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        symm_tma_kernel<<<dim3(rb, splits), SY_WARPS * 32 + 32, smem, st>>>(tmB, tmQ, n, part, cps);
    } else {
        symm_block_kernel<<<dim3(rb, splits), SY_WARPS * 32, 0, st>>>(d_B, n, d_Q, part, cps);
    }
    if (splits > 1) symm_reduce_kernel<<<(unsigned)(((uint64_t)n * SB + 255) / 256), 256, 0, st>>>(part, n, splits, d_Y);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

/* test hook: 0 = always the register-pipelined product kernel (default 1: TMA-fed tiles when n is even) */
void kam_symm_set_tma(int on) { g_symm_tma = on != 0; }

}  // extern "C"
