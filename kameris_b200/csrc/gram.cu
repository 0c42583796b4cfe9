// gram.cu -- euclid / cosine / pearson distance triangles from ONE Gram-matrix pass G = X X^T on
// the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// These three metrics do not exist in the reference (kameris/schemas/job_options.json:97 lists
// only manhat and info; usage string of generation_dists @0x100027d20); BASELINE.json's north star
// adds them.  They are defined as scipy.spatial.distance {euclidean, cosine, correlation}
//  and derived from the integer Gram matrix of the COUNT vectors:
//     cosine  = 1 - G_ij / sqrt(G_ii G_jj)
//     pearson = 1 - (G_ij - S_i S_j / D) / sqrt((G_ii - S_i^2/D)(G_jj - S_j^2/D))
//     euclid  = sqrt(G_ii + G_jj - 2 G_ij)                      on counts
//             = sqrt(G_ii/S_i^2 + G_jj/S_j^2 - 2 G_ij/(S_i S_j)) on frequencies x/sum(x)
// with S_i = sum of row i, D = number of bins.
//
// Exactness: every partial sum of a Gram entry is a non-negative integer bounded by max_i G_ii.
//   kind::i8  : counts <= 255 are exact uint8 operands, int32 accumulation is exact while
//               max G_ii < 2^31 (the usual case for k-mer counts: half the operand bytes and twice
//               the MMA rate of fp16);
//   kind::f16 : counts <= 2048 are exact in fp16, fp32 accumulation is exact while max G_ii < 2^24.
// Inside its envelope each path produces the exact integer Gram matrix; the only rounding is the
// float64 epilogue + the final float32 store.  Inputs outside both envelopes take an exact integer
// CUDA-core Gram kernel instead (dist_tile.cu, OP_DOT).  The pre-pass measures the envelope.
//
// Kernel (gram_tcgen05_kernel<I8, CL>): persistent, one CTA per SM in clusters of CL=2, 192 threads:
//   warp 0    TMA producer   : 128-row A box + its half of the 256-row B panel (multicast to both
//                              CTAs of the cluster), 128 bytes of K per row, SWIZZLE_128B, 4-stage ring
//   warp 1    MMA issuer     : tcgen05.mma.cta_group::1.kind::{i8,f16}, M=128 N=256 K=32 bytes, int32 /
//                              fp32 accumulators in TMEM, two 256-column accumulators so the epilogue
//                              of tile t overlaps the MMAs of tile t+1; tcgen05.commit (multicast)
//                              releases smem stages in both CTAs / publishes TMEM
//   warps 2-9 epilogue       : tcgen05.ld 32x32b -> float64 distance formulas -> shared-memory transpose ->
//                              row-contiguous stores into the packed triangle
// Only tiles that contain a pair i<j are scheduled (upper triangle).
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "dist_internal.cuh"
#include "tma.cuh"

namespace {

constexpr int BM = 128, BN = 256;
constexpr int BK_BYTES = 128;                        // one SWIZZLE_128B row of K per stage: 64 fp16 or 128 uint8
constexpr int G_STAGES = 4;
constexpr int EP_WARPS = 8;                           // epilogue warps: two per TMEM lane quarter
constexpr int G_THREADS = 64 + 32 * EP_WARPS;
constexpr uint32_t A_BYTES = BM * BK_BYTES, B_BYTES = BN * BK_BYTES, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;
// exactness envelopes (see the header comment)
constexpr uint32_t MAX_I8_COUNT = 255;                       // uint8 operands, int32 accumulation
constexpr unsigned long long MAX_I8_GII = 1ull << 31;
constexpr uint32_t MAX_F16_COUNT = 2048;                     // fp16 holds integers up to 2048 exactly
constexpr unsigned long long MAX_F16_GII = 1ull << 24;       // fp32 holds integers up to 2^24 exactly

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format at bits 4-5 (1 = F32, 2 = S32),
// a/b format at bits 7-9 / 10-12 (kind::f16: 0 = F16; kind::i8: 0 = unsigned 8-bit), both K-major
// (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_I8 = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, SWIZZLE_128B:
// start>>4 | LBO(16 B)>>4 at bit 16 | SBO(1024 B: next 8-row group)>>4 at bit 32 | version 1 at
// bit 46 | layout type 2 (SWIZZLE_128B) at bit 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// D[tmem] (+)= A[smem] * B[smem]^T : M128 x N256 x K(32 bytes) per instruction, issued by ONE thread
template <bool I8>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    if constexpr (I8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC_I8), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC_F16), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// commit that arrives on the same barrier offset in every CTA of cta_mask (cluster multicast)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Epilogue stores.  A TMEM lane is a ROW of the tile: after tcgen05.ld every lane of an epilogue warp holds 32
// consecutive columns of its own row, so storing straight from the registers makes one warp-wide store touch
// 32 different rows -- 32 separate 32-byte sectors for 128 bytes of payload.  At D = 4096 (32 K-steps per tile)
// that, not the MMAs, bounded the kernel (14.4 ms for 5.4e8 pairs x 3 metrics; cosine alone took 4.3 ms).  So
// each 32 x 32 chunk goes through a padded shared-memory tile per warp and leaves as 32 row-contiguous
// 128-byte stores.
constexpr int EP_PITCH = 33;
// Which flag a tile waits for is decided per WAVE of the static schedule (tile_list_segments): all CTA pairs of a
// wave wait for the same chunk -- the last one any tile of the wave (or of an earlier wave) needs -- and so resume
// together.  Pairs that waited for their own tile's chunk only drifted apart in their K loops by the length of
// their different waits, and the ~18 operand panels the tiles in flight share no longer met in L2: measured at 8
// ranks, 0.14 ms of waiting per CTA cost 5 ms of a 13 ms launch (profiles/r2_ring_fused_8gpu.jsonl).
// fused exchange: the TMA producer of a tile whose B rows are still on their way (copy engines pull them from a
// peer's memory and write ready[chunk] behind each chunk, peer.cu: kam_peer_pull) waits here.  The flag is read
// with acquire semantics at system scope; the rows were written by a copy engine and are read next by the TMA
// unit (async proxy), hence the proxy fence.  A flag that never comes (a peer died) traps after ~10 s instead of
// hanging the device.
__device__ __forceinline__ void wait_rows_ready(const uint32_t *flag, unsigned long long *probe) {
    uint32_t v, spin = 0;
    const long long t0 = probe ? clock64() : 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v) break;
        __nanosleep(256);
        if (++spin > (1u << 25)) __trap();
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    if (probe && spin) {
        const unsigned long long dt = (unsigned long long)(clock64() - t0);
        atomicAdd(probe, dt);
        atomicMax(probe + 1, dt);
        atomicAdd(probe + 2, 1ull);
    }
}

constexpr int EP_ROWS = 16;                                            // rows per staging round (half a warp's 32)
constexpr uint32_t EP_TILE_BYTES = 8 * EP_ROWS * EP_PITCH * 4;         // eight epilogue warps

template <bool I8>
__device__ __forceinline__ void gram_store_chunk(const GramGeom &gg, const GramOut &go, float *stile,
                                                 const uint32_t (&v)[32], const RowK &ri, const RowK *colk, uint32_t cb,
                                                 uint32_t row_first, uint32_t lane) {
    const uint32_t j = cb + lane;                       // the column this lane stores
    const bool j_ok = j < gg.n_b;
    float *mine = stile + (lane & (EP_ROWS - 1)) * EP_PITCH;
    // the lanes' 32 values each -> global memory, 16 rows at a time through the warp's staging tile
    auto put = [&](float *dst, const float (&val)[32]) {
#pragma unroll
        for (uint32_t h = 0; h < 32 / EP_ROWS; ++h) {
            if (lane / EP_ROWS == h) {
#pragma unroll
                for (int c = 0; c < 32; ++c) mine[c] = val[c];
            }
            __syncwarp();
            // rows of this round that this launch writes; the address of row r+1 follows from that of row r
            const uint32_t first = row_first + h * EP_ROWS;
            const uint32_t lo = first > gg.row_begin ? first : gg.row_begin;
            uint32_t hi = first + EP_ROWS < gg.row_end ? first + EP_ROWS : gg.row_end;
            hi = hi < gg.n_a ? hi : gg.n_a;
            if (lo < hi && j_ok) {
                const long long base = gg.ld ? (long long)lo * (long long)gg.ld
                                             : (long long)gg.n_a * lo - (long long)lo * (lo + 3) / 2 - 1;
                float *p = dst + base + j;
                const float *sp = stile + (lo - first) * EP_PITCH + lane;
#pragma unroll 8
                for (uint32_t row = lo; row < hi; ++row) {
                    if (!gg.tri || j > row) *p = *sp;
                    p += gg.ld ? (long long)gg.ld : (long long)gg.n_a - (long long)row - 2;
                    sp += EP_PITCH;
                }
            }
            __syncwarp();
        }
    };
    float val[32];
    if (I8 && go.raw) {
#pragma unroll
        for (int c = 0; c < 32; ++c) val[c] = __uint_as_float(v[c]);
        put(reinterpret_cast<float *>(go.raw), val);
        return;
    }
    if (I8 && (go.manhat || go.info)) {   // thermometer operands: integer epilogue, the reference's own formulas
        if (go.manhat) {
            const uint32_t si = (uint32_t)__double_as_longlong(ri.a);
#pragma unroll
            for (int c = 0; c < 32; ++c)
                val[c] = __uint_as_float(si + (uint32_t)__double_as_longlong(colk[c].a) - 2u * v[c]);
            put(reinterpret_cast<float *>(go.manhat), val);
        }
        if (go.info) {   // generation_dists @0x10001bc13: U == 0 ? 0 : (1.0f / (float)U) * (float)(U - I), single precision
            const uint32_t ni = (uint32_t)__double_as_longlong(ri.q);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const uint32_t U = ni + (uint32_t)__double_as_longlong(colk[c].q) - v[c];
                val[c] = U == 0 ? 0.0f : __fmul_rn(__fdiv_rn(1.0f, (float)U), (float)(U - v[c]));
            }
            put(go.info, val);
        }
        return;
    }
    if (go.cosine) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double g = I8 ? (double)(int)v[c] : (double)__uint_as_float(v[c]);
            val[c] = (float)(1.0 - g * ri.rn * colk[c].rn);
        }
        put(go.cosine, val);
    }
    if (go.pearson) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double g = I8 ? (double)(int)v[c] : (double)__uint_as_float(v[c]);
            val[c] = (float)(1.0 - (g - ri.S * colk[c].mu) * ri.rp * colk[c].rp);
        }
        put(go.pearson, val);
    }
    if (go.euclid) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const double g = I8 ? (double)(int)v[c] : (double)__uint_as_float(v[c]);
            const double e2 = ri.a + colk[c].a - 2.0 * g * ri.q * colk[c].q;
            val[c] = sqrtf((float)(e2 < 0.0 ? 0.0 : e2));   // cancellation below zero -> 0; NaN (an empty row) stays NaN
        }
        put(go.euclid, val);
    }
}

// CL = cluster size along M (1 or 2).  With CL == 2 the two CTAs of a cluster work on the tiles
// (2p, bj) and (2p+1, bj): they need the same B panel, so each CTA fetches half of it and TMA
// multicasts it into both shared memories -- 32 KB instead of 48 KB of L2 traffic per CTA and stage.
template <bool I8, int CL>
__global__ void __launch_bounds__(G_THREADS, 1) gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const uint2 *__restrict__ tiles, uint32_t n_tiles,
                                                                    uint32_t k_iters, GramGeom gg, GramOut go) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B: 1024-B aligned
    RowK *colk = reinterpret_cast<RowK *>(smem + G_STAGES * STAGE_BYTES);                 // [BN]
    float *stile = reinterpret_cast<float *>(colk + BN) + (((threadIdx.x >> 5) + 6) & 7) * EP_ROWS * EP_PITCH;   // per epilogue warp (warps 2..9)
    __shared__ __align__(8) uint64_t full[G_STAGES], empty[G_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t tmem_holder;

    // tile list entries are taken per CLUSTER (blockIdx.x / CL, stride gridDim.x / CL)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);   // every CTA of the cluster must have consumed the stage
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], EP_WARPS);   // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) {   // TMEM allocation is a warp-wide operation; this warp also frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, seen = 0;
            for (uint32_t t = blockIdx.x / CL; t < n_tiles; t += gridDim.x / CL) {
                uint2 tile = tiles[t];
                if (gg.ready) {   // fused exchange: the high half of .x is the last chunk this tile's WAVE needs; the
                                  // chunks arrive on two copy streams, so every flag up to it is looked at
                    for (const uint32_t need = tile.x >> 16; seen <= need; ++seen) wait_rows_ready(gg.ready + seen, gg.wait_probe);
                    tile.x &= 0xffffu;
                }
                const uint32_t bi = tile.x * CL + crank;
                for (uint32_t kb = 0; kb < k_iters; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    unsigned char *sa = smem + stage * STAGE_BYTES;
                    // inner coordinate in elements: 64 fp16 or 128 uint8 per 128-byte row
                    const int32_t kc = (int32_t)(kb * (I8 ? BK_BYTES : BK_BYTES / 2));
                    tma_load_2d(sa, &tmA, kc, (int32_t)(bi * BM), &full[stage]);
                    if constexpr (CL == 1) {
                        tma_load_2d(sa + A_BYTES, &tmB, kc, (int32_t)(tile.y * BN), &full[stage]);
                    } else {   // my half of the B panel, to both CTAs (tmB has a BN/CL-row box)
                        tma_load_2d_mc(sa + A_BYTES + crank * (B_BYTES / CL), &tmB, kc,
                                       (int32_t)(tile.y * BN + crank * (BN / CL)), &full[stage], CMASK);
                    }
                    if (++stage == G_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = blockIdx.x / CL; t < n_tiles; t += gridDim.x / CL) {
                mbar_wait(&tempty[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < k_iters; ++kb) {
                    mbar_wait(&full[stage], phase);        // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t a_desc = umma_desc(sa), b_desc = umma_desc(sa + A_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < BK_BYTES / 32; ++k)   // +32 B per MMA inside the 128-B swizzle row
                        umma<I8>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, (kb | k) != 0 ? 1u : 0u);
                    // smem slot free once these MMAs retire (in every CTA that was sent data for it)
                    if constexpr (CL == 1) umma_commit(&empty[stage]);
                    else umma_commit_mc(&empty[stage], CMASK);
                    if (++stage == G_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tfull[acc]);                  // accumulator complete
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: 8 warps, warp w may touch TMEM lanes [32*(w%4), +32); the two warps of a lane quarter
        // take alternate 32-column chunks =====
        const uint32_t q = (uint32_t)warp & 3u, half = ((uint32_t)warp - 2u) >> 2;
        const uint32_t et = threadIdx.x - 64;              // 0..255
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x / CL; t < n_tiles; t += gridDim.x / CL) {
            uint2 tile = tiles[t];
            if (gg.ready) tile.x &= 0xffffu;
            const uint32_t r0 = (tile.x * CL + crank) * BM, c0 = tile.y * BN;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EP_WARPS) : "memory");   // previous tile's colk fully consumed
            for (uint32_t c = et; c < (uint32_t)BN; c += 32 * EP_WARPS) {
                RowK z;
                z.rn = z.S = z.mu = z.rp = z.a = z.q = 0.0;
                colk[c] = (c0 + c < gg.n_b) ? gg.rk_b[c0 + c] : z;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EP_WARPS) : "memory");
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t row = r0 + q * 32 + lane;
            const bool row_ok = row < gg.n_a && row >= gg.row_begin && row < gg.row_end;
            RowK ri;
            ri.rn = ri.S = ri.mu = ri.rp = ri.a = ri.q = 0.0;
            if (row_ok) ri = gg.rk_a[row];
            for (uint32_t ch = half; ch < (uint32_t)BN / 32; ch += 2) {
                const uint32_t cb = c0 + ch * 32;
                if (gg.tri && cb + 31 <= r0 + q * 32) continue;   // whole chunk on or below the diagonal (warp-uniform)
                if (cb >= gg.n_b) continue;                         // whole chunk past the last column
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * BN + ch * 32 + ((q * 32) << 16), v);
                gram_store_chunk<I8>(gg, go, stile, v, ri, colk + ch * 32, cb, r0 + q * 32, (uint32_t)lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    if constexpr (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still write its smem / barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}


// ---- CTA-pair variant: tcgen05.mma.cta_group::2, M256 x N256 per instruction --------------------------
// The two CTAs of a cluster (one TPC) form ONE 256 x 256 tile: CTA r stages the A rows [128r, 128r+128)
// and the B rows [128r, 128r+128) of the tile -- 32 KB per stage and CTA, against 48 KB for the
// multicast variant above, so the shared-memory port carries 64 B/clk of TMA writes plus 64 B/clk of
// operand reads instead of 96 + 96 (the port moves 128 B/clk: the cta_group::1 kernel is bound by it at
// ~75 % tensor-pipe activity).  Only the leader CTA (rank 0) issues MMAs; the hardware reads the other
// half of B from the peer's shared memory and writes each CTA's 128 accumulator rows into its own TMEM.
// Both CTAs run a TMA producer (their loads complete on the LEADER's `full` barrier, address bit 24
// cleared) and four epilogue warps (which release the accumulator on the leader's `tempty` barrier).
constexpr int P_STAGES = 6;
constexpr uint32_t PB_BYTES = (BN / 2) * BK_BYTES, P_STAGE_BYTES = A_BYTES + PB_BYTES;
constexpr uint32_t IDESC2_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr uint32_t IDESC2_I8 = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA

template <bool I8>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    if constexpr (I8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC2_I8), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC2_F16), "r"(accumulate)
            : "memory");
    }
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// TMA load whose completion is counted on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tm, int32_t c_inner, int32_t c_outer,
                                                 uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c_inner), "r"(c_outer), "r"(smem_u32(bar) & PEER_MASK)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (spin > (1u << 26)) __trap();
    }
}

template <bool I8>
__global__ void __launch_bounds__(G_THREADS, 1) gram_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                 const __grid_constant__ CUtensorMap tmB,
                                                                 const uint2 *__restrict__ tiles, uint32_t n_tiles,
                                                                 uint32_t k_iters, GramGeom gg, GramOut go) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    RowK *colk = reinterpret_cast<RowK *>(smem + P_STAGES * P_STAGE_BYTES);   // [BN]
    float *stile = reinterpret_cast<float *>(colk + BN) + (((threadIdx.x >> 5) + 6) & 7) * EP_ROWS * EP_PITCH;   // per epilogue warp (warps 2..9)
    __shared__ __align__(8) uint64_t full[P_STAGES], empty[P_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t tmem_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(&full[s], 1);    // leader only: its producer's arrive.expect_tx (bytes of BOTH CTAs)
            mbar_init(&empty[s], 1);   // one multicast commit from the leader's MMA thread
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);   // one multicast commit
            mbar_init(&tempty[a], 2 * EP_WARPS);  // leader only: the epilogue warps of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 1) {   // the same warp of both CTAs allocates (and later frees) the pair's TMEM
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_holder;

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, seen = 0;
            for (uint32_t t = blockIdx.x / 2; t < n_tiles; t += gridDim.x / 2) {
                uint2 tile = tiles[t];
                if (gg.ready) {   // fused exchange: the high half of .x is the last chunk this tile's WAVE needs; the
                                  // chunks arrive on two copy streams, so every flag up to it is looked at
                    for (const uint32_t need = tile.x >> 16; seen <= need; ++seen) wait_rows_ready(gg.ready + seen, gg.wait_probe);
                    tile.x &= 0xffffu;
                }
                const int32_t arow = (int32_t)((tile.x * 2 + crank) * BM);
                const int32_t brow = (int32_t)(tile.y * BN + crank * (BN / 2));
                for (uint32_t kb = 0; kb < k_iters; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx(&full[stage], 2 * P_STAGE_BYTES);
                    unsigned char *sa = smem + stage * P_STAGE_BYTES;
                    const int32_t kc = (int32_t)(kb * (I8 ? BK_BYTES : BK_BYTES / 2));
                    tma_load_2d_pair(sa, &tmA, kc, arow, &full[stage]);
                    tma_load_2d_pair(sa + A_BYTES, &tmB, kc, brow, &full[stage]);
                    if (++stage == P_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA, for both CTAs =====
        if (leader && lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = blockIdx.x / 2; t < n_tiles; t += gridDim.x / 2) {
                mbar_wait_cluster(&tempty[acc], acc_phase ^ 1u);   // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < k_iters; ++kb) {
                    mbar_wait(&full[stage], phase);                // both CTAs' TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * P_STAGE_BYTES);
                    const uint64_t a_desc = umma_desc(sa), b_desc = umma_desc(sa + A_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < BK_BYTES / 32; ++k)
                        umma2<I8>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, (kb | k) != 0 ? 1u : 0u);
                    umma2_commit_mc(&empty[stage], 3);             // frees the stage in both CTAs
                    if (++stage == P_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma2_commit_mc(&tfull[acc], 3);                   // accumulator complete, both epilogues
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: 8 warps per CTA, CTA r owns rows [128r, 128r+128) of the tile; the two warps of a TMEM
        // lane quarter take alternate 32-column chunks =====
        const uint32_t q = (uint32_t)warp & 3u, half = ((uint32_t)warp - 2u) >> 2;
        const uint32_t et = threadIdx.x - 64;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x / 2; t < n_tiles; t += gridDim.x / 2) {
            uint2 tile = tiles[t];
            if (gg.ready) tile.x &= 0xffffu;
            const uint32_t r0 = (tile.x * 2 + crank) * BM, c0 = tile.y * BN;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EP_WARPS) : "memory");
            for (uint32_t c = et; c < (uint32_t)BN; c += 32 * EP_WARPS) {
                RowK z;
                z.rn = z.S = z.mu = z.rp = z.a = z.q = 0.0;
                colk[c] = (c0 + c < gg.n_b) ? gg.rk_b[c0 + c] : z;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EP_WARPS) : "memory");
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t row = r0 + q * 32 + lane;
            const bool row_ok = row < gg.n_a && row >= gg.row_begin && row < gg.row_end;
            RowK ri;
            ri.rn = ri.S = ri.mu = ri.rp = ri.a = ri.q = 0.0;
            if (row_ok) ri = gg.rk_a[row];
            for (uint32_t ch = half; ch < (uint32_t)BN / 32; ch += 2) {
                const uint32_t cb = c0 + ch * 32;
                if (gg.tri && cb + 31 <= r0 + q * 32) continue;
                if (cb >= gg.n_b) continue;
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * BN + ch * 32 + ((q * 32) << 16), v);
                gram_store_chunk<I8>(gg, go, stile, v, ri, colk + ch * 32, cb, r0 + q * 32, (uint32_t)lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[acc]);
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- pre-pass: ONE pass over the counts writes the tensor-core operand rows (uint8, saturating, or
// fp16; zero-padded to dpad) and the exact per-row sum / sum of squares, and records the largest
// count and the largest G_ii (the exactness envelope).  One CTA per row, 16 elements per thread and
// step with 16-byte loads and stores.
template <typename T, typename OP>
__global__ void __launch_bounds__(256) gram_prepass_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                           uint64_t dpad, OP *__restrict__ h,
                                                           unsigned long long *__restrict__ S,
                                                           unsigned long long *__restrict__ G,
                                                           unsigned long long *__restrict__ flags, int vec_ok) {
    __shared__ unsigned long long red[3][8];
    const uint32_t row = blockIdx.x;
    const T *xr = x + (uint64_t)row * ld;
    OP *hr = h + (uint64_t)row * dpad;
    unsigned long long s = 0, g = 0;
    unsigned mx = 0;
    if (vec_ok) {   // d % 16 == 0 and 16-byte aligned rows
        for (uint64_t e = (uint64_t)threadIdx.x * 16; e < dpad; e += 256 * 16) {
            unsigned c[16];
            if (e < d) {
                constexpr int NV = (int)sizeof(T);            // uint4 loads per 16 elements
                uint4 v[NV];
#pragma unroll
                for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const uint4 *>(xr + e) + i);
                const T *t = reinterpret_cast<const T *>(v);
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = (unsigned)t[i];
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = 0;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                s += c[i];
                g += (unsigned long long)c[i] * c[i];
                mx = max(mx, c[i]);
            }
            store_operand16<OP>(hr + e, c);
        }
    } else {
        for (uint64_t e = threadIdx.x; e < dpad; e += 256) {
            const unsigned c = e < d ? (unsigned)xr[e] : 0u;
            s += c;
            g += (unsigned long long)c * c;
            mx = max(mx, c);
            if constexpr (sizeof(OP) == 1) hr[e] = (OP)min(c, 255u);
            else hr[e] = __uint2half_rn(c);
        }
    }
    s = warp_sum64(s);
    g = warp_sum64(g);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = g;
        red[2][threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ts = 0, tg = 0, tm = 0;
        for (int w = 0; w < 8; ++w) {
            ts += red[0][w];
            tg += red[1][w];
            tm = red[2][w] > tm ? red[2][w] : tm;
        }
        S[row] = ts;
        G[row] = tg;
        atomicMax(&flags[0], tm);
        atomicMax(&flags[1], tg);
    }
}

// ---- limb-split path: counts up to 65535 and sums of squares far beyond 2^31 on the integer tensor cores ---
// c = 256 hi + lo  =>  G_ij = LL_ij + 256 (LH_ij + HL_ij) + 65536 HH_ij with LL = lo lo^T, ... : three
// kind::i8 Grams.  Every int32 accumulator stays exact as long as its FINAL value does (all terms are
// non-negative, so partial sums never exceed it), and by Cauchy-Schwarz LL_ij <= max_r sum lo_r^2 etc.; the
// pre-pass measures those maxima (flags[3], flags[4]).  Operand rows: XA = [lo | hi], XB = [hi | lo] (2 dpad
// bytes each): LL and HH are Grams of the two halves of XA, and LH + HL is ONE Gram of XA against XB over the
// doubled K.  The three int32 results of a block of rows go to scratch and a combine kernel applies the metric
// formulas to LL + 256 MID + 65536 HH in float64 (< 2^49: exact).
template <typename T>
__global__ void __launch_bounds__(256) gram_prepass_limb_kernel(const T *__restrict__ x, uint32_t n, uint64_t d,
                                                                uint64_t ld, uint64_t dpad, uint8_t *__restrict__ xa,
                                                                uint8_t *__restrict__ xb,
                                                                unsigned long long *__restrict__ S,
                                                                unsigned long long *__restrict__ G,
                                                                unsigned long long *__restrict__ flags) {
    __shared__ unsigned long long red[5][8];
    const uint32_t row = blockIdx.x;
    const T *xr = x + (uint64_t)row * ld;
    uint8_t *ra = xa + (uint64_t)row * 2 * dpad, *rb = xb + (uint64_t)row * 2 * dpad;
    unsigned long long s = 0, g = 0, ll = 0, hh = 0;
    unsigned mx = 0;
    for (uint64_t e = (uint64_t)threadIdx.x * 4; e < dpad; e += 256 * 4) {   // dpad is a multiple of 128
        uint32_t lo4 = 0, hi4 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned c = e + i < d ? (unsigned)xr[e + i] : 0u;
            const unsigned lo = c & 0xffu, hi = (c >> 8) & 0xffu;
            s += c;
            g += (unsigned long long)c * c;
            ll += lo * lo;
            hh += hi * hi;
            mx = max(mx, c);
            lo4 |= lo << (8 * i);
            hi4 |= hi << (8 * i);
        }
        *reinterpret_cast<uint32_t *>(ra + e) = lo4;
        *reinterpret_cast<uint32_t *>(ra + dpad + e) = hi4;
        *reinterpret_cast<uint32_t *>(rb + e) = hi4;
        *reinterpret_cast<uint32_t *>(rb + dpad + e) = lo4;
    }
    s = warp_sum64(s);
    g = warp_sum64(g);
    ll = warp_sum64(ll);
    hh = warp_sum64(hh);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = g;
        red[2][threadIdx.x >> 5] = mx;
        red[3][threadIdx.x >> 5] = ll;
        red[4][threadIdx.x >> 5] = hh;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < 8; ++w) {
            t[0] += red[0][w];
            t[1] += red[1][w];
            t[2] = red[2][w] > t[2] ? red[2][w] : t[2];
            t[3] += red[3][w];
            t[4] += red[4][w];
        }
        S[row] = t[0];
        G[row] = t[1];
        atomicMax(&flags[0], t[2]);
        atomicMax(&flags[1], t[1]);
        atomicMax(&flags[3], t[3]);
        atomicMax(&flags[4], t[4]);
    }
}

// one thread per (row, column) of the row block [r0, r1): exact G from the three partial Grams -> metrics
__global__ void __launch_bounds__(256) gram_limb_combine_kernel(const int *__restrict__ ll, const int *__restrict__ mid,
                                                                const int *__restrict__ hh, uint32_t n, uint32_t r0,
                                                                uint32_t r1, const RowK *__restrict__ rk, GramOut go) {
    const uint32_t row = r0 + blockIdx.y;
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (row >= r1 || j >= n || j <= row) return;
    const size_t at = (size_t)(row - r0) * n + j;
    const double g = (double)ll[at] + 256.0 * (double)mid[at] + 65536.0 * (double)hh[at];
    const long long idx = (long long)n * row - (long long)row * (row + 3) / 2 - 1 + j;
    gram_store(go, idx, g, rk[row], rk[j]);
}

// ---- thermometer operands: the reference's two metrics, `manhat` and `info`, as contractions ------------------
// sum_e |a_e - b_e| = sum(a) + sum(b) - 2 sum_e min(a_e, b_e), and min(a_e, b_e) = sum_{t >= 1} [a_e >= t][b_e >= t]:
// an integer Gram of rows of indicator bytes.  Column e gets as many bytes as the largest count ANY row holds there
// (colmax[e]; K = sum of the column maxima, typically 2-3 x 4^k for k-mer counts: a few hot k-mers cost a few extra
// bytes, not a level for every column).  One byte per column, (c_e >= 1), is the presence vector, whose Gram is
// |supp_a & supp_b| -- all `info` needs (union = nnz_a + nnz_b - intersection).  kind::i8 tcgen05 then does K MACs
// per pair where the CUDA cores do 4^k VABSDIFF4 / POPC lane-operations at ~1/25 of the rate.

// presence bytes (info): one CTA per row, 16 columns per thread and step
template <typename T>
__global__ void __launch_bounds__(256) thermo_presence_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                              uint64_t dpad, uint8_t *__restrict__ op,
                                                              RowK *__restrict__ rk, int vec_ok) {
    __shared__ unsigned long long red[2][8];
    const uint32_t row = blockIdx.x;
    const T *xr = x + (uint64_t)row * ld;
    uint8_t *out = op + (uint64_t)row * dpad;
    unsigned long long s = 0, nz = 0;
    for (uint64_t e = (uint64_t)threadIdx.x * 16; e < dpad; e += 256 * 16) {   // dpad is a multiple of 128
        unsigned c[16];
        if (vec_ok && e < d) {   // d % 16 == 0 and 16-byte aligned rows
            constexpr int NV = (int)sizeof(T);
            uint4 v[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const uint4 *>(xr + e) + i);
            const T *t = reinterpret_cast<const T *>(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = (unsigned)t[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = e + i < d ? (unsigned)xr[e + i] : 0u;
        }
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            w[q] = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s += c[4 * q + i];
                nz += c[4 * q + i] != 0;
                w[q] |= (c[4 * q + i] != 0 ? 1u : 0u) << (8 * i);
            }
        }
        *reinterpret_cast<uint4 *>(out + e) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    s = warp_sum64(s);
    nz = warp_sum64(nz);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = nz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ts = 0, tz = 0;
        for (int w = 0; w < 8; ++w) {
            ts += red[0][w];
            tz += red[1][w];
        }
        RowK r;
        r.rn = r.S = r.mu = r.rp = 0.0;
        r.a = __longlong_as_double((long long)ts);   // bit patterns of the integers (GramOut::manhat / info)
        r.q = __longlong_as_double((long long)tz);
        rk[row] = r;
    }
}

// per-column maxima: a thread owns VEC adjacent columns (one 4-byte load per row), a CTA a slab of rows
template <typename T, int VEC>
__global__ void __launch_bounds__(256) thermo_colmax_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                            uint32_t rows_per_cta, uint32_t *__restrict__ colmax) {
    const uint64_t e = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * VEC;
    if (e >= d) return;
    const uint32_t r0 = blockIdx.y * rows_per_cta, r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    unsigned m[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = 0;
    const T *p = x + (uint64_t)r0 * ld + e;
#pragma unroll 8
    for (uint32_t r = r0; r < r1; ++r, p += ld) {
        if constexpr (VEC == 1) {
            m[0] = max(m[0], (unsigned)__ldg(p));
        } else {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(p));
            if constexpr (VEC == 2) {
                m[0] = max(m[0], w & 0xffffu);
                m[1] = max(m[1], w >> 16);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) m[i] = max(m[i], (w >> (8 * i)) & 0xffu);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i)
        if (m[i] && e + i < d) atomicMax(&colmax[e + i], m[i]);
}

// off[e] = sum of colmax[0..e) (d + 1 entries), the largest count and the total; ONE CTA (d is 4^k: at most a few MB)
__global__ void __launch_bounds__(1024) thermo_scan_kernel(const uint32_t *__restrict__ colmax, uint64_t d,
                                                           uint32_t *__restrict__ off, unsigned long long *__restrict__ stats) {
    __shared__ unsigned long long wsum[32], step_total;
    __shared__ unsigned wmax[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long carry = 0;
    unsigned mx = 0;
    for (uint64_t base = 0; base < d; base += 4096) {
        const uint64_t e = base + (uint64_t)threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = e + i < d ? colmax[e + i] : 0u;
            mx = max(mx, v[i]);
        }
        unsigned long long t = (unsigned long long)v[0] + v[1] + v[2] + v[3], inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += u;
            }
            wsum[lane] = winc - w;            // exclusive prefix of the warp totals
            if (lane == 31) step_total = winc;
        }
        __syncthreads();
        unsigned long long ex = carry + wsum[warp] + (inc - t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (e + i < d) off[e + i] = (uint32_t)ex;
            ex += v[i];
        }
        carry += step_total;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) wmax[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m = 0;
        for (int w = 0; w < 32; ++w) m = max(m, wmax[w]);
        off[d] = (uint32_t)carry;
        stats[0] = m;
        stats[1] = carry;
    }
}

// operand rows: a CTA takes TH_ROWS rows and walks the columns in chunks (sized by the host so that a chunk's bytes
// fit TH_SB); the bytes of a chunk are composed in shared memory (a column's bytes sit at off[e] - chunk origin) and
// leave as 16-byte stores.  k-mer count rows are mostly zeros (a 150 bp read fills 3 % of 4^6 bins, a 9 kb genome
// 3 % of 4^9), so the chunk is zero-filled with vector stores and only the NON-ZERO counts touch it afterwards: a
// thread owns 16 columns -- their 17 offsets are loaded once, with 16-byte loads, and serve all rows -- reads them
// row by row with 16-byte loads and writes c one-bytes for the few columns that hold something.  (The first version
// wrote every byte of every column from a per-column loop through a generic pointer: ~90 instructions per element,
// 2.0 ms for 131 072 x 4096 against 0.74 ms for the Gram it feeds.)  A chunk whose bytes do not fit the buffer
// (very large column maxima) is written directly.  Also sum(c), nnz(c).
constexpr int TH_ROWS = 4, TH_SB = 12160;   // 4 x 12160 + the reduction scratch = the 48 KB of static shared memory
template <typename T>
__global__ void __launch_bounds__(256, 3) thermo_encode_kernel(const T *__restrict__ x, uint32_t n, uint64_t d, uint64_t ld,
                                                            const uint32_t *__restrict__ off, uint64_t kpad,
                                                            uint8_t *__restrict__ op, RowK *__restrict__ rk, int vec_ok,
                                                            uint32_t chunk) {
    __shared__ __align__(16) uint8_t sbuf[TH_ROWS][TH_SB];
    __shared__ unsigned long long red[2][TH_ROWS][8];
    const uint32_t row0 = blockIdx.x * TH_ROWS, tid = threadIdx.x;
    const uint32_t nrows = n - row0 < (uint32_t)TH_ROWS ? n - row0 : (uint32_t)TH_ROWS;
    const bool off_vec = (((uintptr_t)off) & 15) == 0;
    unsigned long long s[TH_ROWS], nz[TH_ROWS];
#pragma unroll
    for (int r = 0; r < TH_ROWS; ++r) s[r] = nz[r] = 0;
    for (uint64_t e0 = 0; e0 < d; e0 += chunk) {
        const uint64_t e1 = e0 + chunk < d ? e0 + chunk : d;
        const uint32_t kb = off[e0], ke = off[e1];
        const uint32_t a0 = kb & ~15u;
        const bool staged = ke - a0 <= (uint32_t)TH_SB;
        if (staged) {
            const uint32_t words = (ke - a0 + 15u) >> 4;
            for (uint32_t r = 0; r < nrows; ++r)
                for (uint32_t g = tid; g < words; g += 256) reinterpret_cast<uint4 *>(sbuf[r])[g] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
        for (uint64_t e = e0 + (uint64_t)tid * 16; e < e1; e += 256 * 16) {
            uint32_t o[17];
            if (off_vec && e + 16 <= d) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(off + e) + i);
                    o[4 * i] = v.x, o[4 * i + 1] = v.y, o[4 * i + 2] = v.z, o[4 * i + 3] = v.w;
                }
                o[16] = __ldg(off + e + 16);
            } else {
#pragma unroll
                for (int i = 0; i < 17; ++i) o[i] = off[e + i < d ? e + i : d];
            }
            const bool full = vec_ok && e + 16 <= e1;   // d % 16 == 0 and 16-byte aligned rows
            constexpr int NV = (int)sizeof(T);
            uint4 v[TH_ROWS][NV];                        // all rows' loads in flight before any is looked at
            if (full) {
#pragma unroll
                for (int r = 0; r < TH_ROWS; ++r)
                    if ((uint32_t)r < nrows) {
                        const T *xr = x + (uint64_t)(row0 + r) * ld;
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[r][i] = __ldg(reinterpret_cast<const uint4 *>(xr + e) + i);
                    }
            }
#pragma unroll
            for (int r = 0; r < TH_ROWS; ++r) {
                if ((uint32_t)r >= nrows) break;
                const T *xr = x + (uint64_t)(row0 + r) * ld;
                unsigned c[16];
                if (full) {
                    const T *t = reinterpret_cast<const T *>(v[r]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) c[i] = (unsigned)t[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) c[i] = e + i < e1 ? (unsigned)xr[e + i] : 0u;
                }
                unsigned any = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    s[r] += c[i];
                    nz[r] += c[i] != 0;
                    any |= c[i];
                }
                if (staged) {
                    if (any) {
                        uint8_t *rowbuf = sbuf[r] - a0;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c[i]) {
                                uint8_t *dst = rowbuf + o[i];
                                dst[0] = 1;
                                for (uint32_t t = 1; t < c[i]; ++t) dst[t] = 1;
                            }
                    }
                } else {   // direct: every byte of every column, zeros included
                    uint8_t *grow = op + (uint64_t)(row0 + r) * kpad;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (e + i < e1)
                            for (uint32_t t = 0, cm = o[i + 1] - o[i]; t < cm; ++t) grow[o[i] + t] = c[i] > t ? 1 : 0;
                }
            }
        }
        if (staged) {
            __syncthreads();
            const uint32_t up = (kb + 15u) & ~15u;
            const uint32_t head_end = up < ke ? up : ke;
            const uint32_t tail_begin = (ke & ~15u) > head_end ? (ke & ~15u) : head_end;
            for (uint32_t r = 0; r < nrows; ++r) {
                uint8_t *base = op + (uint64_t)(row0 + r) * kpad;
                for (uint32_t g = head_end + tid * 16; g < tail_begin; g += 256 * 16)
                    *reinterpret_cast<uint4 *>(base + g) = *reinterpret_cast<const uint4 *>(&sbuf[r][g - a0]);
                if (tid < 16) {
                    const uint32_t p = kb + tid;
                    if (p < head_end) base[p] = sbuf[r][p - a0];
                } else if (tid < 32) {
                    const uint32_t p = tail_begin + (tid - 16);
                    if (p < ke) base[p] = sbuf[r][p - a0];
                }
            }
            __syncthreads();
        }
    }
    const uint32_t k_total = off[d];
    for (uint32_t r = 0; r < nrows; ++r)
        for (uint64_t p = (uint64_t)k_total + tid; p < kpad; p += 256) op[(uint64_t)(row0 + r) * kpad + p] = 0;
#pragma unroll
    for (int r = 0; r < TH_ROWS; ++r) {
        const unsigned long long a = warp_sum64(s[r]), b = warp_sum64(nz[r]);
        if ((tid & 31) == 0) {
            red[0][r][tid >> 5] = a;
            red[1][r][tid >> 5] = b;
        }
    }
    __syncthreads();
    if (tid < nrows) {
        unsigned long long ts = 0, tz = 0;
        for (int w = 0; w < 8; ++w) {
            ts += red[0][tid][w];
            tz += red[1][tid][w];
        }
        RowK r;
        r.rn = r.S = r.mu = r.rp = 0.0;
        r.a = __longlong_as_double((long long)ts);
        r.q = __longlong_as_double((long long)tz);
        rk[row0 + tid] = r;
    }
}

__global__ void rowk_kernel(const unsigned long long *__restrict__ S, const unsigned long long *__restrict__ G,
                            uint32_t n, double D, int euclid_on_freq, RowK *__restrict__ rk) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = (double)S[i], g = (double)G[i];
    RowK r;
    r.S = s;
    r.mu = s / D;
    r.rn = 1.0 / sqrt(g);
    r.rp = 1.0 / sqrt(g - s * s / D);
    if (euclid_on_freq) {
        r.q = 1.0 / s;
        r.a = g / (s * s);
    } else {
        r.q = 1.0;
        r.a = g;
    }
    rk[i] = r;
}

// thermometer operands across processes: the integers behind RowK.a / .q travel as plain uint64 arrays
__global__ void thermo_stats_from_rowk(const RowK *__restrict__ rk, uint32_t n, unsigned long long *__restrict__ S,
                                       unsigned long long *__restrict__ NZ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S[i] = (unsigned long long)__double_as_longlong(rk[i].a);
    NZ[i] = (unsigned long long)__double_as_longlong(rk[i].q);
}
__global__ void thermo_rowk_from_stats(const unsigned long long *__restrict__ S, const unsigned long long *__restrict__ NZ,
                                       uint32_t n, RowK *__restrict__ rk) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowK r;
    r.rn = r.S = r.mu = r.rp = 0.0;
    r.a = __longlong_as_double((long long)S[i]);
    r.q = __longlong_as_double((long long)NZ[i]);
    rk[i] = r;
}

bool g_force_f16 = false;
bool g_float_recover = true;   // test hook: false = floating-point rows always take the float64 CUDA-core Gram
bool g_limb = true;        // test hook: false = inputs outside the uint8 / fp16 envelopes skip the limb-split tensor path
bool g_pair = true;        // cta_group::2 kernel (gram_pair_kernel, default) or the cluster-multicast one
uint32_t g_cluster = 2;
uint32_t g_strip = 12;      // row tiles (of 128) per strip of the tile order (tuning hook)
unsigned long long *g_wait_probe = nullptr;   // diagnostics hook of the fused exchange (kam_gram_set_wait_probe)

inline uint64_t dpad_of(uint64_t d) { return (d + BK_BYTES - 1) / BK_BYTES * BK_BYTES; }   // elements (either width)

inline size_t max_tiles(uint32_t n_a, uint32_t n_b) {
    return ((size_t)(n_a + BM - 1) / BM + 1) * ((size_t)(n_b + BN - 1) / BN);
}

template <typename T>
int run_prepass(const void *x, uint32_t n, uint64_t d, uint64_t ld, bool i8, void *h, unsigned long long *S,
                unsigned long long *G, unsigned long long *flags, cudaStream_t st) {
    const uint64_t dpad = dpad_of(d);
    const int vec_ok = (d % 16 == 0) && ((ld * sizeof(T)) % 16 == 0) && (((uintptr_t)x & 15) == 0);
    if (i8)
        gram_prepass_kernel<T, uint8_t><<<n, 256, 0, st>>>((const T *)x, n, d, ld, dpad, (uint8_t *)h, S, G, flags, vec_ok);
    else
        gram_prepass_kernel<T, __half><<<n, 256, 0, st>>>((const T *)x, n, d, ld, dpad, (__half *)h, S, G, flags, vec_ok);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

int prepass_dispatch(const void *x, int elem, uint32_t n, uint64_t d, uint64_t ld, bool i8, void *h,
                     unsigned long long *S, unsigned long long *G, unsigned long long *flags, cudaStream_t st) {
    if (elem == KAM_F32 || elem == KAM_F64)   // frequency rows: integer recovery fused into the sweep (freq_gram.cu)
        return kam_float_prepass(x, elem, n, d, ld, dpad_of(d), i8, h, S, G, flags, st);
    return elem == KAM_U8    ? run_prepass<uint8_t>(x, n, d, ld, i8, h, S, G, flags, st)
           : elem == KAM_U16 ? run_prepass<uint16_t>(x, n, d, ld, i8, h, S, G, flags, st)
                             : run_prepass<uint32_t>(x, n, d, ld, i8, h, S, G, flags, st);
}

// kext: K extent of the operand rows in elements (0 = dpad_of(d)); pitch: their row pitch in bytes (0 = packed).
// The limb-split path runs on sub-ranges of rows that hold [low bytes | high bytes] side by side.
template <bool I8, int CL, bool PAIR = false>
int launch_gram(const void *opA, const void *opB, const uint2 *d_tiles, size_t n_tiles, uint64_t d,
                const GramGeom &gg, GramOut go, uint32_t sm_limit, cudaStream_t st, uint64_t kext = 0,
                uint64_t pitch = 0) {
    const uint32_t esz = I8 ? 1 : 2, bk = BK_BYTES / esz;
    const uint64_t dpad = kext ? kext : dpad_of(d);
    if (!pitch) pitch = dpad * esz;
    const CUtensorMapDataType dt = I8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB;
    int rc;
    if ((rc = kam_make_tmap_2d(&tmA, dt, esz, opA, dpad, gg.n_a, pitch, bk, BM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = kam_make_tmap_2d(&tmB, dt, esz, opB, dpad, gg.n_b, pitch, bk, BN / CL, CU_TENSOR_MAP_SWIZZLE_128B)))
        return rc;
    const size_t smem = (PAIR ? (size_t)P_STAGES * P_STAGE_BYTES : (size_t)G_STAGES * STAGE_BYTES) + BN * sizeof(RowK) +
                        EP_TILE_BYTES + 1024;
    auto kern = PAIR ? gram_pair_kernel<I8> : gram_tcgen05_kernel<I8, CL>;
    KAM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t sms = (uint32_t)kam_sm_count();
    if (sm_limit && sm_limit < sms) sms = sm_limit;
    uint32_t clusters = sms / CL;
    if (clusters < 1) clusters = 1;
    if (clusters > n_tiles) clusters = (uint32_t)n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    KAM_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, d_tiles, (uint32_t)n_tiles, (uint32_t)(dpad / bk), gg, go));
    return KAM_OK;
}

// Tile lists are cached per geometry: the list of a block depends only on (first row group, last row group,
// column tiles, triangular?, cluster size), it lives in device memory owned by the library next to the pinned
// host copy it was uploaded from, so a launch neither waits for the upload nor re-builds the list (the first
// version built a std::vector per call and synchronised the stream before it went out of scope).
struct TileList {
    int dev = -1;
    uint32_t tr0 = 0, tr1 = 0, tc = 0, tri = 0, cl = 0, strip = 0;
    uint64_t sig = 0;           // 0: the list of one block; else the hash of a fused call's segments
    uint2 *d = nullptr, *h = nullptr;
    cudaEvent_t up = nullptr;   // the upload; launches on other streams wait for it
    size_t n = 0;
    uint64_t stamp = 0;
};
constexpr int TILE_CACHE = 32;
TileList g_tiles[TILE_CACHE];
uint64_t g_tile_stamp = 0;
std::mutex g_tile_mu;

// one strip-ordered rectangle of tiles: row groups [g0, g1) x column tiles [c0, c1).  A list entry is
// (row-tile group, column tile); a group is CL consecutive 128-row tiles handled by the CL CTAs of a cluster.
// Strips of ~12 row tiles, column-major inside a strip, so the ~148 tiles in flight share about a dozen A panels
// and a dozen B panels in L2.
void push_rect(std::vector<uint2> &tiles, uint32_t CL, uint32_t g0, uint32_t g1, uint32_t c0, uint32_t c1, bool tri,
               uint32_t strip_tiles = 0) {
    const uint32_t GM = BM * CL;
    if (!strip_tiles) strip_tiles = g_strip;
    const uint32_t STRIP = strip_tiles / CL ? strip_tiles / CL : 1;
    for (uint32_t s0 = g0; s0 < g1; s0 += STRIP) {
        const uint32_t s1 = s0 + STRIP < g1 ? s0 + STRIP : g1;
        for (uint32_t bj = c0; bj < c1; ++bj)
            for (uint32_t bg = s0; bg < s1; ++bg)
                if (!tri || (uint64_t)bj * BN + BN - 1 > (uint64_t)bg * GM) tiles.push_back(make_uint2(bg, bj));
    }
}

// The list of a key, built by `build` on a miss and kept in device memory owned by the library next to the
// pinned host copy it was uploaded from.
template <class Build>
int tile_list_cached(const TileList &key, cudaStream_t st, const uint2 **d_list, size_t *n_list, Build build) {
    int dev = 0;
    KAM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tile_mu);
    TileList *slot = &g_tiles[0];
    for (auto &t : g_tiles) {
        if (t.d && t.dev == dev && t.tr0 == key.tr0 && t.tr1 == key.tr1 && t.tc == key.tc && t.tri == key.tri &&
            t.cl == key.cl && t.strip == g_strip && t.sig == key.sig) {
            KAM_CUDA(cudaStreamWaitEvent(st, t.up, 0));
            t.stamp = ++g_tile_stamp;
            *d_list = t.d;
            *n_list = t.n;
            return KAM_OK;
        }
        if (t.stamp < slot->stamp) slot = &t;
    }
    std::vector<uint2> tiles;
    build(tiles);
    if (slot->d) {   // evict: earlier launches that read this list were enqueued on streams we do not know
        int cur = dev;
        if (slot->dev != dev) cudaSetDevice(slot->dev);
        cudaDeviceSynchronize();
        cudaFree(slot->d);
        cudaFreeHost(slot->h);
        cudaEventDestroy(slot->up);
        if (slot->dev != cur) cudaSetDevice(cur);
        slot->d = slot->h = nullptr;
        slot->up = nullptr;
    }
    slot->n = tiles.size();
    *d_list = nullptr;
    *n_list = 0;
    if (tiles.empty()) return KAM_OK;
    KAM_CUDA(cudaMalloc(&slot->d, tiles.size() * sizeof(uint2)));
    if (cudaMallocHost(&slot->h, tiles.size() * sizeof(uint2)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(slot->d);
        slot->d = nullptr;
        kam_set_error("cudaMallocHost failed for a Gram tile list");
        return KAM_E_NOMEM;
    }
    memcpy(slot->h, tiles.data(), tiles.size() * sizeof(uint2));
    KAM_CUDA(cudaMemcpyAsync(slot->d, slot->h, tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
    KAM_CUDA(cudaEventCreateWithFlags(&slot->up, cudaEventDisableTiming));
    KAM_CUDA(cudaEventRecord(slot->up, st));
    slot->dev = dev;
    slot->tr0 = key.tr0, slot->tr1 = key.tr1, slot->tc = key.tc, slot->tri = key.tri, slot->cl = key.cl;
    slot->strip = g_strip, slot->sig = key.sig;
    slot->stamp = ++g_tile_stamp;
    *d_list = slot->d;
    *n_list = slot->n;
    return KAM_OK;
}

// Tile lists are cached per geometry: the list of a block depends only on (first row group, last row group,
// column tiles, triangular?, cluster size), so a launch neither waits for the upload nor re-builds the list (the
// first version built a std::vector per call and synchronised the stream before it went out of scope).
int tile_list(const GramGeom &gg, uint32_t CL, cudaStream_t st, const uint2 **d_list, size_t *n_list) {
    const uint32_t GM = BM * CL;
    TileList key;
    key.tr0 = gg.row_begin / GM, key.tr1 = (gg.row_end - 1) / GM + 1, key.tc = (gg.n_b + BN - 1) / BN;
    key.tri = gg.tri, key.cl = CL;
    return tile_list_cached(key, st, d_list, n_list, [&](std::vector<uint2> &tiles) {
        push_rect(tiles, CL, key.tr0, key.tr1, 0, key.tc, gg.tri != 0);
    });
}

// A fused call's list: the rectangles of its segments one after the other, in the order their B rows arrive.
// A segment is (first column tile, end column tile, first row tile, end row tile) in units of 256 rows.
int tile_list_segments(const GramGeom &gg, const uint32_t *segs, uint32_t n_segs, uint32_t CL, uint32_t pairs,
                       cudaStream_t st, const uint2 **d_list, size_t *n_list) {
    TileList key;
    key.tr0 = 0, key.tr1 = (gg.n_a + BM * CL - 1) / (BM * CL), key.tc = (gg.n_b + BN - 1) / BN;
    key.tri = gg.tri, key.cl = CL;
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < 4 * n_segs; ++i) h = (h ^ segs[i]) * 1099511628211ull;
    h = (h ^ pairs) * 1099511628211ull;
    h = (h ^ (gg.ready ? gg.ready_tiles : 0u)) * 1099511628211ull;
    key.sig = h | 1ull;
    const uint32_t per = 256 / (BM * CL);   // row groups per 256-row tile (1 for clusters of 2, 2 without)
    return tile_list_cached(key, st, d_list, n_list, [&](std::vector<uint2> &tiles) {
        for (uint32_t i = 0; i < n_segs; ++i) {
            const uint32_t *q = segs + 4 * i;
            const uint32_t c1 = q[1] < key.tc ? q[1] : key.tc;
            const uint32_t g1 = q[3] * per < key.tr1 ? q[3] * per : key.tr1;
            // strips twice as tall as for a resident operand: the ~74 tiles in flight then span 6 columns x 12 row
            // groups instead of 12 x 6 -- the same number of panels in L2, half the look-ahead into rows that are
            // still arriving (8 ranks x 5000 rows: no wait at all instead of 0.3 ms per wave)
            push_rect(tiles, CL, q[2] * per, g1, q[0], c1, gg.tri != 0, gg.ready ? 2 * g_strip : 0);
        }
        if (gg.ready) {   // per wave of `pairs` tiles: the chunk to wait for, in the high half of .x
            const size_t P = pairs < tiles.size() ? pairs : tiles.size();
            uint32_t need = 0;
            for (size_t w0 = 0; w0 < tiles.size(); w0 += P) {
                const size_t w1 = w0 + P < tiles.size() ? w0 + P : tiles.size();
                for (size_t t = w0; t < w1; ++t) need = std::max(need, tiles[t].y / gg.ready_tiles);
                for (size_t t = w0; t < w1; ++t) tiles[t].x |= need << 16;
            }
        }
    });
}

int gram_block(const void *opA, const void *opB, bool i8, uint64_t d, const GramGeom &gg, GramOut go, uint32_t sm_limit,
               cudaStream_t st, uint64_t kext = 0, uint64_t pitch = 0, const uint32_t *segs = nullptr,
               uint32_t n_segs = 0) {
    const uint32_t CL = g_cluster;
    const uint2 *d_tiles = nullptr;
    size_t n_tiles = 0;
    uint32_t sms = (uint32_t)kam_sm_count();
    if (sm_limit && sm_limit < sms) sms = sm_limit;
    const uint32_t pairs = sms / CL ? sms / CL : 1u;      // clusters of the launch (launch_gram)
    int rc = segs ? tile_list_segments(gg, segs, n_segs, CL, pairs, st, &d_tiles, &n_tiles)
                  : tile_list(gg, CL, st, &d_tiles, &n_tiles);
    if (rc) return rc;
    if (n_tiles == 0) return KAM_OK;
    if (CL == 2 && g_pair)
        return i8 ? launch_gram<true, 2, true>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch)
                  : launch_gram<false, 2, true>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch);
    if (CL == 2)
        return i8 ? launch_gram<true, 2>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch)
                  : launch_gram<false, 2>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch);
    return i8 ? launch_gram<true, 1>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch)
              : launch_gram<false, 1>(opA, opB, d_tiles, n_tiles, d, gg, go, sm_limit, st, kext, pitch);
}

// limb-split path: rows per scratch block (a multiple of 256) and the bytes it needs next to the operands
inline uint32_t limb_rows(uint32_t n) {
    uint64_t r = (256ull << 20) / ((uint64_t)n * 12) / 256 * 256;
    const uint64_t nup = ((uint64_t)n + 255) / 256 * 256;
    if (r < 256) r = 256;
    if (r > nup) r = nup;
    return (uint32_t)r;
}
inline size_t limb_bytes(uint32_t n, uint64_t d) {
    return kam_align_up((size_t)n * dpad_of(d) * 4, 1024) + (size_t)3 * limb_rows(n) * n * 4 + 1024;
}

struct GramWs {
    void *h;
    unsigned long long *S, *G, *flags;
    RowK *rk;
    uint2 *tiles;
    uint32_t *xt;   // integer fallback operand
};

// with_int: also reserve room for the integer-fallback operand (4 bytes per element, much larger)
inline size_t gram_layout(uint32_t n, uint64_t d, void *base, GramWs *w, bool with_int) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *p = base ? (char *)base + off : nullptr;
        off += kam_align_up(bytes, 1024);
        return p;
    };
    void *flags = take(64);
    void *S = take((size_t)n * 8), *G = take((size_t)n * 8);
    void *rk = take((size_t)n * sizeof(RowK));
    void *tiles = take(max_tiles(n, n) * sizeof(uint2));
    // the fp16 matrix and the integer-fallback operand are never live together: share the space
    const size_t hb = (size_t)n * dpad_of(d) * 2, ib = with_int ? kam_int_gram_xt_bytes(n, d) : 0;
    const size_t lb = with_int ? limb_bytes(n, d) : 0;
    void *big = take(std::max(hb, std::max(ib, lb)));
    if (w) {
        w->flags = (unsigned long long *)flags;
        w->S = (unsigned long long *)S;
        w->G = (unsigned long long *)G;
        w->rk = (RowK *)rk;
        w->tiles = (uint2 *)tiles;
        w->h = big;
        w->xt = (uint32_t *)big;
    }
    return off;
}

int read_flags(const unsigned long long *d_flags, unsigned long long (&h)[3], cudaStream_t st) {
    KAM_CUDA(cudaMemcpyAsync(h, d_flags, 24, cudaMemcpyDeviceToHost, st));
    KAM_CUDA(cudaStreamSynchronize(st));
    return KAM_OK;
}

}  // namespace

size_t kam_thermo_cols_bytes(uint64_t d) { return kam_align_up((size_t)(2 * d + 2) * 4 + 16, 1024); }
uint64_t kam_thermo_kpad(uint64_t k_total) { return k_total ? (k_total + BK_BYTES - 1) / BK_BYTES * BK_BYTES : BK_BYTES; }
size_t kam_thermo_bytes(uint32_t n, uint64_t kpad) {
    return kam_align_up((size_t)n * sizeof(RowK), 1024) + kam_align_up((size_t)n * kpad, 1024) + 1024;
}

int kam_thermo_colmax(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *colmax, cudaStream_t st) {
    if (n == 0) return KAM_OK;
    const size_t esz = elem == KAM_U8 ? 1 : elem == KAM_U16 ? 2 : 4;
    const bool vec = ((ld * esz) % 4 == 0) && (((uintptr_t)d_x & 3) == 0) && (d * esz) % 4 == 0;
    const uint32_t per = vec ? (uint32_t)(4 / esz) : 1u;
    const uint32_t gx = (uint32_t)((d + 256ull * per - 1) / (256ull * per));
    uint32_t slabs = (uint32_t)kam_sm_count() * 8 / (gx ? gx : 1);
    if (slabs < 1) slabs = 1;
    if (slabs > (n + 63) / 64) slabs = (n + 63) / 64;       // at least 64 rows per CTA
    if (slabs > 65535) slabs = 65535;
    const uint32_t rows_per = (n + slabs - 1) / slabs;
    const dim3 grid(gx, (n + rows_per - 1) / rows_per);
    if (elem == KAM_U8) {
        if (vec) thermo_colmax_kernel<uint8_t, 4><<<grid, 256, 0, st>>>((const uint8_t *)d_x, n, d, ld, rows_per, colmax);
        else thermo_colmax_kernel<uint8_t, 1><<<grid, 256, 0, st>>>((const uint8_t *)d_x, n, d, ld, rows_per, colmax);
    } else if (elem == KAM_U16) {
        if (vec) thermo_colmax_kernel<uint16_t, 2><<<grid, 256, 0, st>>>((const uint16_t *)d_x, n, d, ld, rows_per, colmax);
        else thermo_colmax_kernel<uint16_t, 1><<<grid, 256, 0, st>>>((const uint16_t *)d_x, n, d, ld, rows_per, colmax);
    } else {
        thermo_colmax_kernel<uint32_t, 1><<<grid, 256, 0, st>>>((const uint32_t *)d_x, n, d, ld, rows_per, colmax);
    }
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

int kam_thermo_layout(const uint32_t *colmax, uint64_t d, uint32_t *off, unsigned long long *d_stats,
                      unsigned long long (&h_stats)[2], cudaStream_t st) {
    thermo_scan_kernel<<<1, 1024, 0, st>>>(colmax, d, off, d_stats);
    KAM_CUDA(cudaGetLastError());
    KAM_CUDA(cudaMemcpyAsync(h_stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    KAM_CUDA(cudaStreamSynchronize(st));
    return KAM_OK;
}

static int thermo_encode_to(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *off,
                            uint64_t kpad, uint8_t *op, RowK *rk, cudaStream_t st);
static int thermo_presence_to(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint8_t *op, RowK *rk,
                              cudaStream_t st);

int kam_thermo_encode(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *off,
                      uint64_t kpad, void *d_ws, ThermoOp *out, cudaStream_t st) {
    RowK *rk = (RowK *)d_ws;
    uint8_t *op = (uint8_t *)d_ws + kam_align_up((size_t)n * sizeof(RowK), 1024);
    int rc = thermo_encode_to(d_x, elem, n, d, ld, off, kpad, op, rk, st);
    if (rc) return rc;
    out->op = op;
    out->rk = rk;
    out->n = n;
    out->d = d;
    out->kpad = kpad;
    return KAM_OK;
}

static int thermo_encode_to(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *off,
                            uint64_t kpad, uint8_t *op, RowK *rk, cudaStream_t st) {
    if (n) {
        const uint32_t grid = (n + TH_ROWS - 1) / TH_ROWS;
        const size_t esz = elem == KAM_U8 ? 1 : elem == KAM_U16 ? 2 : 4;
        const int vec_ok = (d % 16 == 0) && ((ld * esz) % 16 == 0) && (((uintptr_t)d_x & 15) == 0);
        // columns per chunk: as many as keep a chunk's bytes inside the kernel's buffer at the average bytes per
        // column (x 1.25 for unevenness), whole groups of 4096 where that allows: 256 threads x 16 columns
        const uint64_t avg_x4 = (kpad * 5 + d - 1) / d;                       // 1.25 x 4 x bytes per column
        uint64_t chunk = (uint64_t)TH_SB * 4 / (avg_x4 ? avg_x4 : 1);
        chunk = chunk >= 4096 ? 4096 : chunk >= 16 ? chunk / 16 * 16 : 16;
        if (kpad <= (uint64_t)TH_SB && d <= 4096) chunk = 4096;              // a whole row fits: one chunk
        if (elem == KAM_U8)
            thermo_encode_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)d_x, n, d, ld, off, kpad, op, rk, vec_ok, (uint32_t)chunk);
        else if (elem == KAM_U16)
            thermo_encode_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)d_x, n, d, ld, off, kpad, op, rk, vec_ok, (uint32_t)chunk);
        else
            thermo_encode_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t *)d_x, n, d, ld, off, kpad, op, rk, vec_ok, (uint32_t)chunk);
        KAM_CUDA(cudaGetLastError());
    }
    return KAM_OK;
}

int kam_thermo_presence(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, void *d_ws, ThermoOp *out,
                        cudaStream_t st) {
    RowK *rk = (RowK *)d_ws;
    uint8_t *op = (uint8_t *)d_ws + kam_align_up((size_t)n * sizeof(RowK), 1024);
    int rc = thermo_presence_to(d_x, elem, n, d, ld, op, rk, st);
    if (rc) return rc;
    out->op = op;
    out->rk = rk;
    out->n = n;
    out->d = d;
    out->kpad = dpad_of(d);
    return KAM_OK;
}

static int thermo_presence_to(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint8_t *op, RowK *rk,
                              cudaStream_t st) {
    const uint64_t dpad = dpad_of(d);
    const size_t esz = elem == KAM_U8 ? 1 : elem == KAM_U16 ? 2 : 4;
    const int vec_ok = (d % 16 == 0) && ((ld * esz) % 16 == 0) && (((uintptr_t)d_x & 15) == 0);
    if (n) {
        if (elem == KAM_U8)
            thermo_presence_kernel<uint8_t><<<n, 256, 0, st>>>((const uint8_t *)d_x, n, d, ld, dpad, op, rk, vec_ok);
        else if (elem == KAM_U16)
            thermo_presence_kernel<uint16_t><<<n, 256, 0, st>>>((const uint16_t *)d_x, n, d, ld, dpad, op, rk, vec_ok);
        else
            thermo_presence_kernel<uint32_t><<<n, 256, 0, st>>>((const uint32_t *)d_x, n, d, ld, dpad, op, rk, vec_ok);
        KAM_CUDA(cudaGetLastError());
    }
    return KAM_OK;
}

int kam_thermo_pairs(const ThermoOp &A, const ThermoOp &B, bool tri, uint32_t row_begin, uint32_t row_end,
                     uint32_t *d_manhat, float *d_info, uint64_t out_ld, cudaStream_t st) {
    if (A.n == 0 || B.n == 0 || row_begin >= row_end) return KAM_OK;
    if (A.kpad != B.kpad || (d_manhat != nullptr) == (d_info != nullptr)) {
        kam_set_error("kam_thermo_pairs: operands of different layouts, or not exactly one output");
        return KAM_E_INVALID;
    }
    GramGeom gg;
    gg.n_a = A.n;
    gg.n_b = B.n;
    gg.row_begin = row_begin;
    gg.row_end = row_end < A.n ? row_end : A.n;
    gg.tri = tri ? 1u : 0u;
    gg.ld = out_ld;
    gg.rk_a = A.rk;
    gg.rk_b = B.rk;
    GramOut go;
    go.euclid = go.cosine = go.pearson = nullptr;
    go.manhat = d_manhat;
    go.info = d_info;
    return gram_block(A.op, B.op, true, A.d, gg, go, 0, st, A.kpad, A.kpad);
}

extern "C" {

size_t kam_gram_workspace_bytes(uint32_t n, uint64_t d) { return gram_layout(n, d, nullptr, nullptr, false); }
size_t kam_gram_workspace_bytes_int(uint32_t n, uint64_t d) { return gram_layout(n, d, nullptr, nullptr, true); }
uint64_t kam_gram_dpad(uint64_t d) { return dpad_of(d); }

int kam_gram_prepare(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, int operand_kind,
                     void *d_operand, uint64_t *d_S, uint64_t *d_G, uint64_t *d_flags, void *stream) {
    KAM_REQUIRE(d_x && d_operand && d_S && d_G && d_flags, "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32 || elem == KAM_F32 || elem == KAM_F64,
                "euclid/cosine/pearson take count vectors (uint8/16/32) or frequency vectors (float32/64)");
    KAM_REQUIRE(operand_kind == KAM_GRAM_U8 || operand_kind == KAM_GRAM_F16, "bad operand kind");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    KAM_REQUIRE(((uintptr_t)d_operand & 15) == 0, "operand must be 16-byte aligned");
    if (n == 0) return KAM_OK;
    return prepass_dispatch(d_x, elem, n, d, ld, operand_kind == KAM_GRAM_U8, d_operand, (unsigned long long *)d_S,
                            (unsigned long long *)d_G, (unsigned long long *)d_flags, (cudaStream_t)stream);
}

size_t kam_gram_block_workspace_bytes(uint32_t n_a, uint32_t n_b) {
    return kam_align_up((size_t)n_a * sizeof(RowK), 1024) + kam_align_up((size_t)n_b * sizeof(RowK), 1024) +
           kam_align_up(max_tiles(n_a, n_b) * sizeof(uint2), 1024);
}

static int gram_block_impl(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_G_a, uint32_t n_a, const void *d_op_b,
                   const uint64_t *d_S_b, const uint64_t *d_G_b, uint32_t n_b, int operand_kind, uint64_t d,
                   int triangular, uint32_t row_begin, uint32_t row_end, unsigned metrics, int euclid_on_freq,
                   float *d_out_euclid, float *d_out_cosine, float *d_out_pearson, uint64_t out_ld, int sm_limit,
                   void *d_ws, size_t ws_bytes, void *stream, const uint32_t *segs, uint32_t n_segs,
                           const uint32_t *d_ready, uint32_t ready_tiles) {
    KAM_REQUIRE(d_op_a && d_op_b && d_S_a && d_G_a && d_S_b && d_G_b && d_ws, "null pointer");
    KAM_REQUIRE(operand_kind == KAM_GRAM_U8 || operand_kind == KAM_GRAM_F16, "bad operand kind");
    KAM_REQUIRE((metrics & ~(KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON)) == 0 && metrics != 0, "bad metrics mask");
    KAM_REQUIRE(!(metrics & KAM_M_EUCLID) || d_out_euclid, "null euclid output");
    KAM_REQUIRE(!(metrics & KAM_M_COSINE) || d_out_cosine, "null cosine output");
    KAM_REQUIRE(!(metrics & KAM_M_PEARSON) || d_out_pearson, "null pearson output");
    KAM_REQUIRE(!triangular || (d_op_a == d_op_b && (segs ? n_a <= n_b : n_a == n_b)), "a triangular block needs A == B");
    KAM_REQUIRE(triangular || out_ld >= n_b, "dense output needs out_ld >= n_b");
    KAM_REQUIRE(out_ld == 0 || out_ld >= n_b, "out_ld smaller than n_b");
    if (n_a == 0 || n_b == 0 || row_begin >= row_end) return KAM_OK;
    if (row_end > n_a) row_end = n_a;
    if (ws_bytes < kam_gram_block_workspace_bytes(n_a, n_b)) {
        kam_set_error("kam_gram_block: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RowK *rk_a = (RowK *)d_ws;
    RowK *rk_b = (RowK *)((char *)d_ws + kam_align_up((size_t)n_a * sizeof(RowK), 1024));
    rowk_kernel<<<(n_a + 255) / 256, 256, 0, st>>>((const unsigned long long *)d_S_a, (const unsigned long long *)d_G_a,
                                                   n_a, (double)d, euclid_on_freq, rk_a);
    if (d_op_b != d_op_a || d_S_b != d_S_a || n_b != n_a)
        rowk_kernel<<<(n_b + 255) / 256, 256, 0, st>>>((const unsigned long long *)d_S_b,
                                                       (const unsigned long long *)d_G_b, n_b, (double)d,
                                                       euclid_on_freq, rk_b);
    else
        rk_b = rk_a;
    KAM_CUDA(cudaGetLastError());
    GramOut go;
    go.euclid = (metrics & KAM_M_EUCLID) ? d_out_euclid : nullptr;
    go.cosine = (metrics & KAM_M_COSINE) ? d_out_cosine : nullptr;
    go.pearson = (metrics & KAM_M_PEARSON) ? d_out_pearson : nullptr;
    GramGeom gg;
    gg.n_a = n_a;
    gg.n_b = n_b;
    gg.row_begin = row_begin;
    gg.row_end = row_end;
    gg.tri = triangular ? 1u : 0u;
    gg.ld = out_ld;
    gg.rk_a = rk_a;
    gg.rk_b = rk_b;
    gg.ready = d_ready;
    gg.ready_tiles = ready_tiles ? ready_tiles : 1u;
    gg.wait_probe = g_wait_probe;
    return gram_block(d_op_a, d_op_b, operand_kind == KAM_GRAM_U8, d, gg, go, sm_limit > 0 ? (uint32_t)sm_limit : 0u, st,
                      0, 0, segs, n_segs);
}

int kam_gram_block(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_G_a, uint32_t n_a, const void *d_op_b,
                   const uint64_t *d_S_b, const uint64_t *d_G_b, uint32_t n_b, int operand_kind, uint64_t d,
                   int triangular, uint32_t row_begin, uint32_t row_end, unsigned metrics, int euclid_on_freq,
                   float *d_out_euclid, float *d_out_cosine, float *d_out_pearson, uint64_t out_ld, int sm_limit,
                   void *d_ws, size_t ws_bytes, void *stream) {
    return gram_block_impl(d_op_a, d_S_a, d_G_a, n_a, d_op_b, d_S_b, d_G_b, n_b, operand_kind, d, triangular, row_begin,
                           row_end, metrics, euclid_on_freq, d_out_euclid, d_out_cosine, d_out_pearson, out_ld,
                           sm_limit, d_ws, ws_bytes, stream, nullptr, 0, nullptr, 1);
}

int kam_gram_fused(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_G_a, uint32_t n_a, const void *d_op_b,
                   const uint64_t *d_S_b, const uint64_t *d_G_b, uint32_t n_b, int operand_kind, uint64_t d,
                   int triangular_head, const uint32_t *h_segments, uint32_t n_segments, const uint32_t *d_ready,
                   uint32_t ready_tiles, unsigned metrics, int euclid_on_freq, float *d_out_euclid,
                   float *d_out_cosine, float *d_out_pearson, uint64_t out_ld, int sm_limit, void *d_ws,
                   size_t ws_bytes, void *stream) {
    KAM_REQUIRE(h_segments && n_segments >= 1 && n_segments <= 64, "kam_gram_fused: 1..64 segments");
    KAM_REQUIRE(out_ld >= n_b, "kam_gram_fused writes dense rows: out_ld >= n_b");
    KAM_REQUIRE(n_a <= 65535u * 128u, "kam_gram_fused: at most 65535 row tiles");
    KAM_REQUIRE(!d_ready || (n_b + BN - 1) / BN / (ready_tiles ? ready_tiles : 1u) <= 65535u, "kam_gram_fused: at most 65536 chunks");
    KAM_REQUIRE(!triangular_head || d_op_a == d_op_b, "a triangular head needs B to start with the rows of A");
    for (uint32_t i = 0; i < n_segments; ++i)
        KAM_REQUIRE(h_segments[4 * i] <= h_segments[4 * i + 1] && h_segments[4 * i + 2] <= h_segments[4 * i + 3],
                    "kam_gram_fused: a segment is (col tile begin <= end, row tile begin <= end)");
    return gram_block_impl(d_op_a, d_S_a, d_G_a, n_a, d_op_b, d_S_b, d_G_b, n_b, operand_kind, d, triangular_head, 0,
                           n_a, metrics, euclid_on_freq, d_out_euclid, d_out_cosine, d_out_pearson, out_ld, sm_limit,
                           d_ws, ws_bytes, stream, h_segments, n_segments, d_ready, ready_tiles);
}

// ---- manhat / info across ranks: the pieces of the tensor path as entry points (parallel.thermo_ring) ----------
int kam_dist_colmax(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *d_colmax, void *stream) {
    KAM_REQUIRE(d_colmax && (d_x || n == 0), "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32, "the tensor path takes uint8/16/32 counts");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    return kam_thermo_colmax(d_x, elem, n, d, ld, d_colmax, (cudaStream_t)stream);
}

int kam_dist_layout(const uint32_t *d_colmax, uint64_t d, uint32_t *d_off, uint64_t *d_stats, uint64_t *h_stats,
                    void *stream) {
    KAM_REQUIRE(d_colmax && d_off && d_stats && h_stats, "null pointer");
    unsigned long long hs[2] = {0, 0};
    int rc = kam_thermo_layout(d_colmax, d, d_off, (unsigned long long *)d_stats, hs, (cudaStream_t)stream);
    h_stats[0] = hs[0];
    h_stats[1] = hs[1];
    return rc;
}

uint64_t kam_dist_kpad(uint64_t k_total) { return kam_thermo_kpad(k_total); }

int kam_dist_encode(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *d_off, uint64_t kpad,
                    void *d_operand, uint64_t *d_S, uint64_t *d_NZ, void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_operand && d_S && d_NZ && d_ws && (d_x || n == 0), "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32, "the tensor path takes uint8/16/32 counts");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    KAM_REQUIRE(kpad % BK_BYTES == 0 && kpad >= (d_off ? (uint64_t)BK_BYTES : dpad_of(d)), "bad kpad");
    KAM_REQUIRE(((uintptr_t)d_operand & 127) == 0, "operand must be 128-byte aligned");
    if (ws_bytes < (size_t)n * sizeof(RowK)) {
        kam_set_error("kam_dist_encode: workspace too small (48 bytes per row)");
        return KAM_E_WORKSPACE;
    }
    if (n == 0) return KAM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    RowK *rk = (RowK *)d_ws;
    int rc = d_off ? thermo_encode_to(d_x, elem, n, d, ld, d_off, kpad, (uint8_t *)d_operand, rk, st)
                   : thermo_presence_to(d_x, elem, n, d, ld, (uint8_t *)d_operand, rk, st);
    if (rc) return rc;
    if (!d_off && kpad != dpad_of(d)) {
        kam_set_error("kam_dist_encode: presence rows are dpad bytes long");
        return KAM_E_INVALID;
    }
    thermo_stats_from_rowk<<<(n + 255) / 256, 256, 0, st>>>(rk, n, (unsigned long long *)d_S, (unsigned long long *)d_NZ);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

int kam_thermo_fused(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_NZ_a, uint32_t n_a, const void *d_op_b,
                     const uint64_t *d_S_b, const uint64_t *d_NZ_b, uint32_t n_b, uint64_t d, uint64_t kpad,
                     int triangular_head, const uint32_t *h_segments, uint32_t n_segments, const uint32_t *d_ready,
                     uint32_t ready_tiles, unsigned metric, void *d_out, uint64_t out_ld, void *d_ws, size_t ws_bytes,
                     void *stream) {
    KAM_REQUIRE(d_op_a && d_op_b && d_S_a && d_NZ_a && d_S_b && d_NZ_b && d_ws && d_out, "null pointer");
    KAM_REQUIRE(metric == KAM_M_MANHAT || metric == KAM_M_INFO, "kam_thermo_fused computes manhat or info");
    KAM_REQUIRE(h_segments && n_segments >= 1 && n_segments <= 64, "kam_thermo_fused: 1..64 segments");
    KAM_REQUIRE(out_ld >= n_b, "kam_thermo_fused writes dense rows: out_ld >= n_b");
    KAM_REQUIRE(!triangular_head || (d_op_a == d_op_b && n_a <= n_b), "a triangular head needs B to start with the rows of A");
    KAM_REQUIRE(kpad % BK_BYTES == 0 && kpad >= (uint64_t)BK_BYTES && kpad < (1ull << 31), "bad kpad");
    KAM_REQUIRE(n_a <= 65535u * 128u, "kam_thermo_fused: at most 65535 row tiles");
    KAM_REQUIRE(!d_ready || (n_b + BN - 1) / BN / (ready_tiles ? ready_tiles : 1u) <= 65535u, "kam_thermo_fused: at most 65536 chunks");
    if (n_a == 0 || n_b == 0) return KAM_OK;
    if (ws_bytes < kam_gram_block_workspace_bytes(n_a, n_b)) {
        kam_set_error("kam_thermo_fused: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RowK *rk_a = (RowK *)d_ws;
    RowK *rk_b = (RowK *)((char *)d_ws + kam_align_up((size_t)n_a * sizeof(RowK), 1024));
    thermo_rowk_from_stats<<<(n_a + 255) / 256, 256, 0, st>>>((const unsigned long long *)d_S_a,
                                                              (const unsigned long long *)d_NZ_a, n_a, rk_a);
    thermo_rowk_from_stats<<<(n_b + 255) / 256, 256, 0, st>>>((const unsigned long long *)d_S_b,
                                                              (const unsigned long long *)d_NZ_b, n_b, rk_b);
    KAM_CUDA(cudaGetLastError());
    GramOut go;
    go.euclid = go.cosine = go.pearson = nullptr;
    if (metric == KAM_M_MANHAT) go.manhat = (uint32_t *)d_out;
    else go.info = (float *)d_out;
    GramGeom gg;
    gg.n_a = n_a;
    gg.n_b = n_b;
    gg.row_begin = 0;
    gg.row_end = n_a;
    gg.tri = triangular_head ? 1u : 0u;
    gg.ld = out_ld;
    gg.rk_a = rk_a;
    gg.rk_b = rk_b;
    gg.ready = d_ready;
    gg.ready_tiles = ready_tiles ? ready_tiles : 1u;
    gg.wait_probe = g_wait_probe;
    return gram_block(d_op_a, d_op_b, true, d, gg, go, 0, st, kpad, kpad, h_segments, n_segments);
}

int kam_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t row_begin, uint32_t row_end,
                 unsigned metrics, int euclid_on_freq, float *d_tri_euclid, float *d_tri_cosine, float *d_tri_pearson,
                 void *d_ws, size_t ws_bytes, void *stream) {
    KAM_REQUIRE(d_x && d_ws, "null pointer");
    KAM_REQUIRE(elem == KAM_U8 || elem == KAM_U16 || elem == KAM_U32 || elem == KAM_F32 || elem == KAM_F64,
                "euclid/cosine/pearson take count vectors (uint8/16/32) or frequency vectors (float32/64)");
    KAM_REQUIRE(d >= 1 && d < (1ull << 32) && ld >= d, "bad d / ld");
    KAM_REQUIRE((metrics & ~(KAM_M_EUCLID | KAM_M_COSINE | KAM_M_PEARSON)) == 0 && metrics != 0, "bad metrics mask");
    KAM_REQUIRE(!(metrics & KAM_M_EUCLID) || d_tri_euclid, "null euclid output");
    KAM_REQUIRE(!(metrics & KAM_M_COSINE) || d_tri_cosine, "null cosine output");
    KAM_REQUIRE(!(metrics & KAM_M_PEARSON) || d_tri_pearson, "null pearson output");
    if (n < 2 || row_begin >= row_end) return KAM_OK;
    if (row_end > n) row_end = n;
    if (ws_bytes < kam_gram_workspace_bytes(n, d)) {
        kam_set_error("kam_gram_tri: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    GramWs w;
    gram_layout(n, d, d_ws, &w, false);
    GramOut go;
    go.euclid = (metrics & KAM_M_EUCLID) ? d_tri_euclid : nullptr;
    go.cosine = (metrics & KAM_M_COSINE) ? d_tri_cosine : nullptr;
    go.pearson = (metrics & KAM_M_PEARSON) ? d_tri_pearson : nullptr;

    // pre-pass: operand rows + exact row statistics + the envelope flags in one sweep.  Operand type:
    // uint8 x uint8 -> int32 (kind::i8) when every count fits a byte; fp16 -> fp32 (kind::f16) up to
    // 2048; otherwise the exact integer Gram on the CUDA cores.  The uint8 operand is written
    // speculatively (the common case); a second sweep rewrites it as fp16 only when it did not fit.
    // Floating-point rows (a `frequencies` file): the same sweep recovers the integer counts behind the
    // quotients, row by row, and accepts them only when they reproduce every element bit for bit
    // (freq_gram.cu); euclid is then the distance between the rows as stored, i.e. between frequency
    // vectors.  Rows that are not frequency vectors send the whole matrix to the float64 CUDA-core Gram.
    const bool is_float = elem == KAM_F32 || elem == KAM_F64;
    if (is_float) euclid_on_freq = 1;
    unsigned long long hflags[3] = {0, 0, 0};
    int rc;
    bool have_i8 = false;
    if (is_float && !g_float_recover) return kam_float_gram_tri(d_x, elem, n, d, ld, row_begin, row_end, w.rk, go, st);
    if (!g_force_f16) {
        KAM_CUDA(cudaMemsetAsync(w.flags, 0, 64, st));
        if ((rc = prepass_dispatch(d_x, elem, n, d, ld, true, w.h, w.S, w.G, w.flags, st))) return rc;
        if ((rc = read_flags(w.flags, hflags, st))) return rc;
        if (is_float && hflags[2] != 0)   // not frequency vectors: no second sweep, straight to the float64 Gram
            return kam_float_gram_tri(d_x, elem, n, d, ld, row_begin, row_end, w.rk, go, st);
        have_i8 = hflags[0] <= MAX_I8_COUNT && hflags[1] < MAX_I8_GII;
    }
    if (!have_i8 && (g_force_f16 || (hflags[0] <= MAX_F16_COUNT && hflags[1] < MAX_F16_GII))) {
        KAM_CUDA(cudaMemsetAsync(w.flags, 0, 64, st));
        if ((rc = prepass_dispatch(d_x, elem, n, d, ld, false, w.h, w.S, w.G, w.flags, st))) return rc;
        if ((rc = read_flags(w.flags, hflags, st))) return rc;
    }
    const bool use_i8 = have_i8;
    const bool use_f16 = !use_i8 && hflags[0] <= MAX_F16_COUNT && hflags[1] < MAX_F16_GII;
    if (is_float && (hflags[2] != 0 || (!use_i8 && !use_f16)))
        return kam_float_gram_tri(d_x, elem, n, d, ld, row_begin, row_end, w.rk, go, st);
    rowk_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.S, w.G, n, (double)d, euclid_on_freq, w.rk);
    KAM_CUDA(cudaGetLastError());
    if (!use_i8 && !use_f16) {
        if (ws_bytes < kam_gram_workspace_bytes_int(n, d)) {
            kam_set_error("kam_gram_tri: counts outside the tensor-exact envelope (max count %llu, max sum of squares "
                          "%llu) need the integer path: pass a workspace of kam_gram_workspace_bytes_int()",
                          hflags[0], hflags[1]);
            return KAM_E_WORKSPACE;
        }
        // counts up to 65535 whose byte limbs keep every int32 partial Gram exact: limb-split tensor path
        if (g_limb && !is_float && hflags[0] <= 65535) {
            uint8_t *xa = (uint8_t *)w.h, *xb = xa + (size_t)n * dpad_of(d) * 2;
            int *scratch = (int *)((char *)w.h + kam_align_up((size_t)n * dpad_of(d) * 4, 1024));
            KAM_CUDA(cudaMemsetAsync(w.flags, 0, 64, st));
            if (elem == KAM_U8)
                gram_prepass_limb_kernel<uint8_t><<<n, 256, 0, st>>>((const uint8_t *)d_x, n, d, ld, dpad_of(d), xa, xb, w.S, w.G, w.flags);
            else if (elem == KAM_U16)
                gram_prepass_limb_kernel<uint16_t><<<n, 256, 0, st>>>((const uint16_t *)d_x, n, d, ld, dpad_of(d), xa, xb, w.S, w.G, w.flags);
            else
                gram_prepass_limb_kernel<uint32_t><<<n, 256, 0, st>>>((const uint32_t *)d_x, n, d, ld, dpad_of(d), xa, xb, w.S, w.G, w.flags);
            KAM_CUDA(cudaGetLastError());
            unsigned long long lf[5] = {0, 0, 0, 0, 0};
            KAM_CUDA(cudaMemcpyAsync(lf, w.flags, 40, cudaMemcpyDeviceToHost, st));
            KAM_CUDA(cudaStreamSynchronize(st));
            const double lim = 2147483647.0;
            if ((double)lf[3] <= lim && (double)lf[4] <= lim && 2.0 * sqrt((double)lf[3] * (double)lf[4]) <= lim) {
                const uint64_t dp = dpad_of(d);
                const uint32_t R = limb_rows(n);
                int *sll = scratch, *smid = scratch + (size_t)R * n, *shh = scratch + (size_t)2 * R * n;
                for (uint32_t rb = row_begin / 256 * 256; rb < row_end; rb += R) {
                    GramGeom gg;
                    gg.n_a = gg.n_b = n;
                    gg.row_begin = rb > row_begin ? rb : row_begin;
                    gg.row_end = rb + R < row_end ? rb + R : row_end;
                    gg.tri = 1;
                    gg.ld = n;
                    gg.rk_a = gg.rk_b = w.rk;
                    const size_t shift = (size_t)gg.row_begin * n;     // scratch row 0 = row gg.row_begin
                    GramOut gr;
                    gr.euclid = gr.cosine = gr.pearson = nullptr;
                    gr.raw = sll - shift;
                    if ((rc = gram_block(xa, xa, true, d, gg, gr, 0, st, dp, 2 * dp))) return rc;
                    gr.raw = shh - shift;
                    if ((rc = gram_block(xa + dp, xa + dp, true, d, gg, gr, 0, st, dp, 2 * dp))) return rc;
                    gr.raw = smid - shift;
                    if ((rc = gram_block(xa, xb, true, d, gg, gr, 0, st, 2 * dp, 2 * dp))) return rc;
                    gram_limb_combine_kernel<<<dim3((n + 255) / 256, gg.row_end - gg.row_begin), 256, 0, st>>>(
                        sll, smid, shh, n, gg.row_begin, gg.row_end, w.rk, go);
                    KAM_CUDA(cudaGetLastError());
                }
                return KAM_OK;
            }
        }
        return kam_int_gram_tri(d_x, elem, n, d, ld, row_begin, row_end, w.rk, go, w.xt, st);
    }
    GramGeom gg;
    gg.n_a = gg.n_b = n;
    gg.row_begin = row_begin;
    gg.row_end = row_end;
    gg.tri = 1;
    gg.ld = 0;
    gg.rk_a = gg.rk_b = w.rk;
    return gram_block(w.h, w.h, use_i8, d, gg, go, 0, st);
}

// test / tuning hooks
void kam_gram_force_f16(int on) { g_force_f16 = on != 0; }
void kam_gram_set_float_recover(int on) { g_float_recover = on != 0; }
void kam_gram_set_cluster(int cl) { g_cluster = cl == 1 ? 1u : 2u; }
void kam_gram_set_pair(int on) { g_pair = on != 0; }
void kam_gram_set_limb(int on) { g_limb = on != 0; }
void kam_gram_set_wait_probe(uint64_t *d_three_words) { g_wait_probe = (unsigned long long *)d_three_words; }
void kam_gram_set_strip(int row_tiles) { g_strip = row_tiles >= 2 ? (uint32_t)row_tiles : 2u; }

}  // extern "C"
