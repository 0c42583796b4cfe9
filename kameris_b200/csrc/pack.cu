// pack.cu -- ASCII records -> flat 2-bit "planes" stream (see include/kameris_b200.h).
//
// Replaces the character classification of the reference CGR loop
// (generation_cgr_mac_avx2 @0x10000e22e-0x10000e248): only 'A','C','G','T' are bases,
// everything else is skipped without resetting the k-mer window, so removing the other
// characters up front is exact -- except for the raw-index emit test (quirk Q1), which
// head_mask preserves.
//
// Three launches over tiles of 4096 characters:
//   pack_count_kernel   classifies every character ONCE, four at a time inside a 32-bit word (a
//                       byte-permute perfect hash gives the letter each byte would have to be; an exact
//                       zero-byte test and three multiply "movemasks" give 4-bit valid / x / y masks),
//                       stores the three 16-bit masks of each thread's 16 characters (8 B) and the
//                       valid bases per tile
//   pack_scan_kernel    exclusive scan of the tile counts (one CTA), grand total
//   pack_scatter_kernel compaction from the stored masks (the characters are not read again): x/y
//                       bits land at their flat position; per-sequence starts
// HBM-bound byte work: 16-byte loads, one uint2 store per 32 bases.
#include "common.cuh"

namespace {

constexpr int PACK_THREADS = 256;
constexpr int PACK_CPT = 16;                         // characters per thread
constexpr int PACK_TILE = PACK_THREADS * PACK_CPT;   // 4096
constexpr int SCAN_THREADS = 1024;

// bit (c & 31) of these masks, for c in 0x40..0x5f:  A=0x41 C=0x43 G=0x47 T=0x54
constexpr uint32_t LUT_VALID = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);
constexpr uint32_t LUT_X = (1u << 7) | (1u << 20);   // G, T
constexpr uint32_t LUT_Y = (1u << 1) | (1u << 20);   // A, T

struct Cls {
    uint32_t cnt, x, y, vmask;   // compacted x/y bits (cnt of them), validity of the 16 raw chars
};

// 4 characters in one word -> 4-bit masks (bit i = character i): valid, x-bit (G,T), y-bit (A,T).
// (c >> 1) & 3 is a perfect hash of the four letters (A:0 C:1 T:2 G:3); a byte permute looks up the
// letter each byte would have to be, and the byte is a base iff it equals that letter.  For bases,
// x = bit 2 of the character and y = the complement of bit 1.
__device__ __forceinline__ void cls_word(uint32_t w, uint32_t &v4, uint32_t &x4, uint32_t &y4) {
    const uint32_t h4 = (w >> 1) & 0x03030303u;
    const uint32_t t = h4 | (h4 >> 4);                                  // byte 0: h0 | h1<<4, byte 2: h2 | h3<<4
    const uint32_t expect = __byte_perm(0x47544341u, 0u, __byte_perm(t, 0u, 0x4420u));
    const uint32_t z = w ^ expect;
    const uint32_t nz = ((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z;          // bit 7 of each byte: byte != 0 (no cross-byte carry)
    const uint32_t v1 = (~nz >> 7) & 0x01010101u;                       // bit 0 of each byte: valid
    constexpr uint32_t GATHER = 0x01020408u;                            // bits 0, 8, 16, 24 -> bits 24..27
    v4 = (v1 * GATHER) >> 24;
    x4 = (((w >> 2) & v1) * GATHER) >> 24;
    y4 = (((~w >> 1) & v1) * GATHER) >> 24;
}

__device__ __forceinline__ void cls_byte(uint32_t c, uint32_t &vm, uint32_t &xm, uint32_t &ym, int i) {
    const uint32_t sel = c & 31u;
    const uint32_t v = ((c & 0xE0u) == 0x40u) ? ((LUT_VALID >> sel) & 1u) : 0u;
    vm |= v << i;
    xm |= (v & (LUT_X >> sel)) << i;
    ym |= (v & (LUT_Y >> sel)) << i;
}

// masks of the 16 characters at chars[pos .. pos+16) (those at or beyond `total` are invalid):
// .x = valid | x << 16, .y = y
__device__ __forceinline__ uint2 classify16(const uint8_t *__restrict__ chars, uint64_t pos, uint64_t total) {
    uint32_t vm = 0, xm = 0, ym = 0;
    if (pos + PACK_CPT <= total) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(chars + pos));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t v4, x4, y4;
            cls_word(w[q], v4, x4, y4);
            vm |= v4 << (4 * q);
            xm |= x4 << (4 * q);
            ym |= y4 << (4 * q);
        }
    } else {
        for (int i = 0; i < PACK_CPT; ++i)
            if (pos + i < total) cls_byte(chars[pos + i], vm, xm, ym, i);
    }
    return make_uint2(vm | (xm << 16), ym);
}

// stored masks -> compacted x/y bits: drop the bit positions of the non-base characters (usually none,
// or one IUPAC symbol), from the top so that the positions below stay put
__device__ __forceinline__ Cls compact16(uint2 m) {
    Cls r;
    r.vmask = m.x & 0xffffu;
    r.x = m.x >> 16;
    r.y = m.y;
    r.cnt = __popc(r.vmask);
    uint32_t inv = ~r.vmask & 0xffffu;
    while (inv) {
        const uint32_t p = 31u - (uint32_t)__clz((int)inv);
        const uint32_t low = (1u << p) - 1u;
        r.x = (r.x & low) | ((r.x >> (p + 1)) << p);
        r.y = (r.y & low) | ((r.y >> (p + 1)) << p);
        inv ^= 1u << p;
    }
    return r;
}

__global__ void __launch_bounds__(PACK_THREADS) pack_count_kernel(const uint8_t *__restrict__ chars,
                                                                  uint64_t total, uint32_t *__restrict__ tile_count,
                                                                  uint2 *__restrict__ masks) {
    __shared__ uint32_t wsum[PACK_THREADS / 32];
    const uint64_t pos = (uint64_t)blockIdx.x * PACK_TILE + (uint64_t)threadIdx.x * PACK_CPT;
    const uint2 m = pos < total ? classify16(chars, pos, total) : make_uint2(0, 0);
    masks[(uint64_t)blockIdx.x * PACK_THREADS + threadIdx.x] = m;
    uint32_t c = __popc(m.x & 0xffffu);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < PACK_THREADS / 32; ++i) s += wsum[i];
        tile_count[blockIdx.x] = s;
    }
}

// one CTA: tile_off[t] = sum of tile_count[0..t) (here: over the chunk totals of the two-level scan below);
// base_start[s] = grand total for every trailing sequence that starts at `total` (they own no tile);
// base_start[n_seqs] likewise.  The counts are walked in chunks of 1024 consecutive entries (coalesced
// loads and stores), each chunk scanned with warp shuffles + one shared-memory pass over the 32 warp
// totals, the running total carried along.
__global__ void __launch_bounds__(SCAN_THREADS) pack_scan_kernel(const uint32_t *__restrict__ tile_count,
                                                                 uint64_t n_tiles, uint64_t *__restrict__ tile_off,
                                                                 const uint64_t *__restrict__ char_start,
                                                                 uint32_t n_seqs, uint64_t total,
                                                                 uint64_t *__restrict__ base_start) {
    __shared__ unsigned long long wtot[SCAN_THREADS / 32];
    __shared__ unsigned long long carry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t c0 = 0; c0 < n_tiles; c0 += SCAN_THREADS) {
        const uint64_t i = c0 + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_count[i] : 0u;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += u;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        unsigned long long woff = 0, chunk_total = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            const unsigned long long t = wtot[w];
            if (w < (int)warp) woff += t;
            chunk_total += t;
        }
        const unsigned long long base = carry;
        if (i < n_tiles) tile_off[i] = base + woff + inc - v;
        __syncthreads();                                           // everyone has read wtot and carry
        if (threadIdx.x == 0) carry = base + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long grand = carry;
        tile_off[n_tiles] = grand;
        base_start[n_seqs] = grand;
        for (int64_t q = (int64_t)n_seqs - 1; q >= 0 && char_start[q] >= total; --q) base_start[q] = grand;
    }
}

// two-level scan of the tile counts: chunk totals (one CTA per 1024 tiles), pack_scan_kernel over the chunk
// totals, then every chunk scans its own 1024 tiles on top of its chunk offset -- three short launches
// instead of one CTA walking all the tiles.
__global__ void __launch_bounds__(SCAN_THREADS) chunk_sum_kernel(const uint32_t *__restrict__ tile_count,
                                                                 uint64_t n_tiles, uint32_t *__restrict__ chunk_tot) {
    __shared__ uint32_t wtot[SCAN_THREADS / 32];
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    uint32_t v = warp_sum(i < n_tiles ? tile_count[i] : 0u);
    if ((threadIdx.x & 31) == 0) wtot[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = warp_sum(wtot[threadIdx.x]);
        if (threadIdx.x == 0) chunk_tot[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) tile_off_kernel(const uint32_t *__restrict__ tile_count,
                                                                uint64_t n_tiles,
                                                                const uint64_t *__restrict__ chunk_off,
                                                                uint64_t *__restrict__ tile_off) {
    __shared__ uint32_t wtot[SCAN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const uint32_t v = i < n_tiles ? tile_count[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (uint32_t)o) inc += u;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w)
        if (w < (int)warp) woff += wtot[w];
    if (i < n_tiles) tile_off[i] = chunk_off[blockIdx.x] + woff + inc - v;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) tile_off[n_tiles] = chunk_off[gridDim.x];
}

__global__ void __launch_bounds__(PACK_THREADS) pack_scatter_kernel(const uint2 *__restrict__ masks, uint64_t total,
                                                                    const uint64_t *__restrict__ tile_off,
                                                                    const uint64_t *__restrict__ char_start,
                                                                    uint32_t n_seqs, uint2 *__restrict__ planes,
                                                                    uint64_t *__restrict__ base_start,
                                                                    const uint32_t *__restrict__ tile_first_seq) {
    constexpr int NW = PACK_THREADS / 32;
    constexpr int SW = PACK_TILE / 32 + 2;   // words a tile's compacted bases can touch
    __shared__ uint32_t sx[SW], sy[SW];
    __shared__ uint32_t wsum[NW];
    __shared__ uint32_t pref[PACK_THREADS];
    __shared__ uint32_t vmask[PACK_THREADS];

    const uint64_t tile_begin = (uint64_t)blockIdx.x * PACK_TILE;
    const uint64_t tile_end = tile_begin + PACK_TILE < total ? tile_begin + PACK_TILE : total;
    const uint64_t goff = tile_off[blockIdx.x];            // flat base index of this tile's first base
    const uint32_t bit0 = (uint32_t)(goff & 31);
    const uint64_t word0 = goff >> 5;

    for (int i = threadIdx.x; i < SW; i += PACK_THREADS) {
        sx[i] = 0;
        sy[i] = 0;
    }
    const Cls r = compact16(__ldg(masks + (uint64_t)blockIdx.x * PACK_THREADS + threadIdx.x));

    // block exclusive scan of r.cnt
    uint32_t inc = r.cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += v;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        if (w < (int)(threadIdx.x >> 5)) woff += wsum[w];
    const uint32_t excl = woff + inc - r.cnt;
    pref[threadIdx.x] = excl;
    vmask[threadIdx.x] = r.vmask;

    if (r.cnt) {
        const uint32_t bitpos = bit0 + excl;
        const uint32_t lw = bitpos >> 5, sh = bitpos & 31;
        atomicOr(&sx[lw], r.x << sh);
        atomicOr(&sy[lw], r.y << sh);
        if (sh + r.cnt > 32) {
            atomicOr(&sx[lw + 1], r.x >> (32 - sh));
            atomicOr(&sy[lw + 1], r.y >> (32 - sh));
        }
    }
    __syncthreads();

    const uint32_t tile_cnt = wsum[0] + wsum[1] + wsum[2] + wsum[3] + wsum[4] + wsum[5] + wsum[6] + wsum[7];
    static_assert(NW == 8, "tile_cnt sum assumes 8 warps");
    const uint32_t nwords = (bit0 + tile_cnt + 31) >> 5;   // words touched (0 when the tile has no base)
    for (uint32_t i = threadIdx.x; i < nwords; i += PACK_THREADS) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(planes + word0 + i);
        if (i == 0 || i == nwords - 1) {       // may be shared with a neighbouring tile (planes pre-zeroed)
            if (sx[i]) atomicOr(dst, sx[i]);
            if (sy[i]) atomicOr(dst + 1, sy[i]);
        } else {
            planes[word0 + i] = make_uint2(sx[i], sy[i]);
        }
    }

    // sequences that start inside this tile
    for (uint32_t s = tile_first_seq[blockIdx.x] + threadIdx.x; s < n_seqs; s += PACK_THREADS) {
        const uint64_t c0 = char_start[s];
        if (c0 >= tile_end) break;             // char_start is non-decreasing
        const uint32_t rel = (uint32_t)(c0 - tile_begin);
        const uint32_t th = rel / PACK_CPT, o = rel % PACK_CPT;
        base_start[s] = goff + pref[th] + __popc(vmask[th] & ((1u << o) - 1u));
    }
}


// ---- single pass: classify, chained scan with decoupled look-back, compaction --------------------------
// The three-launch version above writes 8 bytes of masks per 16 characters and reads them back; here a
// tile is classified once and scattered by the same CTA: 1 byte read + 1/4 byte written per character.
// Chunks of 8 tiles (32 768 characters) are handed out by an atomic ticket (so every lower-numbered chunk
// is already running).  A CTA classifies its chunk (8 independent 16-byte loads per thread, masks kept
// in 16 KB of shared memory), publishes the chunk's base count, then warp 0 walks back over the published
// words -- 32 predecessors per step -- summing `aggregate` entries until it meets an `inclusive prefix`
// entry, and publishes its own inclusive prefix.  The chunk's tiles are then compacted one after the
// other: x/y bit streams built in shared memory at bit offset 0, funnel-shifted to the tile's flat
// position on the way out.  state word: bits 63..62 = 0 empty / 1 aggregate / 2 inclusive prefix,
// bits 61..0 = bases.
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_VAL = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int PF_SUB = 8;                               // tiles per CTA ("chunk": 32 768 characters)

__global__ void __launch_bounds__(PACK_THREADS) pack_fused_kernel(const uint8_t *__restrict__ chars, uint64_t total,
                                                                  uint64_t n_tiles, const uint64_t *__restrict__ char_start,
                                                                  uint32_t n_seqs, uint2 *__restrict__ planes,
                                                                  uint64_t *__restrict__ base_start,
                                                                  const uint32_t *__restrict__ tile_first_seq,
                                                                  unsigned long long *__restrict__ state,
                                                                  unsigned int *__restrict__ ticket) {
    constexpr int NW = PACK_THREADS / 32;
    constexpr int SW = PACK_TILE / 32 + 1;   // words of a tile's compacted bases, plus a zero word behind them
    __shared__ uint2 smask[PF_SUB][PACK_THREADS];          // the classification of the whole chunk (16 KB)
    __shared__ uint32_t sx[SW], sy[SW];
    __shared__ uint32_t wcnt[NW][PF_SUB];                  // bases per warp segment (512 characters) and tile
    __shared__ uint32_t sub_cnt[PF_SUB];
    __shared__ uint32_t pref[PACK_THREADS];
    __shared__ uint32_t s_chunk;
    __shared__ unsigned long long s_goff;

    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint64_t chunk = s_chunk;
    const uint64_t n_chunks = (n_tiles + PF_SUB - 1) / PF_SUB;
    const uint64_t tile0 = chunk * PF_SUB;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- phase A: classify the chunk (8 independent 16-byte loads per thread), count its bases
#pragma unroll
    for (int i = 0; i < PF_SUB; ++i) {
        const uint64_t pos = (tile0 + i) * PACK_TILE + (uint64_t)threadIdx.x * PACK_CPT;
        const uint2 m = pos < total ? classify16(chars, pos, total) : make_uint2(0, 0);
        smask[i][threadIdx.x] = m;
        const uint32_t c = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(m.x & 0xffffu));   // REDUX.SUM
        if (lane == 0) wcnt[warp][i] = c;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // ---- per-tile and chunk counts, publication, decoupled look-back (warp 0)
        uint32_t t = 0;
        if (lane < PF_SUB) {
#pragma unroll
            for (int w = 0; w < NW; ++w) t += wcnt[w][lane];
            sub_cnt[lane] = t;
        }
        const uint32_t chunk_cnt = __reduce_add_sync(0xffffffffu, t);
        if (threadIdx.x == 0 && chunk > 0) st_state(&state[chunk], ST_AGG | chunk_cnt);
        unsigned long long before = 0;
        if (chunk > 0) {
            int64_t p = (int64_t)chunk - 1;
            for (uint32_t spin = 0;;) {
                const int64_t idx = p - (int64_t)threadIdx.x;
                const unsigned long long v = idx >= 0 ? ld_state(&state[idx]) : ST_PREFIX;
                const uint32_t flag = (uint32_t)(v >> 62);
                if (__any_sync(0xffffffffu, flag == 0)) {          // a predecessor has not published yet
                    if (++spin > (1u << 24)) __trap();
                    continue;
                }
                const uint32_t isp = __ballot_sync(0xffffffffu, flag == 2);
                const uint32_t upto = isp ? (uint32_t)__ffs((int)isp) - 1u : 31u;   // closest inclusive prefix
                unsigned long long part = threadIdx.x <= upto ? (v & ST_VAL) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                before += part;
                if (isp) break;
                p -= 32;
            }
        }
        if (threadIdx.x == 0) {
            st_state(&state[chunk], ST_PREFIX | (before + chunk_cnt));
            s_goff = before;
        }
    }
    __syncthreads();

    // ---- phase B: compaction of the chunk's tiles from the stored masks
    uint64_t goff = s_goff;                                  // flat base index of the current tile's first base
#pragma unroll 1
    for (int i = 0; i < PF_SUB; ++i) {
        const uint64_t tile = tile0 + i;
        if (tile >= n_tiles) break;
        const uint64_t tile_begin = tile * PACK_TILE;
        const uint64_t tile_end = tile_begin + PACK_TILE < total ? tile_begin + PACK_TILE : total;
        const uint32_t tile_cnt = sub_cnt[i];
        for (int w = threadIdx.x; w < SW; w += PACK_THREADS) {
            sx[w] = 0;
            sy[w] = 0;
        }
        const Cls r = compact16(smask[i][threadIdx.x]);
        uint32_t inc = r.cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (uint32_t)o) inc += v;
        }
        uint32_t woff = 0;                                   // bases of the warps before (phase A counted them)
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (w < (int)warp) woff += wcnt[w][i];
        const uint32_t excl = woff + inc - r.cnt;
        pref[threadIdx.x] = excl;
        __syncthreads();                                     // streams cleared
        if (r.cnt) {                                         // bit streams at offset 0
            const uint32_t lw = excl >> 5, sh = excl & 31;
            atomicOr(&sx[lw], r.x << sh);
            atomicOr(&sy[lw], r.y << sh);
            if (sh + r.cnt > 32) {
                atomicOr(&sx[lw + 1], r.x >> (32 - sh));
                atomicOr(&sy[lw + 1], r.y >> (32 - sh));
            }
        }
        __syncthreads();
        const uint32_t bit0 = (uint32_t)(goff & 31);
        const uint64_t word0 = goff >> 5;
        const uint32_t nwords = (bit0 + tile_cnt + 31) >> 5;   // words touched (0 when the tile has no base)
        for (uint32_t w = threadIdx.x; w < nwords; w += PACK_THREADS) {
            // word w of the output = bits [32 w - bit0, 32 w - bit0 + 32) of the offset-0 stream
            const uint32_t hx = w < (uint32_t)SW ? sx[w] : 0u, hy = w < (uint32_t)SW ? sy[w] : 0u;
            const uint32_t lx = w ? sx[w - 1] : 0u, ly = w ? sy[w - 1] : 0u;
            const uint32_t ox = __funnelshift_l(lx, hx, bit0), oy = __funnelshift_l(ly, hy, bit0);
            uint32_t *dst = reinterpret_cast<uint32_t *>(planes + word0 + w);
            if (w == 0 || w == nwords - 1) {     // may be shared with a neighbouring tile (planes pre-zeroed)
                if (ox) atomicOr(dst, ox);
                if (oy) atomicOr(dst + 1, oy);
            } else {
                planes[word0 + w] = make_uint2(ox, oy);
            }
        }
        // sequences that start inside this tile
        for (uint32_t s = tile_first_seq[tile] + threadIdx.x; s < n_seqs; s += PACK_THREADS) {
            const uint64_t c0 = char_start[s];
            if (c0 >= tile_end) break;           // char_start is non-decreasing
            const uint32_t rel = (uint32_t)(c0 - tile_begin);
            const uint32_t th = rel / PACK_CPT, o = rel % PACK_CPT;
            base_start[s] = goff + pref[th] + __popc(smask[i][th].x & 0xffffu & ((1u << o) - 1u));
        }
        goff += tile_cnt;
        __syncthreads();                                     // pref / streams consumed before the next tile
    }
    if (chunk == n_chunks - 1 && threadIdx.x == 0) {         // grand total: the end marker and trailing empty sequences
        base_start[n_seqs] = goff;
        for (int64_t q = (int64_t)n_seqs - 1; q >= 0 && char_start[q] >= total; --q) base_start[q] = goff;
    }
}

int g_pack_fused = 1;   // test hook: 0 = the three-launch count / scan / scatter version

// tile_first_seq[t] = first sequence whose first character lies at or after the start of tile t
// (lower_bound of char_start): thread s fills the tiles (tile of sequence s-1, tile of sequence s];
// the thread of the last sequence also fills the tiles behind it with n_seqs.
__global__ void tile_first_seq_kernel(const uint64_t *__restrict__ char_start, uint32_t n_seqs, uint64_t n_tiles,
                                      uint32_t *__restrict__ tile_first_seq) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seqs) return;
    const uint64_t c = char_start[s];
    // tiles i with char_start[s-1] < i * TILE <= char_start[s]  (equal starts: the first sequence takes them)
    const uint64_t lo = s ? char_start[s - 1] / PACK_TILE + 1 : 0;
    const uint64_t hi = c / PACK_TILE;
    for (uint64_t i = lo; i <= hi && i < n_tiles; ++i) tile_first_seq[i] = s;
    if (s == n_seqs - 1)
        for (uint64_t i = hi + 1; i < n_tiles; ++i) tile_first_seq[i] = n_seqs;
}

__global__ void head_mask_kernel(const uint8_t *__restrict__ chars, const uint64_t *__restrict__ char_start,
                                 uint32_t n_seqs, uint32_t *__restrict__ head_mask) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seqs) return;
    const uint64_t b = char_start[s], e = char_start[s + 1];
    const uint32_t n = e - b < 32 ? (uint32_t)(e - b) : 32u;
    uint32_t m = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t c = chars[b + i];
        const uint32_t v = ((c & 0xE0u) == 0x40u) ? ((LUT_VALID >> (c & 31u)) & 1u) : 0u;
        m |= v << i;
    }
    head_mask[s] = m;
}

inline uint64_t n_tiles_of(uint64_t total) { return (total + PACK_TILE - 1) / PACK_TILE; }

}  // namespace

extern "C" {

size_t kam_planes_words(uint64_t total_chars) { return (size_t)(total_chars / 32 + 2); }

size_t kam_pack_workspace_bytes(uint64_t total_chars, uint32_t n_seqs) {
    (void)n_seqs;
    const uint64_t nt = n_tiles_of(total_chars);
    return kam_align_up((nt + 1) * sizeof(uint32_t), 256) + kam_align_up((nt + 1) * sizeof(uint64_t), 256) +
           kam_align_up(nt * PACK_THREADS * sizeof(uint2), 256) + kam_align_up((nt + 1) * sizeof(uint32_t), 256) +
           kam_align_up((nt / SCAN_THREADS + 2) * sizeof(uint32_t), 256) +
           kam_align_up((nt / SCAN_THREADS + 2) * sizeof(uint64_t), 256);
}

int kam_pack_ascii(const uint8_t *d_chars, const uint64_t *d_char_start, uint32_t n_seqs, uint64_t total_chars,
                   uint64_t *d_planes, uint64_t *d_base_start, uint32_t *d_head_mask, void *d_ws, size_t ws_bytes,
                   void *stream) {
    KAM_REQUIRE(d_char_start && d_planes && d_base_start && d_head_mask, "null pointer");
    KAM_REQUIRE(total_chars == 0 || d_chars, "null d_chars");
    KAM_REQUIRE(((uintptr_t)d_chars & 15) == 0, "d_chars must be 16-byte aligned");
    if (ws_bytes < kam_pack_workspace_bytes(total_chars, n_seqs)) {
        kam_set_error("kam_pack_ascii: workspace too small");
        return KAM_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nt = n_tiles_of(total_chars);
    KAM_REQUIRE(nt < (1ull << 31), "input too large for one call");
    uint32_t *tile_count = (uint32_t *)d_ws;
    uint64_t *tile_off = (uint64_t *)((char *)d_ws + kam_align_up((nt + 1) * sizeof(uint32_t), 256));
    uint2 *masks = (uint2 *)((char *)tile_off + kam_align_up((nt + 1) * sizeof(uint64_t), 256));
    uint32_t *tile_first = (uint32_t *)((char *)masks + kam_align_up(nt * PACK_THREADS * sizeof(uint2), 256));
    uint32_t *chunk_tot = (uint32_t *)((char *)tile_first + kam_align_up((nt + 1) * sizeof(uint32_t), 256));
    uint64_t *chunk_off = (uint64_t *)((char *)chunk_tot + kam_align_up((nt / SCAN_THREADS + 2) * sizeof(uint32_t), 256));
    const uint64_t n_chunks = (nt + SCAN_THREADS - 1) / SCAN_THREADS;

    KAM_CUDA(cudaMemsetAsync(d_planes, 0, kam_planes_words(total_chars) * sizeof(uint64_t), st));
    if (g_pack_fused) {
        // single pass; `state` (one word per tile) and the ticket live where the masks of the other version do
        unsigned long long *state = (unsigned long long *)masks;
        unsigned int *ticket = (unsigned int *)tile_count;
        if (nt == 0) {
            KAM_CUDA(cudaMemsetAsync(d_base_start, 0, ((size_t)n_seqs + 1) * 8, st));
        } else {
            const uint64_t n_chunks_f = (nt + PF_SUB - 1) / PF_SUB;
            KAM_CUDA(cudaMemsetAsync(state, 0, n_chunks_f * 8, st));
            KAM_CUDA(cudaMemsetAsync(ticket, 0, 4, st));
            if (n_seqs) tile_first_seq_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(d_char_start, n_seqs, nt, tile_first);
            pack_fused_kernel<<<(unsigned)n_chunks_f, PACK_THREADS, 0, st>>>(d_chars, total_chars, nt, d_char_start, n_seqs,
                                                                     (uint2 *)d_planes, d_base_start, tile_first, state,
                                                                     ticket);
        }
        if (n_seqs) head_mask_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(d_chars, d_char_start, n_seqs, d_head_mask);
        KAM_CUDA(cudaGetLastError());
        return KAM_OK;
    }
    if (nt) pack_count_kernel<<<(unsigned)nt, PACK_THREADS, 0, st>>>(d_chars, total_chars, tile_count, masks);
    if (n_chunks) chunk_sum_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(tile_count, nt, chunk_tot);
    pack_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(chunk_tot, n_chunks, chunk_off, d_char_start, n_seqs, total_chars,
                                                 d_base_start);
    if (n_chunks) tile_off_kernel<<<(unsigned)n_chunks, SCAN_THREADS, 0, st>>>(tile_count, nt, chunk_off, tile_off);
    if (nt && n_seqs)
        tile_first_seq_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(d_char_start, n_seqs, nt, tile_first);
    if (nt)
        pack_scatter_kernel<<<(unsigned)nt, PACK_THREADS, 0, st>>>(masks, total_chars, tile_off, d_char_start,
                                                                   n_seqs, (uint2 *)d_planes, d_base_start, tile_first);
    if (n_seqs) head_mask_kernel<<<(n_seqs + 255) / 256, 256, 0, st>>>(d_chars, d_char_start, n_seqs, d_head_mask);
    KAM_CUDA(cudaGetLastError());
    return KAM_OK;
}

/* test hook: 0 = three-launch pack (count / scan / scatter), default 1 = single pass with decoupled look-back */
void kam_pack_set_fused(int on) { g_pack_fused = on; }

}  // extern "C"
