/*
 * kameris_b200.h -- C ABI of libkameris_b200.so, the B200 (sm_100a) backend for the two kameris
 * hot paths: the `kmers` job step (FASTA records -> 4^k CGR / k-mer count vectors) and the
 * `distances` job step (pairwise distance matrices over those vectors).
 *
 * What each entry point replaces in the reference (stephensolis/kameris; binary addresses are
 * kameris/scripts/generation_{cgr,dists}_mac_avx2 virtual addresses):
 *
 *   kam_pack_ascii        the per-character classification of the CGR loop
 *                         (generation_cgr @0x10000e22e-0x10000e248: only 'A','C','G','T' are bases)
 *   kam_kmer_count        the per-file CGR fill (generation_cgr @0x10000d330 / @0x100011320, loop
 *                         @0x10000e1d2-0x10000e304) and, for KAM_OUT_FREQ_*, the numpy pass
 *                         kameris/job_steps/backend.py:42-56
 *   kam_kmers_host        `generation_cgr cgr <dir> <out> <k> <bits>` as invoked by
 *                         kameris/job_steps/backend.py:34-40, minus file I/O (host buffers in/out)
 *   kam_manhattan_*       manhat row lambda (generation_dists @0x10001bd87-0x10001bfc8)
 *   kam_info_tri          info row lambda   (generation_dists @0x10001bab5-0x10001bcb0)
 *   kam_gram_*            euclid / cosine / pearson -- NOT in the reference (schema enum
 *                         kameris/schemas/job_options.json:97 has only manhat, info); defined as
 *                         scipy.spatial.distance {euclidean, cosine, correlation} on the rows as stored:
 *                         count vectors, or the float64 cgr / cgr.sum() rows of a `frequencies` file
 *                         (kameris/job_steps/backend.py:45-56)
 *   kam_dists_host        `generation_dists <in> <prefix> <names>` as invoked by
 *                         kameris/job_steps/backend.py:59-66, minus file I/O
 *   kam_fasta_record0 /   the front half of the per-file task: ifstream + getline + trim + first-record
 *   kam_fasta_read_files  selection (generation_cgr @0x10000d691-0x10000d83e), for one file image / for a whole
 *                         listing on the host threads of mmg::ParallelExecutor (@0x10001a6f0)
 *   kam_cgr_pool,         device forms of what the reference gets by re-running generation_cgr at another k
 *   kam_normalize          and of the numpy pass backend.py:45-49 on existing count tables
 *   kam_col_stats /       sklearn StandardScaler(with_mean=False).fit / .transform as used by
 *   kam_col_scale         kameris/job_steps/classify.py:29-50 (build_pipeline)
 *   kam_mds_center        the numpy double-centring of kameris/job_steps/mds.py:14-24
 *   kam_symm_block        the matrix-block product of the eigen-solve that replaces
 *                         scipy.sparse.linalg.eigsh(B, k=dim), kameris/job_steps/mds.py:26
 *
 * Conventions: every function returns 0 on success and a negative KAM_E_* code on failure;
 * kam_last_error() returns a thread-local message.  No exceptions and no ownership cross the ABI:
 * the caller owns every buffer.  Pointers named d_* are device pointers on the current CUDA
 * device, h_* are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Device entry points are asynchronous on `stream` unless stated otherwise.
 */
#ifndef KAMERIS_B200_H
#define KAMERIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KAM_VERSION 100 /* 0.1.0 */

/* error codes */
#define KAM_OK 0
#define KAM_E_INVALID (-1)   /* bad argument */
#define KAM_E_CUDA (-2)      /* CUDA runtime error, see kam_last_error() */
#define KAM_E_WORKSPACE (-3) /* workspace too small */
#define KAM_E_UNSUPPORTED (-4)
#define KAM_E_NOMEM (-5)

/* element types: the .mm-repr / .mm-dist element_type codes */
#define KAM_U8 0
#define KAM_U16 1
#define KAM_U32 2
#define KAM_U64 3
#define KAM_F32 4
#define KAM_F64 5

/* kam_kmer_count output kinds */
#define KAM_OUT_COUNTS 0   /* uint16 / uint32 counts, wrapping mod 2^bits like the reference */
#define KAM_OUT_FREQ_F32 1 /* (float)((double)count / sum) */
#define KAM_OUT_FREQ_F64 2 /* (double)count / sum: bit-exact numpy true division (backend.py:49) */

/* distance metrics (bit mask) */
#define KAM_M_MANHAT 1u  /* uint32, reference `manhat` */
#define KAM_M_INFO 2u    /* float32, reference `info` */
#define KAM_M_EUCLID 4u  /* float32 */
#define KAM_M_COSINE 8u  /* float32 */
#define KAM_M_PEARSON 16u /* float32 */

int kam_version(void);
const char *kam_last_error(void);
/* SM count, free and total bytes of the current device. */
int kam_device_info(int *sm_count, size_t *free_bytes, size_t *total_bytes);

/* ---------------------------------------------------------------------------------------------
 * Packed sequence layout ("planes").  All sequences of a batch form ONE flat stream of valid
 * bases (non-ACGT characters removed, sequences back to back).  Base p of the stream lives in
 * 64-bit word p/32: bit p%32 of the low half is its x-bit (1 for G,T), bit p%32 of the high half
 * its y-bit (1 for A,T) -- the two CGR coordinate bits of generation_cgr @0x10000e24a-0x10000e2a4.
 * d_base_start[s] .. d_base_start[s+1] is the range of sequence s; d_head_mask[s] has bit i set
 * when raw character i (i < 32) of sequence s is a base, which is all that is needed to reproduce
 * the raw-index emit test of the reference (quirk Q1).
 * ------------------------------------------------------------------------------------------- */

/* number of 64-bit plane words to allocate for `total_chars` input characters */
size_t kam_planes_words(uint64_t total_chars);
size_t kam_pack_workspace_bytes(uint64_t total_chars, uint32_t n_seqs);

/* ASCII -> planes.  d_chars (16-byte aligned) holds the records back to back; d_char_start[n+1]
 * their offsets.  Writes d_planes[kam_planes_words()], d_base_start[n+1], d_head_mask[n]. */
int kam_pack_ascii(const uint8_t *d_chars, const uint64_t *d_char_start, uint32_t n_seqs,
                   uint64_t total_chars, uint64_t *d_planes, uint64_t *d_base_start,
                   uint32_t *d_head_mask, void *d_ws, size_t ws_bytes, void *stream);

/* test hook: 0 = three-launch packer (count / scan / scatter); default 1 = single pass (classification,
 * chained scan with decoupled look-back, compaction in one kernel); identical outputs. */
void kam_pack_set_fused(int on);

size_t kam_kmer_workspace_bytes(uint32_t n_seqs, int k);

/* planes -> n_seqs vectors of 4^k bins (row s at d_out + s*out_stride elements).
 * total_bases_hint: total number of bases (an upper bound such as the character count is fine; it
 * only selects the CTA size for short reads); 0 = unknown, the library reads it from the device.
 * count_bits: 16|32 (wrap width of the counts, also for the frequencies derived from them).
 * out_kind: KAM_OUT_*.  Element type of d_out: uint16/uint32 for COUNTS, float/double for FREQ.
 * Synchronises `stream` once internally (to learn how many sequences need the long-sequence
 * path); the results are complete on return of the stream's work. */
int kam_kmer_count(const uint64_t *d_planes, const uint64_t *d_base_start,
                   const uint32_t *d_head_mask, uint32_t n_seqs, uint64_t total_bases_hint, int k,
                   int count_bits, int out_kind, void *d_out, uint64_t out_stride, void *d_ws,
                   size_t ws_bytes, void *stream);

/* The same fill with a SPARSE result, for vectors that are mostly zeros (a 9 kb genome touches <= 8992 of
 * the 262144 bins of k = 9): vector s becomes d_nnz[s] packed entries (bin << 8 | count), in ascending bin
 * order, at d_entries + d_entry_start[s].  An entry list never has more entries than the sequence has
 * characters, so d_entry_start may simply be the character offsets of the batch (d_char_start of
 * kam_pack_ascii) and d_entries hold total_chars words.  7 <= k <= 12.  A sequence this path cannot take
 * (longer than 131072 bases, or a bin count above 255) gets d_nnz[s] = KAM_SPARSE_DECLINED; the caller
 * recomputes it with kam_kmer_count.  The dense vector (generation_cgr @0x10000d330) is the expansion of the
 * list over zeros; wrap width and output type are applied by whoever expands.  Fully asynchronous on
 * `stream` (no internal synchronisation).  Workspace: kam_kmer_workspace_bytes(n_seqs, k). */
#define KAM_SPARSE_MIN_K 7
#define KAM_SPARSE_MAX_K 12
#define KAM_SPARSE_DECLINED 0xffffffffu
int kam_kmer_count_sparse(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                          uint32_t n_seqs, int k, const uint64_t *d_entry_start, uint32_t *d_entries,
                          uint32_t *d_nnz, void *d_ws, size_t ws_bytes, void *stream);

/* Every order k_lo..k_hi in one call (BASELINE config 2: k = 1..9).  h_outs is a HOST array of
 * k_hi-k_lo+1 device pointers (order k at index k-k_lo), each a dense n_seqs x 4^k matrix.  Orders
 * >= 8 run the ordinary tiled kernel (they are output-bound); for genome-length sequences (<= 65535
 * bases) the orders <= 7 come from ONE fused kernel: a single scan at order 7, whose table fits one
 * shared-memory tile, and every lower order by 2x2 pooling of that tile inside the CTA (SURVEY.md
 * A.3), so orders 1..6 cost only their output bytes.  Workspace: kam_kmer_workspace_bytes(n_seqs, k_hi). */
int kam_kmer_sweep(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                   uint32_t n_seqs, uint64_t total_bases_hint, int k_lo, int k_hi, int count_bits,
                   int out_kind, void *const *h_outs, void *d_ws, size_t ws_bytes, void *stream);

/* tuning / test hook: top order of the fused part of kam_kmer_sweep (6..10, default 7). */
void kam_kmer_sweep_set_fused_top(int k);
/* test hook: 0 sends short reads at k <= 6 through the CTA-per-read kernel instead of the warp-per-read
 * kernel (default 1); both produce identical vectors. */
void kam_kmer_set_warp_path(int on);
/* test hook: 0 sends long sequences (and uint8-counter overflows) straight to the global-memory RED.ADD
 * path instead of the shared-memory uint16 tile kernel (default 1); identical vectors either way. */
void kam_kmer_set_long_path(int on);
/* test hook: 0 keeps long sequences at k <= 9 on the tiled shared-memory kernel (one CTA per (sequence, tile),
 * one scan of the sequence per tile) instead of the one-pass kernel that holds the whole 4^k table of a sequence
 * in the shared memory of one CTA (default 1; at k = 9 only for batches of at least half as many long sequences
 * as the device has SMs -- 2 lifts that condition); identical vectors either way. */
void kam_kmer_set_long_one_pass(int on);
/* test hook: 0 makes the CTA-per-sequence kernel scan 8 k-mer end positions per thread and trip (the first
 * version) instead of one whole plane word of 32 (default 1); identical vectors either way. */
void kam_kmer_set_word_scan(int on);

/* Order-(k+1) count tables -> order-k tables by 2x2 sum-pooling of the CGR image plus the one k-mer
 * that has no (k+1)-mer ending at the same base; exact for dirty heads (quirk Q1) and
 * for wrapped uint16 counts.  d_in rows hold 4^(k+1) bins, d_out rows 4^k bins, both count_bits wide. */
int kam_cgr_pool(const uint64_t *d_planes, const uint64_t *d_base_start, const uint32_t *d_head_mask,
                 uint32_t n_seqs, int k, int count_bits, const void *d_in, uint64_t in_stride,
                 void *d_out, uint64_t out_stride, void *stream);
/* counts (uint16/uint32, n rows of d bins) -> frequencies cgr/cgr.sum() (backend.py:49);
 * out_kind KAM_OUT_FREQ_F32 | KAM_OUT_FREQ_F64.  kam_kmer_count fuses this for fresh tables. */
int kam_normalize(const void *d_counts, int count_bits, uint32_t n, uint64_t d, uint64_t in_stride,
                  int out_kind, void *d_out, uint64_t out_stride, void *stream);

/* Host-buffer form of the whole `kmers` step (what generation_cgr + the numpy pass compute):
 * h_chars/h_char_start are the concatenated FASTA records (record[0] of each file, see
 * kam_fasta_record0), h_out receives n_seqs x 4^k elements.  Copies, packing, counting and the
 * copy back are pipelined over chunks on internal streams; blocking. */
int kam_kmers_host(const uint8_t *h_chars, const uint64_t *h_char_start, uint32_t n_seqs, int k,
                   int count_bits, int out_kind, void *h_out);
/* The same step for a job whose next step is `distances` (kameris/subcommands/run_job.py:175-180 runs the steps
 * of an experiment back to back): besides (or, with h_out == NULL, instead of) returning the vectors to the host,
 * the dense COUNT rows (uint16/uint32 per count_bits, whatever out_kind is) stay in the caller's device matrix
 * d_counts (row s at d_counts + s * counts_stride elements), complete when the call returns.  kam_dists_device
 * then works on them directly: the step neither re-reads the .mm-repr nor uploads it again. */
int kam_kmers_host_keep(const uint8_t *h_chars, const uint64_t *h_char_start, uint32_t n_seqs, int k,
                        int count_bits, int out_kind, void *h_out, void *d_counts, uint64_t counts_stride);
/* How kam_kmers_host brings the vectors back: 0 = always dense (kernel writes the rows, the copy engine moves
 * them), 1 = default: chunks whose vectors are mostly zeros (k >= 7 and entry lists <= 1/4 of the dense bytes)
 * return as kam_kmer_count_sparse lists and are expanded into h_out by the host threads, 2 = sparse whenever
 * 7 <= k <= 12 (test hook).  The rows written are identical either way.  Calls of kam_kmers_host are
 * serialised by the library (one pipeline per device); it may be called from any thread. */
void kam_kmers_host_set_sparse(int mode);
/* Host worker threads of the library (list expansion): default = the cores of the process' affinity mask,
 * divided by LOCAL_WORLD_SIZE under torchrun; KAM_HOST_THREADS in the environment or this call override. */
void kam_host_set_threads(int threads);
int kam_host_threads(void);
/* Host-side expansion of kam_kmer_count_sparse lists (copied back by the caller) into n_seqs dense rows of
 * 4^k elements: uint16/uint32 counts (count_bits) or float/double frequencies count/sum per out_kind, the sum
 * being that of the list -- i.e. exactly what kam_kmer_count would have written.  Every cache line of a row
 * is written once with non-temporal stores, on the library's host threads.  No GPU involved. */
int kam_sparse_expand(const uint32_t *h_entries, const uint64_t *h_entry_start, const uint32_t *h_nnz,
                      uint32_t n_seqs, int k, int count_bits, int out_kind, void *h_out);

/* A.1: first FASTA record of a file image (generation_cgr @0x10000d691-0x10000d83e).  Writes at
 * most n bytes to out; returns the record length or -1 when the file holds no record. */
int64_t kam_fasta_record0(const char *buf, size_t n, char *out);
/* The same for a whole directory listing, on `threads` host threads (0 = all; the task scheme of
 * mmg::ParallelExecutor @0x10001a6f0): kam_fasta_file_sizes stats every path (a non-regular entry has
 * size 0); with file_offsets = the exclusive prefix sum of those sizes (n+1 entries) kam_fasta_read_files
 * reads the files and leaves record[0] of file i at h_chars[start[i] .. start[i+1]) -- contiguous, in path
 * order, ready for kam_kmers_host.  h_chars needs file_offsets[n] bytes.  A path that cannot be read, is not
 * a regular file or holds no record fails the call (KAM_E_INVALID, *bad_index = the first such path). */
int kam_fasta_file_sizes(const char *const *paths, uint32_t n, uint64_t *h_sizes, int threads);
int kam_fasta_read_files(const char *const *paths, uint32_t n, const uint64_t *file_offsets, uint8_t *h_chars,
                         uint64_t *h_start, int threads, uint32_t *bad_index);

/* ---------------------------------------------------------------------------------------------
 * Distances.  x is a row-major n x d matrix (row stride `ld` elements) of element type `elem`.
 * Triangle outputs use the .mm-dist order: for i<j, index n*i - i*(i+3)/2 + j - 1
 * (generation_dists @0x10001bf7e).  [row_begin,row_end) selects the block of rows i a rank
 * computes (all j>i); entries of other rows are left untouched.
 * ------------------------------------------------------------------------------------------- */
/* manhat and info on the tensor cores (the reference's two metrics: generation_dists row lambda, manhat body
 * @0x10001bd87, info body @0x10001bab5).  Neither is a contraction as written, but
 *     sum_e |a_e - b_e| = sum(a) + sum(b) - 2 sum_e min(a_e, b_e),   min(a_e, b_e) = sum_{t >= 1} [a_e >= t][b_e >= t]
 * so the pairwise min-sums are an exact integer Gram (tcgen05 kind::i8, int32 accumulators) of "thermometer" rows
 * of indicator bytes -- column e gets as many bytes as the largest count any row holds there, so a row is
 * K = sum of the column maxima bytes long (2-3 x 4^k for k-mer counts) -- and info's |supp_a & supp_b| is the Gram
 * of the presence bytes.  The same uint32 sums (mod 2^32) and the same float32 formula come out, bit for bit.
 * Taken when every count is <= 255, K <= 12 d, the shape fills 256 x 256 tiles (>= 512 rows, >= 128 columns) and
 * the workspace holds the operand rows (kam_manhattan_workspace_bytes covers K <= 4 d,
 * kam_manhattan_workspace_bytes_levels(rows, d, a) K <= a d); otherwise the CUDA-core tile kernel below runs
 * (VABSDIFF4 / VABSDIFF / POPC).  kam_dist_set_tensor: 0 = never, 1 = as described (default), 2 = for every shape
 * and any K that fits (tests). */
size_t kam_manhattan_workspace_bytes_levels(uint32_t rows, uint64_t d, uint32_t avg_levels);
void kam_dist_set_tensor(int mode);
/* The manhattan entry points synchronise `stream` once (they look at the largest element to pick
 * the tensor path or the 4-elements-per-instruction byte path when every count is <= 255).
 * elem: KAM_U8/U16/U32 (tile kernel); kam_manhattan_tri also takes KAM_U64/F32/F64 through a
 * serial kernel that reproduces the reference's truncating float accumulate (quirk Q5). */
size_t kam_manhattan_workspace_bytes(uint32_t rows, uint64_t d);
int kam_manhattan_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld,
                      uint32_t row_begin, uint32_t row_end, uint32_t *d_tri, void *d_ws,
                      size_t ws_bytes, void *stream);
/* rectangular m x c: out[i*c + j] = manhat(x_i, y_j); workspace = ws(m,d) + ws(c,d) */
int kam_manhattan_rect(const void *d_x, const void *d_y, int elem, uint32_t m, uint32_t c,
                       uint64_t d, uint64_t ldx, uint64_t ldy, uint32_t *d_out, void *d_ws,
                       size_t ws_bytes, void *stream);
/* float32 L1 distance to float32 centroids (BASELINE config 4); out[i*c+j] = sum |x_i - y_j| */
int kam_manhattan_rect_f32(const void *d_x, int elem_x, const float *d_y, uint32_t m, uint32_t c,
                           uint64_t d, uint64_t ldx, uint64_t ldy, float *d_out, void *d_ws,
                           size_t ws_bytes, void *stream);
size_t kam_info_workspace_bytes(uint32_t n, uint64_t d);
int kam_info_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld,
                 uint32_t row_begin, uint32_t row_end, float *d_tri, void *d_ws, size_t ws_bytes,
                 void *stream);

/* Gram-matrix metrics on count vectors (uint8/uint16/uint32): G = X X^T on the tcgen05 tensor cores
 * with exact operands: uint8 x uint8 -> int32 (kind::i8) when every count <= 255, integer-valued
 * fp16 -> fp32 (kind::f16) up to 2048, exact uint64 CUDA-core Gram beyond; fused float64 epilogue.
 * metrics: any of KAM_M_EUCLID|KAM_M_COSINE|KAM_M_PEARSON; d_tri[m] (may be NULL when the metric
 * is not requested) receive float32 triangles.  euclid_on_freq != 0 computes euclidean distance
 * between the frequency vectors x_i/sum(x_i) (what a `frequencies` .mm-repr holds).
 *
 * elem KAM_F32 / KAM_F64 -- the rows of a `mode: frequencies` .mm-repr (backend.py:45-56: cgr / cgr.sum()):
 * the pre-pass recovers, row by row, the integer counts behind the quotients and accepts them only when
 * they reproduce every stored element bit for bit and sum to the divisor; the exact integer Gram of the
 * recovered counts then runs on the tensor cores as above, with euclid between the rows as stored (the
 * frequency vectors; euclid_on_freq is ignored).  If any row is not a frequency vector (or the recovered
 * counts leave the tensor envelopes) the whole matrix takes a float64 Gram on the CUDA cores instead. */
size_t kam_gram_workspace_bytes(uint32_t n, uint64_t d);
/* workspace for inputs outside the tensor-exact envelope (integer CUDA-core Gram; 4 B per element).
 * kam_gram_tri returns KAM_E_WORKSPACE when it needs this size and was given less. */
size_t kam_gram_workspace_bytes_int(uint32_t n, uint64_t d);
/* test hook: non-zero forces the fp16 (kind::f16) tensor path where the uint8 (kind::i8) path would
 * be chosen; both are exact inside their envelopes. */
void kam_gram_force_f16(int on);
/* test hook: 0 sends floating-point rows straight to the float64 CUDA-core Gram (no integer recovery);
 * default 1. */
void kam_gram_set_float_recover(int on);
/* tuning/test hook: thread-block cluster size of the Gram kernel (2 = B panel multicast between the
 * two CTAs of a cluster, default; 1 = no cluster). */
void kam_gram_set_cluster(int cl);
/* tuning/test hook: non-zero (default) runs the CTA-pair kernel (tcgen05.mma.cta_group::2, one 256 x 256
 * tile per cluster of two CTAs, B split between them); 0 the cluster-multicast cta_group::1 kernel.  Needs
 * cluster 2.  Both produce the exact integer Gram matrix, hence identical outputs. */
void kam_gram_set_pair(int on);
/* test hook: 0 sends count vectors outside the uint8 / fp16 envelopes straight to the integer CUDA-core Gram;
 * default 1 tries the limb-split tensor path first (counts <= 65535: c = 256 hi + lo, three kind::i8 Grams
 * combined exactly in float64; taken when the byte limbs keep every int32 accumulator below 2^31). */
void kam_gram_set_limb(int on);
int kam_gram_tri(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld,
                 uint32_t row_begin, uint32_t row_end, unsigned metrics, int euclid_on_freq,
                 float *d_tri_euclid, float *d_tri_cosine, float *d_tri_pearson, void *d_ws,
                 size_t ws_bytes, void *stream);

/* Block interface for multi-GPU runs (kameris_b200/parallel.py): every rank converts ITS rows once
 * (kam_gram_prepare), the ranks exchange operand rows and row statistics, and each rank computes the
 * blocks it owns.  operand_kind is global: KAM_GRAM_U8 needs every count <= 255 and every row's sum
 * of squares < 2^31, KAM_GRAM_F16 needs <= 2048 and < 2^24 (d_flags[0..1] receive, by atomicMax, the
 * largest count and the largest sum of squares seen; the caller zeroes them and reduces over ranks).
 * Operand rows are kam_gram_dpad(d) elements long (uint8 or fp16), zero padded.  Floating-point rows
 * (KAM_F32 / KAM_F64 frequency vectors) are converted through the integer recovery described above;
 * d_flags[2] receives the number of rows that could not be recovered (the block interface has no
 * float64 path: the caller must treat a non-zero count as an error).  d_flags holds >= 3 words. */
#define KAM_GRAM_U8 0
#define KAM_GRAM_F16 1
uint64_t kam_gram_dpad(uint64_t d);
int kam_gram_prepare(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, int operand_kind,
                     void *d_operand, uint64_t *d_S, uint64_t *d_G, uint64_t *d_flags, void *stream);
size_t kam_gram_block_workspace_bytes(uint32_t n_a, uint32_t n_b);
/* rows [row_begin,row_end) of A against all rows of B.  triangular != 0: A and B are the same
 * matrix and only pairs j > i are written.  out_ld == 0 (triangular only): packed triangle of n_a;
 * otherwise dense row-major output, out[i*out_ld + j].  sm_limit: number of SMs the kernel may occupy
 * (0 = all; a multi-GPU caller leaves a few to the concurrent NCCL transfers).  Asynchronous on `stream`:
 * the tile schedule of a geometry is built once and cached on the device by the library. */
int kam_gram_block(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_G_a, uint32_t n_a,
                   const void *d_op_b, const uint64_t *d_S_b, const uint64_t *d_G_b, uint32_t n_b,
                   int operand_kind, uint64_t d, int triangular, uint32_t row_begin, uint32_t row_end,
                   unsigned metrics, int euclid_on_freq, float *d_out_euclid, float *d_out_cosine,
                   float *d_out_pearson, uint64_t out_ld, int sm_limit, void *d_ws, size_t ws_bytes, void *stream);
/* The exchange step of the multi-GPU `distances` step fused into the Gram kernel (the step
 * being sharded: kameris/job_steps/backend.py:59-66).  ONE persistent launch computes every block of a rank:
 * rows of A against the rows of B, where B is the concatenation [rows of A itself | rows of other ranks' shards]
 * and the latter are still being pulled over NVLink by the copy engines when the kernel starts (kam_peer_pull).
 *   triangular_head != 0: B starts with the rows of A (d_op_b == d_op_a); pairs j <= i are skipped there.
 *   h_segments: n_segments x (column tile begin, column tile end, row tile begin, row tile end), tiles of 256
 *     rows; the tiles of a segment are scheduled after those of the segments before it, so list the segments in
 *     the order their rows arrive.
 *   d_ready: one word per chunk of `ready_tiles` column tiles, zero until the rows of those tiles are in place
 *     (NULL: everything is there).  The TMA producer of a tile waits for its word (acquire, system scope) before it
 *     loads the tile's B rows, so transfers overlap the tensor-core work tile by tile and the launch never idles
 *     at a block boundary.  d_S_b / d_G_b (all n_b rows) must be complete at launch.
 * Output: dense rows out[i * out_ld + j], j indexing the rows of B.  Workspace: kam_gram_block_workspace_bytes. */
int kam_gram_fused(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_G_a, uint32_t n_a,
                   const void *d_op_b, const uint64_t *d_S_b, const uint64_t *d_G_b, uint32_t n_b,
                   int operand_kind, uint64_t d, int triangular_head, const uint32_t *h_segments,
                   uint32_t n_segments, const uint32_t *d_ready, uint32_t ready_tiles, unsigned metrics,
                   int euclid_on_freq, float *d_out_euclid, float *d_out_cosine, float *d_out_pearson,
                   uint64_t out_ld, int sm_limit, void *d_ws, size_t ws_bytes, void *stream);
/* manhat / info across ranks (parallel.thermo_ring): the pieces of the tensor path above as
 * entry points, so that every rank encodes ITS rows with a layout all ranks share and the operand rows travel and
 * are consumed exactly like those of the Gram metrics (kam_peer_pull + one fused launch).
 *   kam_dist_colmax   d_colmax[e] = max(d_colmax[e], max_r x[r][e])   (caller zeroes it; all-reduce(MAX) it over ranks)
 *   kam_dist_layout   d_off[0..d] = running sum of d_colmax; h_stats = {largest count, K = their sum}; synchronises
 *   kam_dist_encode   rows of kpad = kam_dist_kpad(K) bytes (d_off != NULL: manhat) or of kam_gram_dpad(d) presence bytes
 *                     (d_off == NULL: info) into d_operand, sum(c) and nnz(c) per row into d_S / d_NZ;
 *                     workspace: 48 bytes per row
 *   kam_thermo_fused  kam_gram_fused's arguments (B = [rows of A | rows of other ranks], segments, arrival flags) with
 *                     the statistics d_S / d_NZ; metric = KAM_M_MANHAT (uint32 out) or KAM_M_INFO (float32 out) */
int kam_dist_colmax(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, uint32_t *d_colmax, void *stream);
int kam_dist_layout(const uint32_t *d_colmax, uint64_t d, uint32_t *d_off, uint64_t *d_stats, uint64_t *h_stats,
                    void *stream);
uint64_t kam_dist_kpad(uint64_t k_total);
int kam_dist_encode(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const uint32_t *d_off,
                    uint64_t kpad, void *d_operand, uint64_t *d_S, uint64_t *d_NZ, void *d_ws, size_t ws_bytes,
                    void *stream);
int kam_thermo_fused(const void *d_op_a, const uint64_t *d_S_a, const uint64_t *d_NZ_a, uint32_t n_a,
                     const void *d_op_b, const uint64_t *d_S_b, const uint64_t *d_NZ_b, uint32_t n_b, uint64_t d,
                     uint64_t kpad, int triangular_head, const uint32_t *h_segments, uint32_t n_segments,
                     const uint32_t *d_ready, uint32_t ready_tiles, unsigned metric, void *d_out, uint64_t out_ld,
                     void *d_ws, size_t ws_bytes, void *stream);
/* diagnostics hook of kam_gram_fused: three device words that receive the cycles its TMA producers spent waiting for
 * rows (sum, longest single wait, number of waits); NULL = off (default). */
void kam_gram_set_wait_probe(uint64_t *d_three_words);
/* tuning hook: row tiles (of 128 rows) per strip of the tile schedule (default 12); larger strips re-use the
 * column panels longer in L2 at the price of more row panels in flight. */
void kam_gram_set_strip(int row_tiles);

/* ---------------------------------------------------------------------------------------------
 * Peer memory (one process per GPU on one NVLink / NVSwitch box).  kam_peer_alloc gives device memory that
 * kam_peer_export turns into a KAM_PEER_HANDLE_BYTES handle; another process passes the handle (sent through
 * any channel, e.g. a torch.distributed all-gather) to kam_peer_open and gets a pointer it can READ with
 * kam_peer_copy -- a copy-engine transfer over NVLink that occupies no SM, unlike NCCL send/recv kernels, so
 * it does not disturb a persistent kernel that owns every SM (the Gram kernel whose next block's operand rows
 * it fetches).  Opened handles are cached by the library; kam_peer_close unmaps one.  The owner must keep the
 * buffer alive and unmodified until its readers are done (the callers use a barrier).
 * ------------------------------------------------------------------------------------------- */
#define KAM_PEER_HANDLE_BYTES 64
int kam_peer_alloc(size_t bytes, void **d_ptr);
int kam_peer_free(void *d_ptr);
int kam_peer_export(const void *d_ptr, unsigned char *handle);
int kam_peer_open(const unsigned char *handle, void **d_mapped);
int kam_peer_close(void *d_mapped);
int kam_peer_copy(void *d_dst, const void *d_src, size_t bytes, void *stream);
/* A batch of copy-engine transfers with arrival flags: the pieces are enqueued in order, all pieces of chunk c on
 * streams[c % n_streams] (two streams = two copy engines), and behind the last piece of chunk c the stream writes
 * d_flags[c] = 1 (cuStreamWriteValue32: no kernel) -- what a running kam_gram_fused launch waits for.  Pieces
 * with chunk == KAM_PEER_NO_FLAG go to streams[0] and set nothing.  kam_peer_signal is the flag write alone. */
#define KAM_PEER_NO_FLAG 0xffffffffu
typedef struct {
    void *d_dst;
    const void *d_src;
    uint64_t bytes;
    uint32_t chunk;
    uint32_t reserved;
} kam_peer_piece;
int kam_peer_pull(const kam_peer_piece *pieces, uint32_t n_pieces, uint32_t *d_flags, void *const *streams,
                  uint32_t n_streams);
int kam_peer_signal(uint32_t *d_flag, void *stream);

/* Host-buffer form of the whole `distances` step: x is the n x d matrix read from a .mm-repr,
 * outputs are packed triangles (NULL = not requested).  euclid is taken between the rows as
 * stored (count vectors of a `counts` file, frequency vectors of a `frequencies` file).  Blocking; may be
 * called from any thread (calls are serialised per process); device buffers are kept between calls
 * (kam_host_release frees them). */
int kam_dists_host(const void *h_x, int elem, uint32_t n, uint64_t d, unsigned metrics,
                   uint32_t *h_manhat, float *h_info, float *h_euclid, float *h_cosine,
                   float *h_pearson);
/* The same with the matrix already on the device (row stride ld elements; complete before the call: the work
 * runs on an internal stream), e.g. the counts kam_kmers_host_keep left there.  euclid_on_freq != 0: euclid
 * between the frequency vectors x / sum(x) -- what the step computes on the `frequencies` file made from these
 * counts (cosine and pearson do not depend on the scaling, info only on which bins are non-zero; manhat on a
 * frequencies file truncates per addition (quirk Q5) and must be computed from the float rows themselves). */
int kam_dists_device(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, unsigned metrics,
                     int euclid_on_freq, uint32_t *h_manhat, float *h_info, float *h_euclid, float *h_cosine,
                     float *h_pearson);
/* frees the device buffers the host-buffer entry points keep between calls */
void kam_host_release(void);

/* ---------------------------------------------------------------------------------------------
 * mds job step (kameris/job_steps/mds.py:10-35): classical multidimensional scaling.
 * kam_mds_center replaces the numpy double-centring of mds.py:14-24 (same operation order, float64):
 * packed `.mm-dist` triangle of n points (elem KAM_U32 or KAM_F32) -> d_B, n x n row-major.
 * kam_symm_block is the n^2 product of the block subspace iteration that replaces
 * scipy.sparse.linalg.eigsh(B, k=dim) of mds.py:26: d_Y = d_B * d_Q with d_Q, d_Y n x 16 row-major
 * (a narrower block is zero-padded by the caller).  Both asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------- */
#define KAM_MDS_BLOCK 16
size_t kam_mds_workspace_bytes(uint32_t n);
int kam_mds_center(const void *d_tri, int elem, uint32_t n, double *d_B, void *d_ws, size_t ws_bytes, void *stream);
size_t kam_symm_workspace_bytes(uint32_t n);
int kam_symm_block(const double *d_B, uint32_t n, const double *d_Q, double *d_Y, void *d_ws, size_t ws_bytes,
                   void *stream);
/* test hook: 0 = always the register-pipelined product kernel; default 1 = TMA-fed shared-memory tiles
 * whenever n is even (16-byte-aligned rows); identical results up to the summation order. */
void kam_symm_set_tma(int on);

/* ---------------------------------------------------------------------------------------------
 * classify pre-processing (kameris/job_steps/classify.py:29-50): the column statistics and
 * the transform of sklearn's StandardScaler(with_mean=False) on a row-major n x d feature matrix (elem
 * KAM_U8/U16/U32 count vectors or KAM_F32/F64 frequency vectors), float64 results:
 *   mean_c = sum x_ic / n_c,  var_c = (sum (x_ic - mean_c)^2 - (sum (x_ic - mean_c))^2 / n_c) / n_c
 * over the n_c non-NaN entries of column c (sklearn/utils/extmath.py _incremental_mean_and_var), and
 * out_ic = x_ic / scale_c.  Asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------- */
size_t kam_colstats_workspace_bytes(uint32_t n, uint64_t d);
int kam_col_stats(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, double *d_mean, double *d_var,
                  int64_t *d_count, void *d_ws, size_t ws_bytes, void *stream);
int kam_col_scale(const void *d_x, int elem, uint32_t n, uint64_t d, uint64_t ld, const double *d_scale,
                  double *d_out, uint64_t out_ld, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KAMERIS_B200_H */
