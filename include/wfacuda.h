/*
 * wfacuda.h -- C ABI of libwfacuda.so, the B200 (sm_100a) wavefront-alignment
 * hot path behind the shenwei356/wfa Go API.
 *
 * The reference has no FFI seam; the seam is its exported Go API and the cut
 * goes through (*Aligner).AlignPointers (reference wfa.go:201-268): everything
 * that function calls -- initComponents (:143-184), extend (:381-458), reduce
 * (:461-540), next (:549-700), backtraceStartPosistion (:270-375), backTrace
 * (:703-983) and AlignmentResult.AddN/process (wfa_cigar.go:118-214) -- runs on
 * the GPU behind the entry points below.  INTEGRATION.md shows the cgo shim a
 * maintainer adds to the Go package; wfa_b200/host/wfa.hpp and wfa_b200/api.py
 * are the C++ and ctypes mirrors used here (no Go toolchain in this image).
 *
 * Plain C, plain pointers and sizes.  There is no CPU fallback: every entry
 * point fails loudly (negative return + wfacuda_last_error) without a GPU.
 *
 * Threading: one ctx per (host thread, device).  A ctx is not re-entrant;
 * different ctxs are independent (the reference's "one Aligner per goroutine",
 * wfa.go:73-78).  Memory: the caller owns every buffer it passes; the library
 * copies and keeps no caller pointer after a call returns (cgo pointer rule).
 */
#ifndef WFACUDA_H
#define WFACUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFACUDA_VERSION 1

/* Per-pair status (wfacuda_result.status). */
#define WFACUDA_OK               0
#define WFACUDA_ERR_EMPTY_SEQ    1   /* ErrEmptySeq,   wfa.go:187, :204-206 */
#define WFACUDA_ERR_SEQ_TOO_LONG 2   /* ErrSeqTooLong, wfa.go:193, :207-209 */
#define WFACUDA_ERR_RESOURCES    3   /* pair needs more device memory than the ctx may use */

/* Call-level return codes (0 = success). */
#define WFACUDA_E_INVALID      (-1)  /* bad argument / config */
#define WFACUDA_E_CUDA         (-2)  /* CUDA runtime error, no device, wrong arch */
#define WFACUDA_E_NOMEM        (-3)  /* host or device allocation failed */
#define WFACUDA_E_OPS_CAPACITY (-4)  /* caller's ops buffer too small; see wfacuda_batch_ops_total */

/* MaxSeqLen, wfa.go:190: offsets carry 3 type bits in a uint32. */
#define WFACUDA_MAX_SEQ_LEN ((1u << 29) - 1u)

/* Penalties (wfa.go:32-36), Options (wfa.go:64-66) and the optional
 * AdaptiveReductionOption (wfa.go:46-50) of one Aligner.  `adaptive` = 0 is the
 * reference's algn.ad == nil.  Unlike the reference's pooled Aligner (wfa.go:
 * 120-131, which keeps a stale `ad`), the setting is explicit per ctx. */
typedef struct wfacuda_config {
    uint32_t mismatch, gap_open, gap_ext;       /* DefaultPenalties = 4, 6, 2 (wfa.go:39-43) */
    uint8_t  global_alignment;                  /* Options.GlobalAlignment */
    uint8_t  adaptive;                          /* 1 after AdaptiveReduction() */
    uint8_t  reserved_[2];
    uint32_t min_wf_len, max_dist_diff;         /* 10, 50 by default (wfa.go:56-60) */
    uint32_t cutoff_step;                       /* carried, unused -- as in the reference */
    /* Tuning, 0 = automatic. */
    uint32_t flags;                             /* WFACUDA_FLAG_* */
    uint64_t arena_budget_bytes;                /* cap on the HBM backtrace arena */
} wfacuda_config;

/* Semi-global only: run to the global corner and search the start cell over
 * all retained scores exactly like wfa.go:270-375, instead of stopping at the
 * first score whose M wavefront touches the last row/column (same result). */
#define WFACUDA_FLAG_SEMIGLOBAL_LITERAL 1u
/* Force every pair through the CTA-per-pair kernel (testing). */
#define WFACUDA_FLAG_FORCE_CTA          2u
/* Force every pair through the 8-bit symbol path (testing). */
#define WFACUDA_FLAG_FORCE_8BIT         4u
/* Keep short pairs off the LANE kernel (32 pairs per warp in lockstep); testing / comparison. */
#define WFACUDA_FLAG_NO_LANE            8u
/* Keep pairs off the SLIM kernel (one warp per pair, offsets-only cells); testing / comparison. */
#define WFACUDA_FLAG_NO_SLIM            16u
/* Keep pairs off the WIDE kernel (one thread-block cluster per pair, live rows in shared memory); testing / comparison. */
#define WFACUDA_FLAG_NO_WIDE            32u

/* AlignmentResult (wfa_cigar.go:29-46) after process() (wfa_cigar.go:136-214).
 * tend/qend are 0 when the alignment has no match run (the reference leaves
 * stale pool values there, wfa_cigar.go:77-89). */
typedef struct wfacuda_result {
    uint32_t score;
    int32_t  tbegin, tend, qbegin, qend;
    uint32_t align_len, matches, gaps, gap_regions;
    uint32_t n_ops;
    uint8_t  status;
    uint8_t  reserved_[3];
} wfacuda_result;

/* Counters of the last batch (device-side work counts are the roofline
 * numerators of SURVEY.md section 8d). */
typedef struct wfacuda_stats {
    uint64_t pairs;              /* pairs aligned on the GPU */
    uint64_t cells;              /* C: sum over scores of the M wavefront width (Hi-Lo+1) before reduce */
    uint64_t cells_written;      /* wavefront cells stored in the arena (x3 components x4 bytes) */
    uint64_t score_steps;        /* existing scores processed */
    uint64_t ops;                /* R: merged ops emitted */
    uint64_t seq_bases;          /* sum n+m */
    uint64_t arena_bytes;        /* arena allocated */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t kernel_launches;    /* launches of our kernels */
    uint32_t align_launches;     /* of which wavefront kernels */
    uint32_t retries;            /* pairs re-queued with a bigger arena / wider kernel */
    uint32_t pairs_warp, pairs_cta, pairs_8bit;
    float    ms_pack, ms_align, ms_total_device;   /* CUDA-event times on the ctx stream */
    uint32_t pairs_lane;         /* pairs aligned by the LANE kernel (not counted in pairs_warp) */
    uint32_t pairs_slim;         /* pairs aligned by the SLIM kernel (not counted in pairs_warp) */
    uint32_t pairs_wide;         /* pairs aligned by the WIDE kernel (not counted in pairs_cta) */
} wfacuda_stats;

typedef struct wfacuda_ctx   wfacuda_ctx;
typedef struct wfacuda_batch wfacuda_batch;

/* Number of usable sm_100 devices (0 if none; negative on CUDA error). */
int wfacuda_device_count(void);

/* wfa.New(p, opt) (+ AdaptiveReduction) -- wfa.go:120-140.  NULL on failure;
 * wfacuda_last_error(NULL) then describes why. */
wfacuda_ctx *wfacuda_create(int device, const wfacuda_config *cfg);
/* RecycleAligner, wfa.go:102-116. */
void wfacuda_destroy(wfacuda_ctx *ctx);
/* Re-point an existing ctx at new penalties/options (wfa.New on a pooled Aligner). */
int wfacuda_set_config(wfacuda_ctx *ctx, const wfacuda_config *cfg);

/* (*Aligner).Align for many pairs (wfa.go:196-268), blocking, host buffers.
 *   seq_bytes            byte pool holding all sequences (arbitrary bytes, case-sensitive)
 *   q_off/q_len, t_off/t_len   per pair, offsets into seq_bytes
 *                        The usual pool holds the pairs back to back (q0 t0 q1 t1 ...): every
 *                        pipeline chunk then uploads one dense byte range.  Scattered pools are
 *                        fine too -- all queries then all targets, windows into one shared
 *                        reference, even per-pair heap strings addressed relative to the lowest
 *                        one (seq_bytes = that address): when the range a chunk touches is much
 *                        larger than its sequences, the library gathers them into a page-locked
 *                        pool of its own first (one host copy per sequence) and uploads that.
 *   results[n_pairs]     filled for every pair (status says which are valid)
 *   ops / ops_capacity   receives the concatenated AlignmentResult.Ops words
 *                        (op<<32|n, already reversed + merged, wfa_cigar.go:32,123,136-169);
 *                        may be NULL with capacity 0 to skip CIGARs
 *   ops_off[n_pairs]     start of pair i's n_ops words in ops (regions never overlap; their
 *                        order in the buffer is unspecified -- pairs complete in any order on the
 *                        GPU and their ops stay where they were written; the buffer may hold
 *                        unused gaps, wfacuda_last_ops_total() is its used length)
 * Returns 0, or a WFACUDA_E_* code.  On WFACUDA_E_OPS_CAPACITY results are
 * valid and wfacuda_last_ops_total() tells the capacity needed. */
int wfacuda_align_batch(wfacuda_ctx *ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                        const uint64_t *q_off, const uint32_t *q_len,
                        const uint64_t *t_off, const uint32_t *t_len,
                        wfacuda_result *results,
                        uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off);
uint64_t wfacuda_last_ops_total(const wfacuda_ctx *ctx);

/* The same call split in three, so that a caller (or bench.py) can keep a
 * batch resident in HBM: upload = H2D + planning, run = kernels only (inputs
 * and outputs stay in HBM), download = D2H.  The ops of a run live in the ctx's pool until the
 * next run on the same ctx: download a batch before running another one there. */
wfacuda_batch *wfacuda_batch_upload(wfacuda_ctx *ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                                    const uint64_t *q_off, const uint32_t *q_len,
                                    const uint64_t *t_off, const uint32_t *t_len);
int  wfacuda_batch_run(wfacuda_ctx *ctx, wfacuda_batch *b);
int  wfacuda_batch_download(wfacuda_ctx *ctx, wfacuda_batch *b, wfacuda_result *results,
                            uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off);
uint64_t wfacuda_batch_ops_total(const wfacuda_batch *b);
void wfacuda_batch_free(wfacuda_ctx *ctx, wfacuda_batch *b);

/* (*AlignmentResult).CIGAR(onlyAignedRegion) (wfa_cigar.go:236-257) and AlignmentText
 * (wfa_cigar.go:261-333, incl. trimOps :217-233) for every pair of a batch, rendered on the GPU
 * from what wfacuda_batch_run left in HBM (call it before the next run on the same ctx).
 *   cigar / cigar_capacity     receives the CIGAR bytes of all pairs, no terminators
 *   cigar_off / cigar_len[n]   pair i's string is cigar[cigar_off[i], cigar_off[i] + cigar_len[i])
 *   text / text_capacity       receives the alignment text; NULL skips it (text_off / text_len unused)
 *   text_off / text_len[n]     pair i's lines Q, A, T (query, '|' marks, target) have text_len[i] bytes
 *                              each and start at text_off[i] + j * text_len[i], j = 0, 1, 2
 * Pairs with status != 0 get empty strings, and so does the aligned region of an alignment without
 * any M (the reference's trimOps slices ops[-1:0] there and panics).  The order of the pairs'
 * strings in the buffers is unspecified, like that of the ops.  WFACUDA_E_OPS_CAPACITY: a buffer is
 * too small, wfacuda_last_render_total() then tells the sizes needed. */
int wfacuda_batch_render(wfacuda_ctx *ctx, wfacuda_batch *b, int only_aligned_region,
                         uint8_t *cigar, uint64_t cigar_capacity, uint64_t *cigar_off, uint32_t *cigar_len,
                         uint8_t *text, uint64_t text_capacity, uint64_t *text_off, uint32_t *text_len);
void wfacuda_last_render_total(const wfacuda_ctx *ctx, uint64_t *cigar_bytes, uint64_t *text_bytes);

/* One score's wavefronts of a pair's M / I / D components (WaveFront, wfa_wavefront.go:45-60):
 * diagonals lo..hi, three raw words (M, I, D: offset<<3 | backtrace code, 0 = absent,
 * wfa_backtrace_types.go:23-37) per diagonal starting at cells[first_cell]. */
typedef struct wfacuda_wavefront {
    uint32_t score;
    int32_t  lo, hi;
    uint32_t reserved_;
    uint64_t first_cell;
} wfacuda_wavefront;

/* Align ONE pair and hand out Aligner.M / I / D as Align leaves them (wfa.go:80-86; Component,
 * wfa_component.go:37-187): what the reference's Plot / Print / GetRaw read.  The pair is aligned by
 * a single worker whose arena slot is then read back; semi-global alignments run to the global
 * corner like the reference does, so every score the reference retains is present.  [lo, hi] is the
 * M wavefront's Lo / Hi after reduce; I and D cells are reported over the same range (absent = 0).
 * rows / cells may be NULL with capacity 0 to ask for the sizes (returned in n_rows / n_cells
 * together with WFACUDA_E_OPS_CAPACITY).  A debugging interface: O(wavefront cells) host memory. */
int wfacuda_align_components(wfacuda_ctx *ctx, const uint8_t *q, uint32_t q_len, const uint8_t *t, uint32_t t_len,
                             wfacuda_result *result, uint64_t *ops, uint64_t ops_capacity,
                             wfacuda_wavefront *rows, uint32_t rows_capacity, uint32_t *n_rows,
                             uint32_t *cells, uint64_t cells_capacity, uint64_t *n_cells);

/* The C side of AlignBatch over several GPUs (north-star item 4; reference contract: one
 * Aligner per goroutine, wfa.go:73-78): length-binned LPT shards (wfacuda_shard_assign), one host
 * thread and one ctx per device, every device running the chunked upload / kernels / download
 * pipeline of wfacuda_align_batch on its shard; no collective, pairs are independent.  Results land
 * at the caller's indices.  The ops buffer is cut into one region per device, in proportion to
 * the shards' bases: ops_off[i] is an absolute position in `ops`, the regions are not compacted,
 * and wfacuda_last_ops_total() returns the extent of the buffer in use (or, after
 * WFACUDA_E_OPS_CAPACITY, the capacity that makes every region large enough). */
int wfacuda_align_batch_multi(wfacuda_ctx *const *ctxs, int n_ctx, uint64_t n_pairs,
                              const uint8_t *seq_bytes,
                              const uint64_t *q_off, const uint32_t *q_len,
                              const uint64_t *t_off, const uint32_t *t_len,
                              wfacuda_result *results,
                              uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off);

/* The sharding rule of wfacuda_align_batch_multi on its own (pure host logic, no GPU): pairs are
 * binned by length (half octaves of n+m), bins are dealt out longest first, and inside a bin
 * every shard gets one run of consecutive pairs sized to level the shards' estimated loads
 * (cost ~ (n+m)^2 without heuristic or semi-global, ~ n+m with wf-adaptive).  shard_of[i] = shard
 * of pair i; shard_cost (may be NULL) = estimated load per shard. */
int wfacuda_shard_assign(int n_shards, uint64_t n_pairs, const uint32_t *q_len, const uint32_t *t_len,
                         int adaptive, int global_alignment, uint32_t *shard_of, double *shard_cost);

/* Contiguous index ranges of equal estimated cost (what wfacuda_shard_assign yields for reads of
 * one length class): cuts[0..n_shards], shard d owns pairs [cuts[d], cuts[d+1]). */
int wfacuda_shard_plan(int n_shards, uint64_t n_pairs, const uint32_t *q_len, const uint32_t *t_len,
                       int adaptive, uint64_t *cuts);

/* The chunk boundaries wfacuda_align_batch uses for a batch of n_pairs with chunks of about
 * chunk_pairs (pure host logic, no GPU): cuts[0] = 0 < ... < cuts[n_cuts - 1] = n_pairs; small
 * first chunks that double in size, `tail_levels` halving chunks at the end (< 0: uniform). */
int wfacuda_chunk_plan(uint64_t n_pairs, uint64_t chunk_pairs, int tail_levels,
                       uint64_t *cuts, uint32_t cuts_capacity, uint32_t *n_cuts);

/* Geometry of a WIDE launch (one thread-block cluster per pair, the live wavefront rows as 16-bit offsets in
 * the cluster's shared memory -- replaces the Component / WaveFront store of wfa_component.go:37-41 for the
 * scores still needed): for pairs whose widest wavefront spans `max_diagonals` = n + m - 1 diagonals and whose
 * sequences need `seq_entries` 8-byte window entries, with `smem_per_cta` bytes of shared memory per CTA:
 * the cluster size (1, 2, 4, 8), diagonals per CTA, threads per CTA and shared-memory bytes per CTA.
 * Returns WFACUDA_E_INVALID when eight CTAs cannot hold the rows (such pairs go to the CTA worker).  Pure host logic. */
int wfacuda_wide_plan(uint64_t max_diagonals, uint32_t seq_entries, uint64_t smem_per_cta,
                      int *cluster_ctas, uint32_t *diagonals_per_cta, int *threads, uint64_t *smem_bytes);

/* Page-locked host memory for the caller's input / output arrays.  Every entry point accepts
 * any host pointer; arrays that live in memory from wfacuda_host_alloc (or registered with
 * wfacuda_host_register) are moved by the DMA engines directly, without the staging copy
 * through the ctx's own pinned buffers.  The cgo shim keeps its sequence pool and result
 * arrays here (Go memory cannot be pinned, and the cgo pointer rule forbids keeping it). */
void *wfacuda_host_alloc(size_t bytes);
void  wfacuda_host_free(void *p);
int   wfacuda_host_register(void *p, size_t bytes);
int   wfacuda_host_unregister(void *p);

/* Measured INT32 issue peaks of the ctx's device in thread-level Tops/s (the roofline of SURVEY
 * section 8d names INT32 issue as a bound and asks for a measured peak): add / xor only, and add
 * alternating with mad.lo (integer ALU pipe + FMA pipe). */
int wfacuda_measure_issue_peak(wfacuda_ctx *ctx, double *alu_tops, double *mixed_tops);

int wfacuda_get_stats(const wfacuda_ctx *ctx, wfacuda_stats *out);
/* Last error text of the ctx (or of the calling thread when ctx is NULL). */
const char *wfacuda_last_error(const wfacuda_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* WFACUDA_H */
