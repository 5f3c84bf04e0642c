/*
 * wfa_oracle.h -- CPU restatement of shenwei356/wfa's wavefront hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or
 * the CPU baseline.  libwfacuda.so never links, loads or calls it.
 *
 * Parity status: PINNED against the reference's README golden alignments
 * (README.md:18-27, 101-124, 128-149, 231-240, 245-254; see
 * tests/golden/readme_vectors.json) and the two README M-component tables.
 * `reduce` (wf-adaptive) and semi-global tie cases have no golden vector in
 * the reference; they are pinned only by the second, independent restatement
 * in oracle/pyoracle.py (tests/test_oracle_vs_pyoracle.py).
 *
 * The Go toolchain is absent in this image, so the reference itself cannot
 * be compiled into oracle/_ref; this C restatement is the oracle and the
 * "port" CPU baseline.
 */
#ifndef WFA_ORACLE_H
#define WFA_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* wfa.go:32-60 Penalties / AdaptiveReductionOption, wfa.go:64-66 Options */
typedef struct {
    uint32_t mismatch, gap_open, gap_ext;
    uint8_t  global_alignment;   /* Options.GlobalAlignment */
    uint8_t  adaptive;           /* algn.ad != nil */
    uint8_t  pad_[2];
    uint32_t min_wf_len, max_dist_diff;
} oracle_config;

/* AlignmentResult (wfa_cigar.go:29-46) after process().  Same field order as
 * wfacuda_result so the tests can compare records byte-wise. */
typedef struct {
    uint32_t score;
    int32_t  tbegin, tend, qbegin, qend;
    uint32_t align_len, matches, gaps, gap_regions;
    uint32_t n_ops;
    uint8_t  status;             /* 0 ok, 1 ErrEmptySeq, 2 ErrSeqTooLong */
    uint8_t  pad_[3];
} oracle_result;

/* SURVEY.md section 8(d) work counters (roofline numerators). */
typedef struct {
    uint64_t cells;      /* C: sum over existing scores of M width (Hi-Lo+1) before reduce */
    uint64_t visits;     /* V: present diagonals visited by extend */
    uint64_t words;      /* W: sum over extended diagonals of ceil((LCP+1)/16) */
    uint64_t ops;        /* R: merged ops */
    uint64_t scores;     /* number of existing M wavefronts */
    uint64_t max_width;  /* widest M wavefront */
} oracle_counters;

typedef struct oracle_aligner oracle_aligner;

#define ORACLE_OK        0
#define ORACLE_EMPTY     1
#define ORACLE_TOO_LONG  2
#define ORACLE_MAX_SEQ_LEN ((1u << 29) - 1u)   /* wfa.go:190 */

oracle_aligner *oracle_new(const oracle_config *cfg);
void            oracle_free(oracle_aligner *a);

/* Align one pair (wfa.go:201-268).  *ops points into the aligner and stays
 * valid until the next call; it holds res->n_ops reversed+merged words
 * op<<32|n exactly like AlignmentResult.Ops after process(). */
int oracle_align(oracle_aligner *a, const uint8_t *q, uint32_t n,
                 const uint8_t *t, uint32_t m, oracle_result *res,
                 const uint64_t **ops, oracle_counters *ctr);

/* Wavefront inspection after oracle_align (the aligner keeps M/I/D like the
 * reference does for Plot).  comp: 0=M 1=I 2=D.  Returns 1 if present. */
int oracle_get_raw(const oracle_aligner *a, int comp, uint32_t s, int k, uint32_t *raw);
int oracle_krange(const oracle_aligner *a, int comp, uint32_t s, int *lo, int *hi);
uint32_t oracle_max_score(const oracle_aligner *a);

/* Batch, nthreads worker threads with one aligner each.  Same buffer layout
 * as wfacuda_align_batch.  ops may be NULL (results only).  ops_off[i] is the
 * start of pair i's ops; pairs are laid out in index order.  Returns 0, or
 * -1 if ops_capacity is too small (ops_needed then holds the total). */
int oracle_align_batch(const oracle_config *cfg, uint64_t n_pairs,
                       const uint8_t *seq_bytes,
                       const uint64_t *q_off, const uint32_t *q_len,
                       const uint64_t *t_off, const uint32_t *t_len,
                       oracle_result *results,
                       uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off,
                       uint64_t *ops_needed,
                       int nthreads, oracle_counters *ctr_sum);

#ifdef __cplusplus
}
#endif
#endif
