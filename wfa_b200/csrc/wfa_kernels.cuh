/*
 * wfa_kernels.cuh -- device code of libwfacuda.so (sm_100a).
 *
 * One "worker" aligns one pair at a time: a warp (WARP kernel: diagonals are
 * strided over the 32 lanes, the last max(x,o+e)/g+1 score rows of M and the
 * last e/g+1 rows of I/D live in a per-warp shared-memory ring, no block
 * barrier anywhere) or a whole CTA (CTA kernel: diagonals strided over the
 * block, source rows re-read from the HBM arena through L1/L2).  Workers are
 * persistent and pull pairs from an atomic queue sorted by decreasing cost.
 *
 * Every (score, diagonal) cell of M, I and D is written exactly once, as the
 * reference's raw word offset<<3|code, to the worker's slot of the HBM
 * backtrace arena; `next` and `extend` are fused per cell (the value stored
 * for M is the extended one, as after the reference's Increase).  Backtrace
 * and CIGAR emission follow the reference's control flow literally and run on
 * the same worker right after the forward pass.
 *
 * Semantics follow the reference at /root/reference (cited per function as
 * wfa.go:LINE etc.); equivalences used instead of a literal translation are
 * argued in DESIGN.md section 4.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace wfak {

/* wfa_backtrace_types.go:23-37 */
constexpr uint32_t T_BITS = 3, T_MASK = 7;
enum : uint32_t { T_INS_OPEN = 1, T_INS_EXT = 2, T_DEL_OPEN = 3, T_DEL_EXT = 4, T_MISMATCH = 5, T_MATCH = 6 };

/* internal per-pair states (>= 100 never leave the library) */
enum : uint8_t { ST_OK = 0, ST_EMPTY = 1, ST_TOO_LONG = 2, ST_RESOURCES = 3,
                 ST_ARENA = 100, ST_RING = 101, ST_OPS = 102, ST_NEED8 = 103, ST_PENDING = 255 };

struct PairDesc {
    uint64_t q_byte, t_byte;   /* byte offsets into the raw sequence pool */
    uint64_t q_word, t_word;   /* word offsets into the 2-bit packed pool */
    uint32_t n, m;             /* len(q), len(t) */
};

/* One score's wavefront row in a worker's arena slot.  Cells of M, I, D for
 * diagonals [alo, alo+aw) sit at word offsets off, off+aw, off+2*aw.  [lo,hi]
 * is the reference's M WaveFront.Lo/Hi *after* reduce: a cell outside it reads
 * as absent (that is what reduce's Delete calls achieve, wfa.go:526-537).
 * lo > hi means the score does not exist (HasScore false). */
struct RowHdr {
    int32_t  alo, aw, lo, hi;
    uint64_t off;
};

struct Result {                /* == wfacuda_result */
    uint32_t score;
    int32_t  tbegin, tend, qbegin, qend;
    uint32_t align_len, matches, gaps, gap_regions;
    uint32_t n_ops;
    uint8_t  status;
    uint8_t  pad_[3];
};

struct Counters {              /* device-side work counters, one set per run */
    unsigned long long cells, cells_written, steps, ops, ops_cursor;    /* accumulate over all launches of a run */
    /* --- reset before every class launch with ONE fill: [retry_n, launch_end) ------------------- */
    unsigned long long retry_n, work_next, arena_used_max;
    unsigned long long t_last;            /* globaltimer of the last block end (LANE kernels, profiling aid) */
    unsigned long long retry2_n;          /* hand-over launch (KParams.handover): its own failures, listed in its own retry buffer */
    unsigned long long handover_other;    /* hand-over launch: entries of the LANE list it left alone (not ST_RING) */
    unsigned long long lane_count[8];     /* LANE class: [j] pairs entering stage j, [4 + j] group queue of stage j */
    unsigned int lane_hist[64];           /* LANE class: sampled histogram of the final score index (next batch's stage boundaries) */
    unsigned long long launch_end;        /* (marker: end of the per-launch block) */
    /* ------------------------------------------------------------------------------------------- */
    unsigned long long t_first;           /* globaltimer of the first block start (only initialised under WFACUDA_DEBUG) */
    unsigned long long dump_rows;         /* single-worker launches: row headers the forward pass wrote (wfacuda_align_components) */
};

/* LANE class (wfa_lane.cuh).  Without heuristic the loop range of `next` depends only on which
 * scores exist, so the rows of every group have the same geometry: one table per launch. */
constexpr int LANE_MAX_STAGES = 4;
struct LaneGeom {
    int8_t   lo[64], hi[64];            /* loop range of the row of score index si; lo > hi: no such row */
    uint16_t off[64];                   /* first cell of the row inside its stage's group slot, in units of 32 words */
    int32_t  n_rows;                    /* rows in the table; a pair still running after row n_rows - 1 goes to the WARP worker */
    int32_t  n_stages;
    int32_t  stage_end[LANE_MAX_STAGES];/* last row of stage j; stage_end[n_stages - 1] = n_rows - 1 */
    uint32_t slot_words[LANE_MAX_STAGES];
    uint32_t scratch_words;             /* op scratch (backtrace overflow) at the start of every slot */
    uint32_t state_words;               /* words of a pair's saved state (ring bytes + 3 counters) */
};
struct LaneAux {
    uint32_t *rec;                      /* [ticket][LANE_REC_WORDS]: what the finish kernel needs of a pair */
    uint32_t *state;                    /* [ticket][state_words]: rings of a pair that moves on to the next stage */
    uint32_t *list[LANE_MAX_STAGES];    /* tickets entering stage j >= 1 */
    uint8_t  *arena[LANE_MAX_STAGES];   /* group slots of stage j */
    uint32_t  cap[LANE_MAX_STAGES];     /* groups the arena of stage j holds */
};

struct KParams {
    const PairDesc *pairs;
    const uint32_t *work;      /* pair indices, decreasing cost */
    uint32_t        n_work;
    uint32_t        pair_base; /* LANE kernels without a work list: item i is pair pair_base + i */
    const uint32_t *packed;    /* 2-bit pool */
    const uint32_t *raw;       /* byte pool viewed as words (base is 16-byte aligned) */
    const uint8_t  *pflags;    /* bit0: pair has a non-ACGT byte -> 8-bit path */
    uint8_t        *arena;
    uint64_t        slot_bytes;
    Result         *results;
    uint64_t       *ops_pool;  /* completion-order pool */
    uint64_t        ops_cap;
    uint64_t       *ops_where; /* per pair: start in ops_pool */
    uint64_t       *retry;     /* status<<32 | pair for pairs that ran out of arena / ring / pool */
    Counters       *ctr;
    /* penalties: raw and divided by g = gcd(x, o+e, e) */
    uint32_t x, oe, e, g;
    int32_t  xg, oeg, eg;
    int32_t  dM, dE;           /* ring depths: max(xg,oeg)+1, eg+1 */
    int32_t  ring_cap;         /* diagonals per ring row (WARP kernel) */
    int32_t  group;            /* WARP kernel: pairs per group (1..32), slot_bytes = group * sub-slot */
    int32_t  seq_cap;          /* WARP kernel: 32-bit words of shared memory per warp for the pair's 2-bit sequences (0: read them from global) */
    uint8_t  global_aln, adaptive, semi_literal;
    /* Hand-over launch of the WARP kernel: the work list is the retry list the LANE class just wrote
     * (status << 32 | pair, Counters.retry_n entries, only known on the device); entries with
     * ST_RING are aligned, the others are counted in Counters.handover_other and left to the host. */
    const uint64_t *handover;
    unsigned long long *retry_ctr;   /* where this launch counts its failures: &Counters.retry_n, or &Counters.retry2_n */
    uint8_t  single_worker;    /* only worker 0 takes work: its slot then holds the one pair's whole wavefront store (wfacuda_align_components) */
    int32_t  min_wf_len, max_dist_diff;
    LaneGeom lg;               /* LANE kernels only */
    LaneAux  la;
    /* WIDE kernels only (wfa_wide.cuh): diagonals per CTA of the cluster, 8-byte entries of shared memory for
     * the two sequence windows, where the forward pass leaves item i's outcome for the finish kernel */
    int32_t  wide_seg; uint32_t wide_seq_cap;
    struct FwdOut *wide_rec;
};

/* ------------------------------------------------------------------ sequences
 * 2-bit mode: 16 bases per 32-bit word, base i in bits 2(i%16)..; code =
 * (byte>>1)&3 (A=0 C=1 T=2 G=3), only used when every byte of the pair is one
 * of "ACGT" so that code equality == byte equality (wfa.go:408-454 compares
 * raw bytes).  8-bit mode reads the caller's bytes unchanged. */
template <int BITS, bool SM = false> struct SeqView {
    const uint32_t *w;   /* word holding symbol 0 (aligned down) */
    uint32_t        mis; /* symbols before symbol 0 inside that word (8-bit mode only) */
    uint32_t        sa;  /* SM: shared-window byte address of word 0 (the WARP worker's copy of a short sequence) */
    static constexpr int PER_WORD = 32 / BITS;
    __device__ __forceinline__ uint32_t word(uint32_t wi) const {
        if (SM) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa + wi * 4u)); return v; }
        return __ldg(w + wi);
    }
    __device__ __forceinline__ uint32_t sym(int i) const {
        uint32_t j = (uint32_t)i + mis;
        return (word(j / PER_WORD) >> ((j % PER_WORD) * BITS)) & ((1u << BITS) - 1u);
    }
    /* PER_WORD symbols starting at symbol i, symbol i in the low bits */
    __device__ __forceinline__ uint32_t chunk(int i) const {
        uint32_t j = (uint32_t)i + mis, wi = j / PER_WORD;
        return __funnelshift_r(word(wi), word(wi + 1), (j % PER_WORD) * BITS);
    }
};

/* Longest common prefix of q[v:] and t[h:], at most maxl symbols: exactly what
 * the reference's 8-byte block loop + byte loop compute (wfa.go:411-454),
 * here 16 bases (or 4 bytes) per XOR + find-first-set step. */
template <int BITS, bool SM>
__device__ __forceinline__ int lcp(const SeqView<BITS, SM> &Q, const SeqView<BITS, SM> &T, int v, int h, int maxl)
{
    constexpr int PW = 32 / BITS;
    int l = 0;
    while (l < maxl) {
        uint32_t x = Q.chunk(v + l) ^ T.chunk(h + l);
        if (x) { l += (__ffs((int)x) - 1) / BITS; break; }
        l += PW;
    }
    return l < maxl ? l : maxl;
}

/* Bases matched by a 16-base compare: half the number of trailing zero bits of the XOR of the two chunks, 16 when they
 * agree.  popc((x - 1) & ~x) counts the trailing zeros and is 32 for x = 0 by itself; the decrement goes through an asm
 * statement so that the compiler does not recognise a count-trailing-zeros and wrap it into its guarded sequence for
 * x = 0 (two more instructions in the innermost loop of every worker). */
__device__ __forceinline__ int matched_bases(uint32_t x)
{
    uint32_t t;
    asm("add.u32 %0, %1, -1;" : "=r"(t) : "r"(x));
    return __popc(t & ~x) >> 1;
}

/* ------------------------------------------------------------------ next
 * One diagonal of wfa.go:572-699.  Inputs are raw source words (0 = absent):
 *   mo_l = M[s-o-e][k-1]  ie_l = I[s-e][k-1]
 *   mo_r = M[s-o-e][k+1]  de_r = D[s-e][k+1]   mx = M[s-x][k]
 * Validity uses '>' exactly like the reference (:581,:585,:616,:620,:651), so
 * the one-past-the-end phantom cells are produced identically. */
struct Cell3 { uint32_t M, I, D; };

/* Every candidate is packed as value<<p | priority so that one max() picks both the offset
 * and, on equal offsets, the provenance the reference prefers:
 *   I:  M[s-o-e][k-1] (Open) beats I[s-e][k-1] (Ext) on ties   -- "v1 >= v2" (:592-596)
 *   D:  M[s-o-e][k+1] (Open) beats D[s-e][k+1] (Ext) on ties   -- (:628-632)
 *   M:  Mismatch > I > D on ties                                -- (:656-693)
 * A source is valid iff raw != 0 and its offset is within the bound; with raw = off<<3|type,
 * off in [1, bound]  <=>  (raw - 8) < (bound << 3) as unsigned (raw == 0 wraps to huge). */
__device__ __forceinline__ Cell3 next_cell(uint32_t mo_l, uint32_t ie_l, uint32_t mo_r, uint32_t de_r,
                                           uint32_t mx, int k, int n, int m)
{
    Cell3 r;
    const uint32_t bm = (uint32_t)m << T_BITS;                    /* offset <= m         (:581,:585,:651) */
    /* offset - k <= n (:616,:620,:651); n + k >= 1; offsets never exceed 2^29, so the bound saturates there */
    const uint32_t bk = min((uint32_t)(n + k), 0x1fffffffu) << T_BITS;
    /* insertion (:579-609) */
    uint32_t ca = (mo_l - 8u) < bm ? ((mo_l >> T_BITS) << 1 | 1u) : 0u;
    uint32_t cb = (ie_l - 8u) < bm ? ((ie_l >> T_BITS) << 1) : 0u;
    uint32_t best = max(ca, cb);
    const uint32_t Isk = best ? (best >> 1) + 1u : 0u;
    const uint32_t tI = T_INS_EXT - (best & 1u);                  /* 1 = Open, 2 = Ext */
    r.I = best ? (Isk << T_BITS | tI) : 0u;
    /* deletion (:614-645) */
    ca = (mo_r - 8u) < bk ? ((mo_r >> T_BITS) << 1 | 1u) : 0u;
    cb = (de_r - 8u) < bk ? ((de_r >> T_BITS) << 1) : 0u;
    best = max(ca, cb);
    const uint32_t Dsk = best >> 1;
    const uint32_t tD = T_DEL_EXT - (best & 1u);                  /* 3 = Open, 4 = Ext */
    r.D = best ? (Dsk << T_BITS | tD) : 0u;
    /* mismatch + provenance priority (:650-698).  Msk = max(Isk, Dsk, v1+1); when M[s-x][k] is
     * not usable v1+1 = 1 never beats an existing I (>= 2) or D (>= 1) */
    const uint32_t cx = (mx - 8u) < min(bm, bk) ? (((mx >> T_BITS) + 1u) << 2 | 2u) : 0u;
    const uint32_t ci = Isk ? (Isk << 2 | 1u) : 0u;
    const uint32_t cd = Dsk << 2;                                 /* Dsk >= 1 when present */
    const uint32_t bestM = max(max(cx, ci), cd);
    const uint32_t src = bestM & 3u;
    const uint32_t tM = src == 2u ? T_MISMATCH : (src == 1u ? tI : tD);
    r.M = bestM ? ((bestM >> 2) << T_BITS | tM) : 0u;
    return r;
}

/* distance-to-end of one M cell for reduce, -1 = not counted (wfa.go:474-494) */
__device__ __forceinline__ int dist_of(uint32_t raw, int k, int n, int m)
{
    if (raw == 0) return -1;
    const int h = (int)(raw >> T_BITS), v = h - k;
    if (v < 0 || v >= n || h >= m) return -1;
    return max(m - h, n - v);
}

/* start-cell classification for semi-global (wfa.go:306-323 / :341-358):
 * 0 = keep scanning, 1 = scan stops without a hit, 2 = hit */
__device__ __forceinline__ int hit_class(uint32_t raw, int k, int n, int m)
{
    if (raw == 0) return 0;
    const int h = (int)(raw >> T_BITS), v = h - k;
    if (v <= 0 || v > n || h > m) return 1;
    if ((v == n && h >= n) || (h == m && v >= m)) return 2;
    return 0;
}

/* ------------------------------------------------------------------ groups
 * A worker is a warp (CTA=false) or a block (CTA=true). */
template <bool CTA> struct Grp {
    __device__ static __forceinline__ int tid()  { return CTA ? (int)threadIdx.x : (int)(threadIdx.x & 31); }
    __device__ static __forceinline__ int size() { return CTA ? (int)blockDim.x : 32; }
    __device__ static __forceinline__ void sync() { if (CTA) __syncthreads(); else __syncwarp(); }
    /* all-reduce of (min a, max b, or c) over the group; red = 4*32 ints of block scratch */
    __device__ static __forceinline__ void reduce3(int &a, int &b, int &c, int *red)
    {
        a = __reduce_min_sync(0xffffffffu, a);
        b = __reduce_max_sync(0xffffffffu, b);
        c = (int)__reduce_or_sync(0xffffffffu, (unsigned)c);
        if (CTA) {
            const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
            __syncthreads();                     /* scratch free again */
            if (lane == 0) { red[wid] = a; red[32 + wid] = b; red[64 + wid] = c; }
            __syncthreads();
            a = lane < nw ? red[lane] : INT_MAX;
            b = lane < nw ? red[32 + lane] : INT_MIN;
            c = lane < nw ? red[64 + lane] : 0;
            a = __reduce_min_sync(0xffffffffu, a);
            b = __reduce_max_sync(0xffffffffu, b);
            c = (int)__reduce_or_sync(0xffffffffu, (unsigned)c);
        }
    }
    template <typename T> __device__ static __forceinline__ T bcast0(T v, T *slot)
    {
        if (CTA) {
            __syncthreads();
            if (threadIdx.x == 0) *slot = v;
            __syncthreads();
            return *slot;
        } else {
            return __shfl_sync(0xffffffffu, v, 0);
        }
    }
};

/* ------------------------------------------------------------------ arena reads for backtrace
 * Component.Get / GetRaw (wfa_component.go:142-155) on the arena.  `s` is the
 * reference's uint32 score arithmetic carried in 64 bits: a wrapped value
 * (s - diff with diff > s) is >= len(WaveFronts) there, i.e. absent. */
struct ArenaView {
    const RowHdr   *hdr;
    const uint32_t *cells;
    int             si_last;   /* index (score / g) of the highest score with a header */
    bool            aos;       /* row layout: cell-major M,I,D triples (WARP worker) or three arrays (CTA worker) */
    /* si = score / g; every score reachable by the backtrace is a multiple of g */
    __device__ __forceinline__ uint32_t get(int comp, int si, int k) const
    {
        if (si < 0 || si > si_last) return 0;
        const RowHdr h = hdr[si];
        if (k < h.lo || k > h.hi) return 0;
        const uint32_t d = (uint32_t)(k - h.alo);
        return cells[h.off + (aos ? 3ull * d + (uint32_t)comp : (uint64_t)comp * (uint32_t)h.aw + d)];
    }
    /* the backtrace reads the provenance code only of the cell it stands on; views that do
     * not store codes re-derive them there (LaneView) */
    __device__ __forceinline__ uint32_t get_typed(int comp, int si, int k) const { return get(comp, si, k); }
};

/* Reversed, run-merged op emitter: AddN (wfa_cigar.go:118-124) followed by
 * process()'s merge of equal neighbours (:147-166); merging commutes with the
 * reversal, so runs are merged as they are produced. */
struct OpSink {
    uint64_t *buf; uint32_t cap, n; uint32_t cur_op, cur_n; bool overflow;
    uint32_t stride;               /* words between consecutive ops (1, or 32 for lane-interleaved scratch) */
    __device__ __forceinline__ void add(uint32_t op, uint32_t cnt)
    {
        if (op == cur_op) { cur_n += cnt; return; }
        flush();
        cur_op = op; cur_n = cnt;
    }
    __device__ __forceinline__ void flush()
    {
        if (cur_op == 0) return;
        if (n < cap) buf[(size_t)n * stride] = (uint64_t)cur_op << 32 | cur_n; else overflow = true;
        n++;
        cur_op = 0;
    }
};

__device__ __forceinline__ uint32_t op_of_type(uint32_t t)
{
    /* wfaOps = ". I I D D X M H" (wfa_backtrace_types.go:37) */
    return __byte_perm(0x4449492eu, 0x484d5844u, t & 7u) & 0xffu;     /* one PRMT instead of an indexed constant load */
}

/* backTrace, wfa.go:703-983, executed by one thread.  Returns ops (reversed
 * order, merged) in sink; fills score/begin/end of res. */
template <class View, class Sink>
__device__ __forceinline__ void back_trace_inl(View &A, const KParams &P, int n, int m,
                                               uint32_t s0, int Ak, Result &res, Sink &sink)
{
    const bool semi = !P.global_aln;
    /* scores are carried as indices s/g: the reference's uint32 s-x etc. (wfa.go:760-762)
     * wrap to "absent" exactly when the index goes negative */
    const int x = P.xg, oe = P.oeg, e = P.eg;
    int s = (int)(s0 / P.g);
    res.score = s0;
    res.tbegin = res.tend = res.qbegin = res.qend = 0;
    int k = Ak, h, v, tBegin = 0, qBegin = 0;
    bool firstMatch = true, previousFromM = true, fromItself = false;
    uint32_t offset0 = 0, Isk = 0, Dsk = 0;
    int M0 = 0;                                   /* component of the next cell: 0 M, 1 I, 2 D */

    uint32_t raw = A.get_typed(0, s, k);          /* :738; get_typed = raw word incl. its provenance code */
    uint32_t type = raw & T_MASK;
    h = (int)(raw >> T_BITS);
    v = h - k;
    if (h < m) sink.add('I', (uint32_t)m - (uint32_t)h);          /* :746-750 */
    else if (v < n) sink.add('H', (uint32_t)n - (uint32_t)v);

    while (v > 0 && h > 0) {                      /* :753 */
        const int sX = s - x, sO = s - oe, sE = s - e;
        if (type == T_INS_EXT) {                  /* :767-777 */
            const uint32_t r1 = A.get(0, sO, k - 1), r2 = A.get(1, sE, k - 1);
            offset0 = (r1 | r2) ? max(r1 >> T_BITS, r2 >> T_BITS) + 1 : 0;
            M0 = 1;
        } else if (type == T_DEL_EXT) {           /* :778-788 */
            const uint32_t r1 = A.get(0, sO, k + 1), r2 = A.get(2, sE, k + 1);
            offset0 = (r1 | r2) ? max(r1 >> T_BITS, r2 >> T_BITS) : 0;
            M0 = 2;
        } else {                                  /* :789-817 */
            const uint32_t r1 = A.get(0, sO, k - 1), r2 = A.get(1, sE, k - 1);
            const uint32_t r3 = A.get(0, sO, k + 1), r4 = A.get(2, sE, k + 1);
            const uint32_t r5 = A.get(0, sX, k);
            Isk = (r1 | r2) ? max(r1 >> T_BITS, r2 >> T_BITS) + 1 : 0;
            Dsk = (r3 | r4) ? max(r3 >> T_BITS, r4 >> T_BITS) : 0;
            if (r1 | r2 | r3 | r4 | r5) { offset0 = max(max(Isk, Dsk), (r5 >> T_BITS) + 1); fromItself = false; }
            else fromItself = true;
            M0 = 0;
        }
        if (fromItself) break;                    /* :818-825 */
        if (offset0 == 0) break;
        const int h0 = (int)offset0;
        if (previousFromM) {                      /* :833-869 */
            const int nMatches = h - h0;
            if (nMatches > 0) {
                if (firstMatch) { firstMatch = false; res.tend = h; res.qend = v; }
                sink.add('M', (uint32_t)nMatches);
            }
            h = h0; v = h - k;
            if (type == T_MATCH) { tBegin = h; qBegin = v; }
            else if (nMatches > 0) { tBegin = h + 1; qBegin = v + 1; }
            if (h <= 0 || v <= 0) break;
        }
        sink.add(op_of_type(type), 1);            /* :872-873 */
        if (semi && (h == 1 || v == 1)) break;    /* :876-879 */
        previousFromM = true;                     /* :885-909 */
        bool leave = false;
        switch (type) {
        case T_MISMATCH: s = sX; h--; break;
        case T_INS_OPEN: s = sO; k--; h--; break;
        case T_INS_EXT:  s = sE; k--; h--; previousFromM = false; break;
        case T_DEL_OPEN: s = sO; k++; break;
        case T_DEL_EXT:  s = sE; k++; previousFromM = false; break;
        default: leave = true;
        }
        if (leave) break;
        v = h - k;
        raw = A.get_typed(M0, s, k);              /* :915-920 */
        if (raw == 0) break;
        type = raw & T_MASK;
    }

    if (h > 0 && v > 0) {                         /* :930-968 */
        const int nMatches = min(h, v) - 1;
        if (nMatches > 0) {
            if (firstMatch) { firstMatch = false; res.tend = h; res.qend = v; }
            sink.add('M', (uint32_t)nMatches);
            h -= nMatches; v -= nMatches;
            if (type == T_MATCH) { tBegin = h; qBegin = v; }
            else { tBegin = h + 1; qBegin = v + 1; }
        } else if (type == T_MATCH) {
            tBegin = h; qBegin = v;
            if (firstMatch) { firstMatch = false; res.tend = h; res.qend = v; }
        }
        sink.add(op_of_type(type), 1);
    }
    if (v > 1) sink.add('H', (uint32_t)(v - 1));  /* :970-976 */
    if (h > 1) sink.add('I', (uint32_t)(h - 1));
    res.tbegin = tBegin; res.qbegin = qBegin;     /* :979 */
    sink.flush();
}

/* out-of-line copy for the WARP / CTA workers (their kernels are register-bound in the forward pass) */
template <class View>
__device__ __noinline__ void back_trace(View &A, const KParams &P, int n, int m,
                                        uint32_t s0, int Ak, Result &res, OpSink &sink)
{
    back_trace_inl(A, P, n, m, s0, Ak, res, sink);
}

/* ------------------------------------------------------------------ explicit shared-memory access
 * The WARP worker addresses its ring with 32-bit shared-window byte addresses kept in
 * registers (made opaque to the compiler, which otherwise re-derives them from the kernel
 * parameters inside the diagonal loop). */
template <int OFF> __device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "r"(v) : "memory");
}
__device__ __forceinline__ void keep(uint32_t &x) { asm volatile("" : "+r"(x)); }
__device__ __forceinline__ void keep(int &x) { asm volatile("" : "+r"(x)); }
template <typename T> __device__ __forceinline__ void keep_ptr(T *&p) { asm volatile("" : "+l"(p)); }

/* ------------------------------------------------------------------ one pair
 * Shared-memory layout of a worker:
 *   meta  int4[dM]           ring of the most recent rows' {alo, lo, hi, aw}
 *   roff  u64[dM]            their arena offsets (used by the CTA worker)
 *   red   int[128]           block reduction scratch (CTA only)
 *   bslot u64[2]             broadcast scratch
 *   ring  u32[dM][cap][3]    WARP only: ring of rows, cell-major {M, I, D} triples
 */
template <bool CTA> __host__ __device__ inline size_t worker_smem_bytes(int dM, int dE, int cap, int seq_cap = 0)
{
    size_t b = (size_t)dM * 16 + (size_t)dM * 8 + 128 * sizeof(int) + 16;
    if (!CTA) b += (size_t)dM * (size_t)cap * 12 + (size_t)seq_cap * 4;      /* ring, then the pair's sequences */
    return (b + 15) & ~(size_t)15;
}

/* what the forward pass leaves for the backtrace (uniform over the worker) */
struct FwdOut {
    int      status;
    uint32_t minS;            /* score the backtrace starts from */
    int      lastK, si;       /* its diagonal; index of the last header written */
    int      n, m;
    uint64_t top;             /* first used cell word of the slot (rows occupy [top, slot_words)) */
    unsigned long long c_cells, c_written, c_steps;
    bool     first_eq;        /* REG worker: q[0] == t[0] (which init cell exists, wfa.go:155-158) */
};

template <int BITS, bool CTA, bool SSEQ = false>
__device__ FwdOut forward_pair(const KParams &P, const uint32_t pair, unsigned char *smem, uint8_t *slot, const uint64_t slot_bytes)
{
    using G = Grp<CTA>;
    const int tid = G::tid(), gsz = G::size();
    const PairDesc pd = P.pairs[pair];
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const int dM = P.dM, cap = P.ring_cap;

    SeqView<BITS, SSEQ> Q, T;
    Q.sa = T.sa = 0;
    if (BITS == 2) { Q.w = P.packed + pd.q_word; Q.mis = 0; T.w = P.packed + pd.t_word; T.mis = 0; }
    else { Q.w = P.raw + (pd.q_byte >> 2); Q.mis = (uint32_t)(pd.q_byte & 3); T.w = P.raw + (pd.t_byte >> 2); T.mis = (uint32_t)(pd.t_byte & 3); }

    int4     *meta  = reinterpret_cast<int4 *>(smem);
    uint64_t *roff  = reinterpret_cast<uint64_t *>(meta + dM);
    int      *red   = reinterpret_cast<int *>(roff + dM);
    uint64_t *bslot = reinterpret_cast<uint64_t *>(red + 128);         /* 16 bytes of broadcast scratch */
    uint32_t *ring = reinterpret_cast<uint32_t *>(bslot + 2);
    /* Row layout.  WARP worker: cell-major triples {M,I,D} in the ring and in the arena, so one
     * address register per row serves all three components at immediate offsets.  CTA worker:
     * three arrays per row (its source rows are re-read from L2, where triples would triple the
     * sectors touched by the M-only reads). */
    constexpr int KS = CTA ? 1 : 3;                                     /* words between neighbouring diagonals */
    uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring);
    keep(ring_sa);
    if (SSEQ) {
        /* short pair: its 2-bit words live in the warp's shared memory for the whole forward
         * pass (extend reads two words per sequence and step); one spare word each for the
         * funnel shift */
        const uint32_t wq = ((uint32_t)n + 15u) >> 4, wt = ((uint32_t)m + 15u) >> 4;
        uint32_t *sq = ring + (size_t)dM * cap * 3, *st = sq + wq + 1;
        for (uint32_t i = tid; i <= wq; i += 32) sq[i] = i < wq ? __ldg(Q.w + i) : 0u;
        for (uint32_t i = tid; i <= wt; i += 32) st[i] = i < wt ? __ldg(T.w + i) : 0u;
        Q.sa = (uint32_t)__cvta_generic_to_shared(sq); T.sa = (uint32_t)__cvta_generic_to_shared(st);
        __syncwarp();
    }

    RowHdr   *hdrs  = reinterpret_cast<RowHdr *>(slot);                /* grows up, index s/g */
    uint32_t *cells = reinterpret_cast<uint32_t *>(slot);              /* rows grow down from the end */
    /* word offsets inside the slot: 32 bits are enough for a warp's slot (the host keeps
     * WARP-class slots below 16 GB), the CTA worker may own more */
    typedef typename std::conditional<CTA, uint64_t, uint32_t>::type Off;
    const Off slot_words = (Off)(slot_bytes >> 2);
    Off top = slot_words;
    Off hdr_limit = 2 * (sizeof(RowHdr) / 4) + 8;                      /* words used by headers incl. the next one + slack */

    const int4 EMPTY = make_int4(0, 1, 0, 0);
    for (int i = tid; i < dM; i += gsz) meta[i] = EMPTY;
    G::sync();

    const int x = (int)P.x, xg = P.xg, oeg = P.oeg, eg = P.eg;
    const int ilo = P.global_aln ? 0 : -(n - 1), ihi = P.global_aln ? 0 : m - 1;   /* init cells, wfa.go:160-183 */
    const int maxdiff = P.max_dist_diff;

    int status = ST_OK;
    uint32_t s = 0; int si = 0, cur = 0;
    uint32_t minS = 0; int lastK = Ak;
    typedef typename std::conditional<CTA, unsigned long long, uint32_t>::type Cnt;   /* per-pair work counters */
    Cnt c_cells = 0, c_written = 0, c_steps = 0;

    /* ---------------- forward: wfa.go:228-251 with next+extend fused per cell */
    for (;;) {
        /* source rows s-x, s-o-e, s-e: ring slots kept incrementally (no division) */
        int slX = cur - xg, slO = cur - oeg, slE = cur - eg;
        slX += slX < 0 ? dM : 0; slO += slO < 0 ? dM : 0; slE += slE < 0 ? dM : 0;
        /* slots this pair has not written yet still hold EMPTY from the start of the pair */
        const int4 hX = meta[slX], hO = meta[slO], hE = meta[slE];
        /* loop range (wfa.go:557-563); a superset is harmless, the clamp is not */
        int lo = INT_MAX, hi = INT_MIN;
        if (hX.y <= hX.z) { lo = min(lo, hX.y); hi = max(hi, hX.z); }
        if (hO.y <= hO.z) { lo = min(lo, hO.y); hi = max(hi, hO.z); }
        if (hE.y <= hE.z) { lo = min(lo, hE.y); hi = max(hi, hE.z); }
        if (lo <= hi) { lo = max(lo - 1, -(n - 1)); hi = min(hi + 1, m - 1); }
        const bool has_init = (s == 0) || (s == (uint32_t)x);
        if (has_init) { lo = min(lo, ilo); hi = max(hi, ihi); }

        bool exists = false;
        int wlo = INT_MAX, whi = INT_MIN, endhit = 0;
        int aw = 0; Off off = 0;
        /* WARP worker, row of at most 64 diagonals (the usual wf-adaptive row): the lane's one or
         * two cells stay in registers, so that Lo/Hi, the end test and reduce need no second look
         * at the row */
        uint32_t myM0 = 0, myM1 = 0;
        const bool narrow = !CTA && lo <= hi && hi - lo < 64;
        if (lo <= hi) {
            aw = hi - lo + 1;
            if (!CTA && aw > cap) { status = ST_RING; break; }
            const Off need = (Off)3 * (Off)aw;
            if (top < hdr_limit || top - hdr_limit < need) { status = ST_ARENA; break; }
            off = top - need;
            const uint32_t cntX = (uint32_t)max(hX.z - hX.y + 1, 0), cntO = (uint32_t)max(hO.z - hO.y + 1, 0), cntE = (uint32_t)max(hE.z - hE.y + 1, 0);
            /* one diagonal after its five source words are in: init overlay, extend, bookkeeping */
            auto finish_cell = [&](Cell3 &c, const int k) {
                if (has_init && c.M == 0 && k >= ilo && k <= ihi) {
                    /* initComponents (wfa.go:155-183): cell (k) of the first row / column;
                     * next's Set overwrites it when both write (wfa_wavefront.go:93) */
                    const bool eq = Q.sym(k < 0 ? -k : 0) == T.sym(k > 0 ? k : 0);
                    if (eq ? (s == 0) : (s == (uint32_t)x))
                        c.M = (uint32_t)((k > 0 ? k : 0) + 1) << T_BITS | (eq ? T_MATCH : T_MISMATCH);
                }
                if (c.M) {
                    /* extend (wfa.go:394-455) */
                    int h = (int)(c.M >> T_BITS), v = h - k;
                    if (v > 0 && v < n && h < m) {
                        const int l = lcp(Q, T, v, h, min(n - v, m - h));
                        c.M += (uint32_t)l << T_BITS;
                        h += l;
                    }
                    wlo = min(wlo, k); whi = max(whi, k);
                    if (k == Ak && h >= m) endhit = 1;            /* wfa.go:235-239 */
                }
            };
            if (CTA) {
                /* source rows from the arena (three arrays per row), diagonal k indexes directly.
                 * The block is latency-bound on these L2/HBM reads, so each thread handles U
                 * diagonals per pass: all 5*U source words are requested first, then the 4*U
                 * sequence words of the first extend step, and only then anything is consumed. */
                const uint64_t oX = roff[slX], oO = roff[slO], oE = roff[slE];
                const uint32_t *srcX = cells + oX - hX.x, *srcO = cells + oO - hO.x;
                const uint32_t *srcI = cells + oE + (uint32_t)hE.w - hE.x, *srcD = cells + oE + 2ull * (uint32_t)hE.w - hE.x;
                uint32_t *dstM = cells + off - lo, *dstI = dstM + aw, *dstD = dstI + aw;
                constexpr int U = 4;
                constexpr int PW = 32 / BITS;
                for (int k0 = lo + tid; k0 <= hi; k0 += U * gsz) {
                    uint32_t w[U][5];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int k = k0 + u * gsz;
                        w[u][0] = w[u][1] = w[u][2] = w[u][3] = w[u][4] = 0;
                        if (k <= hi) {
                            if ((uint32_t)(k - 1 - hO.y) < cntO) w[u][0] = srcO[k - 1];
                            if ((uint32_t)(k - 1 - hE.y) < cntE) w[u][1] = srcI[k - 1];
                            if ((uint32_t)(k + 1 - hO.y) < cntO) w[u][2] = srcO[k + 1];
                            if ((uint32_t)(k + 1 - hE.y) < cntE) w[u][3] = srcD[k + 1];
                            if ((uint32_t)(k - hX.y) < cntX) w[u][4] = srcX[k];
                        }
                    }
                    Cell3 c[U]; int ext[U]; uint32_t xr[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int k = k0 + u * gsz;
                        c[u] = next_cell(w[u][0], w[u][1], w[u][2], w[u][3], w[u][4], k, n, m);
                        ext[u] = 0; xr[u] = 0;
                        if (k <= hi) {
                            if (has_init && c[u].M == 0 && k >= ilo && k <= ihi) {
                                const bool eq = Q.sym(k < 0 ? -k : 0) == T.sym(k > 0 ? k : 0);
                                if (eq ? (s == 0) : (s == (uint32_t)x))
                                    c[u].M = (uint32_t)((k > 0 ? k : 0) + 1) << T_BITS | (eq ? T_MATCH : T_MISMATCH);
                            }
                            const int h = (int)(c[u].M >> T_BITS), v = h - k;
                            if (c[u].M && v > 0 && v < n && h < m) {            /* extend applies (wfa.go:404) */
                                ext[u] = min(n - v, m - h);
                                xr[u] = Q.chunk(v) ^ T.chunk(h);                 /* first 16 bases / 4 bytes */
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int k = k0 + u * gsz;
                        if (k > hi) continue;
                        if (c[u].M) {
                            int h = (int)(c[u].M >> T_BITS);
                            if (ext[u]) {
                                int l;
                                if (xr[u]) l = (__ffs((int)xr[u]) - 1) / BITS;
                                else l = PW + (ext[u] > PW ? lcp(Q, T, h - k + PW, h + PW, ext[u] - PW) : 0);
                                l = min(l, ext[u]);
                                c[u].M += (uint32_t)l << T_BITS;
                                h += l;
                            }
                            wlo = min(wlo, k); whi = max(whi, k);
                            if (k == Ak && h >= m) endhit = 1;            /* wfa.go:235-239 */
                        }
                        dstM[k] = c[u].M; dstI[k] = c[u].I; dstD[k] = c[u].D;
                    }
                }
            } else {
                /* source rows from the shared-memory ring: cell-major {M,I,D} triples, 12 bytes per
                 * diagonal; one byte address per row and lane, advanced by 32 diagonals per pass */
                int k = lo + tid;
                uint32_t pO = ring_sa + (uint32_t)((slO * cap - hO.x + k) * 12);
                uint32_t pE = ring_sa + (uint32_t)((slE * cap - hE.x + k) * 12);
                uint32_t pX = ring_sa + (uint32_t)((slX * cap - hX.x + k) * 12);
                uint32_t pC = ring_sa + (uint32_t)((cur * cap - lo + k) * 12);
                uint32_t *gC = cells + off + 3 * (k - lo);
                int rO = k - 1 - hO.y, rE = k - 1 - hE.y, rX = k - hX.y;      /* unsigned range tests: r < cnt */
                keep(pO); keep(pE); keep(pX); keep(pC); keep_ptr(gC); keep(rO); keep(rE); keep(rX);
                for (int pass = 0; k <= hi; k += 32, pass++) {
                    uint32_t mo_l = 0, ie_l = 0, mo_r = 0, de_r = 0, mx = 0;
                    if ((uint32_t)rO < cntO) mo_l = lds32<-12>(pO);
                    if ((uint32_t)(rO + 2) < cntO) mo_r = lds32<12>(pO);
                    if ((uint32_t)rE < cntE) ie_l = lds32<-12 + 4>(pE);
                    if ((uint32_t)(rE + 2) < cntE) de_r = lds32<12 + 8>(pE);
                    if ((uint32_t)rX < cntX) mx = lds32<0>(pX);
                    Cell3 c = next_cell(mo_l, ie_l, mo_r, de_r, mx, k, n, m);
                    finish_cell(c, k);
                    if (pass == 0) myM0 = c.M; else if (pass == 1) myM1 = c.M;
                    sts32<0>(pC, c.M); sts32<4>(pC, c.I); sts32<8>(pC, c.D);
                    gC[0] = c.M; gC[1] = c.I; gC[2] = c.D;
                    pO += 384; pE += 384; pX += 384; pC += 384; gC += 96; rO += 32; rE += 32; rX += 32;
                }
            }
            if (narrow) {
                /* lane j holds diagonals lo + j and lo + 32 + j: first / last present cell from two ballots */
                const unsigned long long pres = (unsigned long long)__ballot_sync(0xffffffffu, myM0 != 0) |
                                                (unsigned long long)__ballot_sync(0xffffffffu, myM1 != 0) << 32;
                endhit = __any_sync(0xffffffffu, endhit != 0);
                if (pres) { wlo = lo + __ffsll((long long)pres) - 1; whi = lo + 63 - __clzll((long long)pres); }
                else { wlo = INT_MAX; whi = INT_MIN; }
            } else G::reduce3(wlo, whi, endhit, red);
            exists = wlo <= whi;
        }

        RowHdr hc; hc.alo = lo; hc.aw = aw; hc.lo = 1; hc.hi = 0; hc.off = off;
        if (!exists) {
            hc.alo = 0; hc.aw = 0; hc.off = 0;
            G::sync();
            if (tid == 0) { meta[cur] = EMPTY; hdrs[si] = hc; }
            G::sync();
            s += P.g; si++; hdr_limit += sizeof(RowHdr) / 4;
            cur = cur + 1 == dM ? 0 : cur + 1;
            continue;
        }
        top = off;
        c_steps++; c_cells += (Cnt)(whi - wlo + 1); c_written += (Cnt)aw;
        int elo = wlo, ehi = whi;
        const uint32_t *rowM;
        if (CTA) { G::sync(); rowM = cells + off; }                     /* arena row visible to the block */
        else { __syncwarp(); rowM = ring + cur * cap * 3; }

        bool finished = endhit != 0;
        if (!finished && P.adaptive && whi - wlo + 1 >= P.min_wf_len) {
            /* reduce (wfa.go:461-540) as three group reductions; see DESIGN.md 4.3 */
            if (narrow) {
                /* the same three reductions on the lanes' own cells: min distance by redux, the
                 * first / last near diagonal and the last valid one below it from ballots */
                const int d0 = dist_of(myM0, lo + tid, n, m), d1 = dist_of(myM1, lo + 32 + tid, n, m);   /* no cell: -1 */
                const int mind = __reduce_min_sync(0xffffffffu, min(d0 >= 0 ? d0 : INT_MAX, d1 >= 0 ? d1 : INT_MAX));
                const unsigned long long valid = (unsigned long long)__ballot_sync(0xffffffffu, d0 >= 0) |
                                                 (unsigned long long)__ballot_sync(0xffffffffu, d1 >= 0) << 32;
                const unsigned long long far = (unsigned long long)__ballot_sync(0xffffffffu, d0 >= 0 && d0 - mind > maxdiff) |
                                               (unsigned long long)__ballot_sync(0xffffffffu, d1 >= 0 && d1 - mind > maxdiff) << 32;
                if (far) {
                    const unsigned long long near = valid & ~far;           /* never empty: the closest cell is near */
                    const int fb = __ffsll((long long)near) - 1, Lb = 63 - __clzll((long long)near);
                    const unsigned long long below = valid & ((1ull << fb) - 1ull);
                    if (below) elo = lo + (63 - __clzll((long long)below)) + 1;
                    ehi = lo + Lb;
                }
            } else {
            int mind = INT_MAX, dummy1 = INT_MIN, dummy2 = 0;
            for (int k = wlo + tid; k <= whi; k += gsz) {
                const int d = dist_of(rowM[KS * (k - lo)], k, n, m);
                if (d >= 0) mind = min(mind, d);
            }
            G::reduce3(mind, dummy1, dummy2, red);
            int f = INT_MAX, L = INT_MIN, anyfar = 0;
            for (int k = wlo + tid; k <= whi; k += gsz) {
                const int d = dist_of(rowM[KS * (k - lo)], k, n, m);
                if (d >= 0) { if (d - mind > maxdiff) anyfar = 1; else { f = min(f, k); L = max(L, k); } }
            }
            G::reduce3(f, L, anyfar, red);
            if (anyfar) {
                int lf = INT_MIN, d0 = INT_MAX, d2 = 0;
                for (int k = wlo + tid; k <= whi && k < f; k += gsz)
                    if (dist_of(rowM[KS * (k - lo)], k, n, m) >= 0) lf = max(lf, k);
                G::reduce3(d0, lf, d2, red);
                if (lf != INT_MIN) elo = lf + 1;
                ehi = L;
            }
            }
        }
        bool hit = false; int hitK = Ak;
        if (!P.global_aln && (finished || !P.semi_literal)) {
            /* backtraceStartPosistion (wfa.go:270-375) for this one score, on the row as the
             * reference leaves it (post-reduce, or un-reduced when it is the final score) */
            int ka = INT_MIN, kb = INT_MAX, d2 = 0;
            const int a_hi = min(Ak, ehi), b_lo = max(Ak + 1, elo);
            for (int k = elo + tid; k <= ehi; k += gsz) {
                const int c = hit_class(rowM[KS * (k - lo)], k, n, m);
                if (c) {
                    const int key = (k + n) * 2 + (c == 2);
                    if (k <= a_hi) ka = max(ka, key);
                    if (k >= b_lo) kb = min(kb, key);
                }
            }
            G::reduce3(kb, ka, d2, red);
            if (ka != INT_MIN && (ka & 1)) { hit = true; hitK = (ka >> 1) - n; }
            if (kb != INT_MAX && (kb & 1)) { hit = true; hitK = (kb >> 1) - n; }     /* scan (b) overrides (a) */
        }
        hc.lo = elo; hc.hi = ehi;
        G::sync();
        if (tid == 0) { meta[cur] = make_int4(lo, elo, ehi, aw); if (CTA) roff[cur] = (uint64_t)off; hdrs[si] = hc; }
        G::sync();
        if (finished) { minS = s; lastK = Ak; if (hit) lastK = hitK; break; }
        if (hit) { minS = s; lastK = hitK; break; }
        s += P.g; si++; hdr_limit += sizeof(RowHdr) / 4;
        cur = cur + 1 == dM ? 0 : cur + 1;
    }

    /* ---------------- semi-global, literal mode: scan every retained score downwards (wfa.go:287-371) */
    if (status == ST_OK && !P.global_aln && P.semi_literal) {
        minS = s; lastK = Ak;
        for (int sj = si; sj >= 0; sj--) {
            const RowHdr h = hdrs[sj];
            if (h.lo > h.hi) continue;
            const uint32_t *rowM = cells + h.off;
            int ka = INT_MIN, kb = INT_MAX, d2 = 0;
            const int a_hi = min(Ak, h.hi), b_lo = max(Ak + 1, h.lo);
            for (int k = h.lo + tid; k <= h.hi; k += gsz) {
                const int c = hit_class(rowM[KS * (k - h.alo)], k, n, m);
                if (c) {
                    const int key = (k + n) * 2 + (c == 2);
                    if (k <= a_hi) ka = max(ka, key);
                    if (k >= b_lo) kb = min(kb, key);
                }
            }
            G::reduce3(kb, ka, d2, red);
            if (ka != INT_MIN && (ka & 1)) { minS = (uint32_t)sj * P.g; lastK = (ka >> 1) - n; }
            if (kb != INT_MAX && (kb & 1)) { minS = (uint32_t)sj * P.g; lastK = (kb >> 1) - n; }
        }
    }

    FwdOut f;
    f.status = status; f.minS = minS; f.lastK = lastK; f.si = si; f.n = n; f.m = m; f.top = (uint64_t)top;
    f.c_cells = c_cells; f.c_written = c_written; f.c_steps = c_steps;
    return f;
}

/* ---------------- backtrace + result, one thread per pair, then the whole group copies the
 * ops out and reduces the stats (CTA worker). */
template <bool CTA>
__device__ void finish_single(const KParams &P, const uint32_t pair, const FwdOut &f, unsigned char *smem, uint8_t *slot, const uint64_t slot_bytes)
{
    using G = Grp<CTA>;
    const int tid = G::tid(), gsz = G::size();
    int      *red   = reinterpret_cast<int *>(reinterpret_cast<uint64_t *>(reinterpret_cast<int4 *>(smem) + P.dM) + P.dM);
    uint64_t *bslot = reinterpret_cast<uint64_t *>(red + 128);
    RowHdr   *hdrs  = reinterpret_cast<RowHdr *>(slot);
    uint32_t *cells = reinterpret_cast<uint32_t *>(slot);
    const uint64_t slot_words = slot_bytes >> 2, top = f.top;
    const int si = f.si;
    int status = f.status;

    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;

    const uint64_t scratch_w = (((uint64_t)(si + 1) * sizeof(RowHdr) + 7) / 8) * 2;   /* word index, 8-byte aligned */
    uint64_t *scratch = reinterpret_cast<uint64_t *>(cells + scratch_w);
    uint32_t n_ops = 0;
    if (P.single_worker && tid == 0) P.ctr->dump_rows = (unsigned long long)(si + 1);
    if (status == ST_OK) {
        G::sync();
        if (tid == 0) {
            ArenaView A; A.hdr = hdrs; A.cells = cells; A.si_last = si; A.aos = !CTA;
            OpSink sink; sink.buf = scratch; sink.cap = (uint32_t)min((uint64_t)0x7fffffff, (top - scratch_w) / 2);
            sink.n = 0; sink.cur_op = 0; sink.cur_n = 0; sink.overflow = false; sink.stride = 1;
            back_trace(A, P, f.n, f.m, f.minS, f.lastK, res, sink);
            res.n_ops = sink.n;
            if (sink.overflow) res.status = ST_ARENA;
        }
        n_ops = G::bcast0(res.n_ops, reinterpret_cast<uint32_t *>(bslot));
        status = G::bcast0((int)res.status, reinterpret_cast<int *>(bslot) + 1);
    }
    if (status == ST_OK) {
        /* process() (wfa_cigar.go:136-214): reverse (merge already done), stats between first and last M */
        unsigned long long base = 0;
        if (tid == 0) base = atomicAdd(&P.ctr->ops_cursor, (unsigned long long)n_ops);
        base = G::bcast0(base, reinterpret_cast<unsigned long long *>(bslot));
        if (P.ops_pool != nullptr && base + n_ops > P.ops_cap) status = ST_OPS;
        else {
            int begin = INT_MAX, end = INT_MIN, d2 = 0;
            for (uint32_t i = tid; i < n_ops; i += gsz) {
                const uint64_t op = scratch[n_ops - 1 - i];
                if (P.ops_pool) P.ops_pool[base + i] = op;
                if ((op >> 32) == 'M') { begin = min(begin, (int)i); end = max(end, (int)i); }
            }
            G::reduce3(begin, end, d2, red);
            if (begin == INT_MAX) { begin = 0; end = 0; }          /* no M: begin = end = 0 (:170-186) */
            unsigned alen = 0, matches = 0, gaps = 0, regions = 0;
            for (int i = begin + tid; i <= end; i += gsz) {
                const uint64_t op = scratch[n_ops - 1 - i];
                const unsigned cnt = (unsigned)(op & 0xffffffffu), o = (unsigned)(op >> 32);
                alen += cnt;
                if (o == 'M') matches += cnt;
                else if (o == 'I' || o == 'D') { gaps += cnt; regions++; }
            }
            for (int d = 16; d > 0; d >>= 1) {
                alen += __shfl_xor_sync(0xffffffffu, alen, d); matches += __shfl_xor_sync(0xffffffffu, matches, d);
                gaps += __shfl_xor_sync(0xffffffffu, gaps, d); regions += __shfl_xor_sync(0xffffffffu, regions, d);
            }
            if (CTA) {
                const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
                unsigned *ured = reinterpret_cast<unsigned *>(red);
                __syncthreads();
                if (lane == 0) { ured[wid] = alen; ured[32 + wid] = matches; ured[64 + wid] = gaps; ured[96 + wid] = regions; }
                __syncthreads();
                alen = matches = gaps = regions = 0;
                for (int w = 0; w < nw; w++) { alen += ured[w]; matches += ured[32 + w]; gaps += ured[64 + w]; regions += ured[96 + w]; }
            }
            res.align_len = alen; res.matches = matches; res.gaps = gaps; res.gap_regions = regions;
            if (tid == 0) P.ops_where[pair] = base;
        }
    }
    if (tid == 0) {
        res.status = (uint8_t)status;
        if (status != ST_OK) {
            const unsigned long long r = atomicAdd(P.retry_ctr, 1ull);
            P.retry[r] = (uint64_t)status << 32 | pair;
        } else {
            atomicAdd(&P.ctr->cells, f.c_cells); atomicAdd(&P.ctr->cells_written, f.c_written);
            atomicAdd(&P.ctr->steps, f.c_steps); atomicAdd(&P.ctr->ops, (unsigned long long)n_ops);
            atomicMax(&P.ctr->arena_used_max, (unsigned long long)((slot_words - top + scratch_w) * 4 + 8ull * n_ops));
        }
        /* score/begin/end live in thread 0's res; stats were reduced to every thread */
        P.results[pair] = res;
    }
    G::sync();
}

/* ---------------- result of up to 32 pairs (lane j owns pair j): one ops-pool reservation per
 * group, process() (wfa_cigar.go:136-214) in one pass per lane, result records and work
 * counters.  `scratch` holds the lane's reversed, run-merged ops, `stride` words apart. */
/* Work counters of a worker that handles many groups: added to the launch's Counters once, when
 * the worker is done, instead of with same-address atomics per group. */
struct WorkAcc { unsigned long long cells, written, steps, ops, arena_max; };

struct ScratchOps {                /* j-th op as the backtrace produced it (reversed order) */
    const uint64_t *scratch; uint32_t stride;
    __device__ __forceinline__ uint64_t operator()(uint32_t j) const { return scratch[(size_t)j * stride]; }
};

template <class GetOp>
__device__ __forceinline__ void group_emit(const KParams &P, const bool have, const uint32_t pair, int status, Result &res,
                                           uint32_t n_ops, const GetOp get_op,
                                           const unsigned long long arena_used, const unsigned long long c_cells,
                                           const unsigned long long c_written, const unsigned long long c_steps, WorkAcc *acc = nullptr)
{
    const int lane = threadIdx.x & 31;
    /* one pool reservation per group: exclusive scan of n_ops over the lanes */
    uint32_t incl = n_ops;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    const uint32_t group_total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && group_total) base = atomicAdd(&P.ctr->ops_cursor, (unsigned long long)group_total);
    base = __shfl_sync(0xffffffffu, base, 0) + (incl - n_ops);
    if (status == ST_OK) {
        if (P.ops_pool != nullptr && base + n_ops > P.ops_cap) status = ST_OPS;
        else {
            /* process() (wfa_cigar.go:136-214) in one pass: copy reversed, and take the stats
             * between the first and the last M as a difference of running sums */
            unsigned alen = 0, matches = 0, gaps = 0, regions = 0;
            unsigned a0 = 0, m0 = 0, g0 = 0, r0 = 0, a1 = 0, m1 = 0, g1 = 0, r1 = 0;
            bool seenM = false;
            for (uint32_t i = 0; i < n_ops; i++) {
                const uint64_t op = get_op(n_ops - 1 - i);
                if (P.ops_pool) P.ops_pool[base + i] = op;
                const unsigned cnt = (unsigned)(op & 0xffffffffu), o = (unsigned)(op >> 32);
                if (o == 'M' && !seenM) { seenM = true; a0 = alen; m0 = matches; g0 = gaps; r0 = regions; }
                alen += cnt;
                if (o == 'M') matches += cnt;
                else if (o == 'I' || o == 'D') { gaps += cnt; regions++; }
                if (o == 'M') { a1 = alen; m1 = matches; g1 = gaps; r1 = regions; }
            }
            if (seenM) { res.align_len = a1 - a0; res.matches = m1 - m0; res.gaps = g1 - g0; res.gap_regions = r1 - r0; }
            else if (n_ops) {                                      /* no M: begin = end = 0 (:170-186) */
                const uint64_t op = get_op(n_ops - 1);
                const unsigned cnt = (unsigned)(op & 0xffffffffu), o = (unsigned)(op >> 32);
                res.align_len = cnt; res.matches = 0;
                res.gaps = (o == 'I' || o == 'D') ? cnt : 0; res.gap_regions = (o == 'I' || o == 'D') ? 1 : 0;
            }
            res.n_ops = n_ops;
            P.ops_where[pair] = base;
        }
    }
    if (have) {
        res.status = (uint8_t)status;
        if (status != ST_OK) {
            const unsigned long long r = atomicAdd(P.retry_ctr, 1ull);
            P.retry[r] = (uint64_t)status << 32 | pair;
        }
        P.results[pair] = res;
    }
    /* work counters: one set of atomics per group (a million same-address atomics per launch
     * serialise in L2), or none at all when the worker accumulates them */
    unsigned long long c0 = (have && status == ST_OK) ? c_cells : 0, c1 = (have && status == ST_OK) ? c_written : 0;
    unsigned long long c2 = (have && status == ST_OK) ? c_steps : 0, c3 = (have && status == ST_OK) ? n_ops : 0;
    unsigned long long c4 = (have && status == ST_OK) ? arena_used : 0;
    for (int d = 16; d > 0; d >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, d); c1 += __shfl_xor_sync(0xffffffffu, c1, d);
        c2 += __shfl_xor_sync(0xffffffffu, c2, d); c3 += __shfl_xor_sync(0xffffffffu, c3, d);
        c4 = max(c4, __shfl_xor_sync(0xffffffffu, c4, d));
    }
    if (acc) { acc->cells += c0; acc->written += c1; acc->steps += c2; acc->ops += c3; acc->arena_max = max(acc->arena_max, c4); }
    else if (lane == 0) {
        atomicAdd(&P.ctr->cells, c0); atomicAdd(&P.ctr->cells_written, c1);
        atomicAdd(&P.ctr->steps, c2); atomicAdd(&P.ctr->ops, c3);
        if (c4) atomicMax(&P.ctr->arena_used_max, c4);
    }
    __syncwarp();
}

/* ---------------- WARP worker: the forward passes of up to 32 pairs are followed by their
 * backtraces run lane-parallel (lane j owns pair j and its sub-slot): the pointer-chasing
 * backtrace costs one warp instruction stream for the whole group instead of one per pair. */
__device__ __noinline__ void finish_group(const KParams &P, const bool have, const uint32_t pair, const FwdOut &f, uint8_t *slot, const uint64_t slot_bytes)
{
    RowHdr   *hdrs  = reinterpret_cast<RowHdr *>(slot);
    uint32_t *cells = reinterpret_cast<uint32_t *>(slot);
    const uint64_t slot_words = slot_bytes >> 2, top = f.top;
    const uint64_t scratch_w = (((uint64_t)(f.si + 1) * sizeof(RowHdr) + 7) / 8) * 2;
    uint64_t *scratch = reinterpret_cast<uint64_t *>(cells + scratch_w);
    int status = have ? f.status : ST_PENDING;
    if (P.single_worker && have) P.ctr->dump_rows = (unsigned long long)(f.si + 1);

    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;
    uint32_t n_ops = 0;
    __syncwarp();
    if (status == ST_OK) {
        ArenaView A; A.hdr = hdrs; A.cells = cells; A.si_last = f.si; A.aos = true;
        OpSink sink; sink.buf = scratch; sink.cap = (uint32_t)min((uint64_t)0x7fffffff, (top - scratch_w) / 2);
        sink.n = 0; sink.cur_op = 0; sink.cur_n = 0; sink.overflow = false; sink.stride = 1;
        back_trace(A, P, f.n, f.m, f.minS, f.lastK, res, sink);
        n_ops = sink.n;
        if (sink.overflow) { status = ST_ARENA; n_ops = 0; }
    }
    __syncwarp();
    group_emit(P, have, pair, status, res, n_ops, ScratchOps{scratch, 1u},
               (unsigned long long)((slot_words - top + scratch_w) * 4 + 8ull * n_ops), f.c_cells, f.c_written, f.c_steps);
}

/* ------------------------------------------------------------------ kernels */
/* One kernel per (symbol width, worker shape): the 2-bit kernels hand pairs with a
 * non-ACGT byte back to the host (ST_NEED8), which re-queues them on the 8-bit
 * kernel -- keeps each kernel's code (and I-cache footprint) to one instantiation. */
#ifndef WFA_WARP_MINB
#define WFA_WARP_MINB 7
#endif
#ifndef WFA_CTA_THREADS
#define WFA_CTA_THREADS 768
#endif
template <int BITS, bool CTA>
__global__ void __launch_bounds__(CTA ? WFA_CTA_THREADS : 128, CTA ? 1 : WFA_WARP_MINB)
align_kernel(const KParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wpb = CTA ? 1 : (int)(blockDim.x >> 5);
    const int wib = CTA ? 0 : (int)(threadIdx.x >> 5);
    const size_t wbytes = worker_smem_bytes<CTA>(P.dM, P.dE, P.ring_cap, P.seq_cap);
    unsigned char *smem = smem_raw + (size_t)wib * wbytes;
    const uint64_t worker = (uint64_t)blockIdx.x * wpb + wib;
    uint8_t *slot = P.arena + worker * P.slot_bytes;
    __shared__ uint32_t next_item[4];
    if (!CTA && P.single_worker && worker != 0) return;

    if (CTA) {
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) next_item[0] = (uint32_t)atomicAdd(&P.ctr->work_next, 1ull);
            __syncthreads();
            const uint32_t item = next_item[0];
            if (item >= P.n_work) break;
            const uint32_t pair = P.work ? P.work[item] : item;
            if (BITS == 2 && (P.pflags[pair] & 1)) {
                if (threadIdx.x == 0) {
                    const unsigned long long r = atomicAdd(P.retry_ctr, 1ull);
                    P.retry[r] = (uint64_t)ST_NEED8 << 32 | pair;
                }
                continue;
            }
            const FwdOut f = forward_pair<BITS, CTA>(P, pair, smem, slot, P.slot_bytes);
            finish_single<CTA>(P, pair, f, smem, slot, P.slot_bytes);
        }
    } else {
        /* groups of P.group pairs: forward passes one after the other (sub-slot j of the
         * warp's slot), then the backtraces of the whole group lane-parallel */
        const int lane = threadIdx.x & 31;
        const uint32_t G = (uint32_t)P.group;
        const uint64_t sub_bytes = P.slot_bytes / G;
        const uint32_t n_work = P.handover ? (uint32_t)P.ctr->retry_n : P.n_work;
        for (;;) {
            uint32_t first = 0;
            if (lane == 0) first = (uint32_t)atomicAdd(&P.ctr->work_next, (unsigned long long)G);
            first = __shfl_sync(0xffffffffu, first, 0);
            if (first >= n_work) break;
            const uint32_t cnt = min(G, n_work - first);
            FwdOut mine; mine.status = ST_PENDING; mine.minS = 0; mine.lastK = 0; mine.si = 0; mine.n = mine.m = 0; mine.top = 0;
            mine.c_cells = mine.c_written = mine.c_steps = 0;
            bool have = false; uint32_t my_pair = 0;
            for (uint32_t j = 0; j < cnt; j++) {
                uint32_t pair;
                if (P.handover) {
                    const uint64_t e = P.handover[first + j];
                    if ((uint32_t)(e >> 32) != ST_RING) { if (lane == 0) atomicAdd(&P.ctr->handover_other, 1ull); continue; }
                    pair = (uint32_t)e;
                } else pair = P.work ? P.work[first + j] : first + j;
                if (BITS == 2 && (P.pflags[pair] & 1)) {
                    if (lane == 0) {
                        const unsigned long long r = atomicAdd(P.retry_ctr, 1ull);
                        P.retry[r] = (uint64_t)ST_NEED8 << 32 | pair;
                    }
                    continue;
                }
                FwdOut f;
                if (BITS == 2 && !CTA && ((P.pairs[pair].n + 15u) >> 4) + ((P.pairs[pair].m + 15u) >> 4) + 2u <= (uint32_t)P.seq_cap)
                    f = forward_pair<BITS, CTA, BITS == 2 && !CTA>(P, pair, smem, slot + (uint64_t)j * sub_bytes, sub_bytes);
                else f = forward_pair<BITS, CTA, false>(P, pair, smem, slot + (uint64_t)j * sub_bytes, sub_bytes);
                if (lane == (int)j) { mine = f; have = true; my_pair = pair; }
            }
            finish_group(P, have, my_pair, mine, slot + (uint64_t)lane * sub_bytes, sub_bytes);
        }
    }
}

/* 2-bit packing of every sequence of the batch + detection of non-ACGT bytes.
 * One warp per pair: lanes 0-15 stride over the 16-base words of the query, lanes 16-31 over
 * those of the target; 5 aligned 32-bit loads in (neighbouring lanes read neighbouring bytes),
 * one 32-bit word out. */
/* one 16-base word of a sequence: bytes 16 w .. 16 w + 15 of the `len` bytes at src (byte phase sh) */
__device__ __forceinline__ uint32_t pack_word(const uint32_t *__restrict__ src, uint32_t w, uint32_t len, uint32_t sh, bool &bad)
{
    uint32_t in[5];
#pragma unroll
    for (int j = 0; j < 5; j++) in[j] = __ldg(src + 4 * w + j);
    const int left = (int)len - (int)(16 * w);                     /* bases from this word on */
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t x = __funnelshift_r(in[j], in[j + 1], sh);       /* bases 16w+4j .. +3 */
        if (left < 16) {                                           /* last word: bytes past the end read as 'A' */
            const int rem = left - 4 * j;
            if (rem < 4) x = rem <= 0 ? 0x41414141u : ((x & ((1u << (8 * rem)) - 1u)) | (0x41414141u << (8 * rem)));
        }
        const uint32_t c = (x >> 1) & 0x03030303u;
        /* valid iff every byte equals the letter its code maps back to (A,C,T,G) */
        const uint32_t sel = (c & 3u) | ((c >> 4) & 0x30u) | ((c >> 8) & 0x300u) | ((c >> 12) & 0x3000u);
        bad |= __byte_perm(0x47544341u, 0u, sel) != x;
        const uint32_t p = (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xffu;
        o |= p << (8 * j);
    }
    return o;
}

__global__ void __launch_bounds__(256)
pack_kernel(const PairDesc *__restrict__ pairs, uint32_t n_pairs, const uint32_t *__restrict__ raw,
            uint32_t *__restrict__ packed, uint8_t *__restrict__ pflags)
{
    const uint32_t lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t pr = warp0; pr < n_pairs; pr += nwarps) {
        const PairDesc pd = pairs[pr];
        const uint64_t boff = half ? pd.t_byte : pd.q_byte;
        const uint32_t len = half ? pd.m : pd.n;
        uint32_t *out = packed + (half ? pd.t_word : pd.q_word);
        const uint32_t nwords = (len + 15) >> 4;
        const uint32_t *src = raw + (boff >> 2);
        const uint32_t sh = (uint32_t)(boff & 3) * 8;
        bool bad = false;
        for (uint32_t w = l16; w < nwords; w += 16) out[w] = pack_word(src, w, len, sh, bad);
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(reinterpret_cast<unsigned int *>(pflags) + (pr >> 2), 1u << ((pr & 3) * 8));
    }
}

/* The same for batches of short reads (no sequence longer than 160 bases = 10 words): three
 * pairs per warp, five lanes per sequence, two words per lane -- 30 of 32 lanes busy for 150-base
 * reads instead of 20 (the kernel is bound by instruction issue, not by bytes in flight). */
__global__ void __launch_bounds__(256)
pack_short_kernel(const PairDesc *__restrict__ pairs, uint32_t n_pairs, const uint32_t *__restrict__ raw,
                  uint32_t *__restrict__ packed, uint8_t *__restrict__ pflags)
{
    const uint32_t lane = threadIdx.x & 31, slot = lane / 5, j5 = lane % 5;     /* slot 0..5 = (pair, sequence), 6 = idle lanes 30, 31 */
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t pr0 = warp0 * 3; pr0 < n_pairs; pr0 += nwarps * 3) {
        const uint64_t pr = pr0 + (slot >> 1);
        bool bad = false;
        if (slot < 6 && pr < n_pairs) {
            const PairDesc pd = pairs[pr];
            const uint32_t half = slot & 1;
            const uint64_t boff = half ? pd.t_byte : pd.q_byte;
            const uint32_t len = half ? pd.m : pd.n;
            uint32_t *out = packed + (half ? pd.t_word : pd.q_word);
            const uint32_t nwords = (len + 15) >> 4;
            const uint32_t *src = raw + (boff >> 2);
            const uint32_t sh = (uint32_t)(boff & 3) * 8;
            if (j5 < nwords) out[j5] = pack_word(src, j5, len, sh, bad);
            if (j5 + 5 < nwords) out[j5 + 5] = pack_word(src, j5 + 5, len, sh, bad);
        }
        const unsigned votes = __ballot_sync(0xffffffffu, bad);
        if (slot < 6 && (slot & 1) == 0 && j5 == 0 && pr < n_pairs && (votes & (0x3ffu << (10 * (slot >> 1)))))
            atomicOr(reinterpret_cast<unsigned int *>(pflags) + (pr >> 2), 1u << ((pr & 3) * 8));
    }
}

} /* namespace wfak */
