/*
 * wfa_wide.cuh -- WIDE worker of libwfacuda.so (sm_100a): one thread-block CLUSTER per pair, for
 * alignments without heuristic whose wavefronts are thousands of diagonals wide (config 4:
 * semi-global, 10 kbp reads in 12 kbp windows: every row spans all n + m - 1 = 22 k diagonals),
 * under penalties of the default shape x : o+e : e = 2 : 4 : 1.
 *
 * Why (profiles/r1_cfg4_cfg5.md, r2_launches.md): the CTA worker of wfa_kernels.cuh re-reads its
 * source rows from the HBM arena through L2 (live rows 195 MB > L2: 62 % of those reads miss,
 * DRAM traffic 1.8x the algorithmic bytes), computes and stores 12-byte {M, I, D} raw words with
 * provenance codes for every cell, and makes separate passes over the row for Lo/Hi, the end test
 * and the start-cell search: 253 thread instructions per cell, 0.13 of the HBM roofline.  Here
 *   - the live rows (M of s-1 .. s-4 and the row being written, I and D of s-1 and s) stay on
 *     chip as 16-bit offsets, column = diagonal + n - 1, cut into one contiguous segment per CTA
 *     of the cluster: 9 rows x 22 k diagonals x 2 B = 396 KB = the shared memory of two SMs.  The
 *     only cells a CTA needs from its neighbours are the two next to its segment; their owners
 *     push them into the neighbour's halo columns through distributed shared memory
 *     (st.shared::cluster), and one cluster barrier per score makes them visible;
 *   - the same barrier carries the row's reductions (first / last present diagonal, end test,
 *     start-cell search): every CTA posts its partial results into every CTA's mailbox before
 *     arriving, so all CTAs take the same decision without a second round -- and, rows being computed
 *     over ranges that depend on the score alone, that decision is only needed one row later: the
 *     barrier of a row overlaps the first cells of the next;
 *   - a thread handles two neighbouring diagonals per step (one 32-bit shared-memory word per row),
 *     only offsets are computed (codes re-derived by the backtrace, as in the LANE / SLIM classes) --
 *     for both cells at once on 16-bit halves (VIMNMX.U16x2 / VIADD.16x2) wherever no source reaches
 *     a bound; both sequences are read through shared-memory windows (one LDS.64 + one funnel shift
 *     per 16-base compare), and the arena gets ONE 64-bit word per cell (M | I << 16 | D << 32):
 *     8 instead of 12 bytes per cell, written once with 128-bit stores, never read by the forward pass;
 *   - Lo/Hi are taken on the way; the end test (wfa.go:235-239) and the semi-global start-cell test
 *     (wfa.go:270-375, early-stop form) are paid by the few cells that can pass them
 *     -- those that have reached the end of a sequence -- through per-row flags in shared memory.
 * The backtraces run afterwards in their own kernel (wide_finish_kernel: lane-parallel over the
 * pairs of the launch; semi-global slots are addressed arithmetically, WideSemiView), so no cluster
 * idles while one thread chases pointers.  Step-by-step measurements: profiles/r2_wide_cfg4.md.
 *
 * Semantics follow the reference at /root/reference (cited as wfa.go:LINE); the recurrences are
 * next_off3 / next_off of wfa_lane.cuh.
 */
#pragma once
#include "wfa_slim.cuh"

namespace wfak {

constexpr uint32_t WIDE_MAX_M = 65534;            /* offsets up to m + 1 must fit 16 bits */
constexpr uint32_t WIDE_HEAD_BYTES = 1280;        /* mailboxes 512 + reduction scratch 640 + work item 16 + the record keeper's 56 bytes + row flags 24, padded */
constexpr int WIDE_MAX_CLUSTER = 8;

/* shared memory of one CTA: head, 9 ring rows of seg + 4 16-bit columns (two halo columns on either
 * side keep the rows word aligned), the two sequence windows (8 bytes per 16 bases + 1 entry each) */
__host__ __device__ inline size_t wide_smem_bytes(uint32_t seg, uint32_t seq_entries)
{
    return WIDE_HEAD_BYTES + 9 * (size_t)(seg + 4) * 2 + 8 + (size_t)seq_entries * 8;      /* (seg / 2 + 2) x (5 x 4 + 2 x 8) bytes of rings */
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
/* address of the same shared-memory location in CTA `rank` of the cluster */
__device__ __forceinline__ uint32_t cluster_map(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(caddr), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_u16(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u16 [%0], %1;" :: "r"(caddr), "h"((uint16_t)v) : "memory"); }
/* all threads of all CTAs of the cluster; orders the distributed-shared-memory stores before it */
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* The per-row barrier: only a handful of threads have written to other CTAs (halo cells, mailboxes); they
 * fence their stores themselves (cluster_fence) and nobody else pays a release fence that would wait for
 * the thread's arena stores to drain. */
__device__ __forceinline__ void cluster_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

/* a whole 2-bit sequence in shared memory: entry j holds words j and j + 1 */
struct SeqAll {
    uint32_t sa;
    __device__ __forceinline__ uint32_t chunk(uint32_t pos) const
    {
        uint32_t a, b;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sa + ((pos >> 4) << 3)));
        return __funnelshift_r(a, b, pos * 2u);              /* the shift count wraps at 32: (pos % 16) * 2 */
    }
};

/* Forward pass of one pair by the whole cluster: wfa.go:228-251 with next + extend fused per cell.
 * Every thread of every CTA takes the same decisions (the reductions are combined identically everywhere);
 * what only the result needs is kept by thread 0 of CTA 0, the record keeper, whose return value is complete.
 *
 * Rows and ranges.  Row si (score si * g) is computed over a range that depends on si alone: all
 * n + m - 1 diagonals for semi-global alignment (the init cells span them, wfa.go:160-183), [-si, si]
 * (clamped) for global alignment -- a superset of the reference's loop range (wfa.go:557-563: the hull
 * of the source rows' ranges +- 1 grows by at most one diagonal per row), and a cell outside the
 * reference's range has no source and comes out absent.  So the loop never waits for a reduction:
 * what the reductions deliver (M WaveFront.Lo/Hi for the header and the work counter, the end test,
 * the start-cell test) is consumed ONE ROW LATER, in the middle of the next row, which is what lets the
 * cluster barrier of row si overlap the first cells of row si + 1.  The row computed past the final
 * one is discarded (its arena cells lie beyond `top`, nobody reads them). */
template <bool SEMI>
__device__ __forceinline__ FwdOut forward_wide(const KParams &P, const uint32_t pair, const uint32_t sbase, uint8_t *slot, const uint64_t slot_bytes,
                                               const uint32_t rank, const uint32_t C)
{
    typedef uint64_t CellT;                                    /* SlimCell<3>: M | I << 16 | D << 32 */
    constexpr uint32_t HDR_CELLS = sizeof(SlimHdr) / sizeof(CellT);
    const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, wid = tid >> 5, nw = T >> 5;
    const uint32_t SEG = (uint32_t)P.wide_seg, HALF = SEG >> 1;
    const uint32_t RSM = (HALF + 2u) * 4u, RSE = (HALF + 2u) * 8u;           /* row strides: M ring (one word per column pair), I|D ring ({I, D} words per column pair) */
    const uint32_t sMail = sbase, sRed = sbase + 512u;
    const uint32_t sRingE = sbase + WIDE_HEAD_BYTES;                          /* 2 rows of {I pair, D pair} */
    const uint32_t sRingM = sRingE + 2u * RSE;                                /* 5 rows */
    const uint32_t sSeq = (sRingM + 5u * RSM + 7u) & ~7u;

    const PairDesc pd = P.pairs[pair];
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const uint32_t W = (uint32_t)(n + m - 1);                  /* columns: j = k + n - 1 in [0, W) */
    const uint32_t qent = (pd.n + 15u) >> 4, tent = (pd.m + 15u) >> 4;

    FwdOut f;
    f.status = ST_OK; f.minS = 0; f.lastK = Ak; f.si = 0; f.n = n; f.m = m; f.top = 0;
    f.c_cells = f.c_written = f.c_steps = 0; f.first_eq = false;
    if (pd.m > WIDE_MAX_M || W > C * SEG || qent + tent + 2u > P.wide_seq_cap) { f.status = ST_RING; return f; }

    /* ---- per-pair set-up: rings zeroed (absent everywhere), the sequences in their windows */
    for (uint32_t a = tid * 4u; a < 2u * RSE + 5u * RSM; a += T * 4u) sts_u32(sRingE + a, 0u);
    SeqAll Q, Tq;
    Q.sa = sSeq; Tq.sa = sSeq + (qent + 1u) * 8u;
    {
        const uint32_t *gq = P.packed + pd.q_word, *gt = P.packed + pd.t_word;
        for (uint32_t j = tid; j < qent; j += T) {
            const uint32_t a = __ldg(gq + j), b = j + 1u < qent ? __ldg(gq + j + 1) : 0u;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(Q.sa + j * 8u), "r"(a), "r"(b) : "memory");
        }
        for (uint32_t j = tid; j < tent; j += T) {
            const uint32_t a = __ldg(gt + j), b = j + 1u < tent ? __ldg(gt + j + 1) : 0u;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(Tq.sa + j * 8u), "r"(a), "r"(b) : "memory");
        }
    }
    /* per row parity: {scan (a) key, scan (b) key, end test} -- written by the few cells that reach the end of a sequence */
    const uint32_t sFlags = sbase + 1224u;
    if (tid == 0) for (uint32_t q = 0; q < 2; q++) { sts_u32(sFlags + q * 12u, (uint32_t)INT_MIN); sts_u32(sFlags + q * 12u + 4u, (uint32_t)INT_MAX); sts_u32(sFlags + q * 12u + 8u, 0u); }
    keep(Q.sa); keep(Tq.sa);                                   /* (opaque: not to be re-derived from the kernel parameters inside the cell loop) */
    cluster_sync_all();                                        /* nobody pushes a halo cell into a ring that is still being cleared */

    /* slot: headers (SlimHdr, index s/g) grow up from its start, rows grow down from its end */
    CellT   *cells = reinterpret_cast<CellT *>(slot);          /* rows grow down from the end */
    const uint32_t slot_cells = (uint32_t)min(slot_bytes / sizeof(CellT), (uint64_t)0xfffffff0u) & ~1u;
    uint32_t top = slot_cells;

    const int nm1 = n - 1;
    uint32_t base_j = rank * SEG;                              /* first column of this CTA's segment */
    keep(base_j);
    const int ilo = SEMI ? -nm1 : 0, ihi = SEMI ? m - 1 : 0;   /* init cells, wfa.go:155-183 */
    const bool edgeL = rank > 0, edgeR = rank + 1u < C;        /* the segment has a neighbour on that side: its first / last column pair needs halo cells */

    /* What only the record keeper (thread 0 of CTA 0) needs -- work counters, the last existing row, where the
     * backtrace starts, the previous row's place in the slot -- lives in its shared memory, not in 2 048 threads'
     * registers (the cell loop runs at the 64-register limit; what spills is re-read after every cluster barrier,
     * which invalidates L1). */
    struct Keep { unsigned long long c_cells, c_written, c_steps; uint32_t top_final, minS; int si_final, lastK; uint32_t p_off, p_ja, p_aw; };
    Keep *K = reinterpret_cast<Keep *>(__cvta_shared_to_generic((size_t)(sbase + 1168u)));
    const bool keeper = rank == 0 && tid == 0;
    if (keeper) { K->c_cells = K->c_written = K->c_steps = 0; K->top_final = slot_cells; K->minS = 0; K->si_final = 0; K->lastK = Ak; K->p_off = K->p_ja = K->p_aw = 0; }
    int status = ST_OK, si = -1;
    uint32_t slotM = 4;                                        /* ring slot of row si (si mod 5), advanced incrementally */
    /* the previous row, whose reductions are still on their way */
    bool pending = false;

    for (;;) {
        si++;
        slotM = slotM == 4u ? 0u : slotM + 1u;
        const bool has_init = si == 0 || si == SLIM_XG;
        const int lo = SEMI ? -nm1 : max(-si, -nm1), hi = SEMI ? m - 1 : min(si, m - 1);
        const uint32_t ja = (uint32_t)(lo + nm1) & ~1u, jb = (uint32_t)(hi + nm1) | 1u;       /* even / odd: whole column pairs */
        const uint32_t aw = jb - ja + 1u;
        /* room between the headers (growing up: this row's, two spare, some slack) and the rows (growing down) */
        if (top < aw + (uint32_t)(si + 4) * HDR_CELLS + 8u) { if (pending) cluster_wait(); status = ST_ARENA; break; }
        const uint32_t off = top - aw;
        /* this CTA's share: column pairs c (columns base_j + 2c, + 1), c in [cl, ch]; the pairs next to a
         * neighbouring segment wait for the halo cells (after the barrier of the previous row) */
        const int gl = (int)(ja >> 1) - (int)(base_j >> 1), gh = (int)(jb >> 1) - (int)(base_j >> 1);
        const int cl = max(gl, 0), ch = min(gh, (int)HALF - 1);
        const bool doL = edgeL && cl == 0 && cl <= ch, doR = edgeR && ch == (int)HALF - 1 && cl <= ch;      /* (a segment has at least 32 column pairs) */
        const int c_first = cl + (doL ? 1 : 0), c_last = ch - (doR ? 1 : 0);
        /* source and destination rows */
        uint32_t s4 = slotM + 1u; s4 = s4 >= 5u ? s4 - 5u : s4;                 /* row si-4 = slot (si+1) mod 5 */
        uint32_t s2 = slotM + 3u; s2 = s2 >= 5u ? s2 - 5u : s2;                 /* row si-2 */
        const uint32_t pe = (uint32_t)(si & 1);
        uint32_t bM4 = sRingM + s4 * RSM, bM2 = sRingM + s2 * RSM, bMc = sRingM + slotM * RSM;
        uint32_t bE1 = sRingE + (pe ^ 1u) * RSE, bEc = sRingE + pe * RSE;
        keep(bM4); keep(bM2); keep(bMc); keep(bE1); keep(bEc);                 /* (not to be re-derived inside the cell loop) */
        uint4 *grow = reinterpret_cast<uint4 *>(cells + off - ja + base_j);     /* column pair c of this CTA at grow[c] */
        keep_ptr(grow);
        int pmin = INT_MAX, pmax = INT_MIN;                                     /* first / last column PAIR with a present cell */
        const uint32_t sF = sFlags + (uint32_t)(si & 1) * 12u;

        /* one column pair: next (wfa.go:572-699) + extend (wfa.go:394-455) of its two cells, stores */
        auto cellpair = [&](const int c, auto initc) {
            constexpr bool INIT = decltype(initc)::value;
            const uint32_t wa = 4u * (uint32_t)c + 4u;
            const uint32_t pM4 = bM4 + wa, pE1 = bE1 + 2u * wa;
            const uint32_t a = lds32<-4>(pM4), b = lds32<0>(pM4), d = lds32<4>(pM4);
            const uint32_t ia = lds32<-8>(pE1), dd = lds32<12>(pE1);
            uint32_t ib, db;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ib), "=r"(db) : "r"(pE1));
            const uint32_t xm = lds32<0>(bM2 + wa);
            const uint32_t j0 = base_j + 2u * (uint32_t)c;                      /* columns j0, j0 + 1; n + k = j + 1 */
            const bool act1 = j0 + 1u < W;                                      /* the padding column past the last diagonal stays absent */
            const uint32_t um = (uint32_t)m;
            /* next (wfa.go:572-699) of both cells.  The sources as pairs of 16-bit offsets (low half: column j0):
             * when none of them reaches a bound -- every cell away from the ends of the sequences -- all validity
             * tests pass and the recurrence is I = max + 1, D = max, M = max(M[s-x] + 1, I, D) on present sources,
             * which sm_100a does on both halves at once (VIMNMX.U16x2, VIADD.16x2); otherwise cell by cell. */
            Cell3O c0, c1;
            uint32_t Iw, Dw;                                                    /* I and D of both cells as they go to the ring */
            {
                const uint32_t L2 = __byte_perm(a, b, 0x5432), I2 = __byte_perm(ia, ib, 0x5432);      /* M[s-o-e][k-1], I[s-e][k-1] */
                const uint32_t R2 = __byte_perm(b, d, 0x5432), D2 = __byte_perm(db, dd, 0x5432);      /* M[s-o-e][k+1], D[s-e][k+1] */
                const uint32_t mi = __vmaxu2(L2, I2), md = __vmaxu2(R2, D2);
                const uint32_t all = __vmaxu2(__vmaxu2(mi, md), xm);
                if (max(all & 0xffffu, all >> 16) <= min(um, j0 + 1u) && act1) {
                    const uint32_t Iw2 = __vadd2(mi, __vminu2(mi, 0x00010001u));                        /* + 1 where present */
                    const uint32_t Ew2 = __vadd2(xm, __vminu2(xm, 0x00010001u));
                    const uint32_t Mw2 = __vmaxu2(__vmaxu2(Ew2, Iw2), md);
                    c0.M = Mw2 & 0xffffu; c1.M = Mw2 >> 16; Iw = Iw2; Dw = md;
                } else {
                    c0 = next_off3(a >> 16, ia >> 16, b >> 16, db >> 16, xm & 0xffffu, um, j0 + 1u);
                    c1 = next_off3(b & 0xffffu, ib & 0xffffu, d & 0xffffu, dd & 0xffffu, xm >> 16, act1 ? um : 0u, act1 ? j0 + 2u : 0u);
                    Iw = c0.I | c1.I << 16; Dw = c0.D | c1.D << 16;
                }
            }
            const int k0 = (int)j0 - nm1;
            if (INIT) {
                /* initComponents (wfa.go:155-183): cell k of the first row / column; next's Set wins when both write */
                auto seed = [&](Cell3O &cc, const int k, const bool act) {
                    if (cc.M == 0u && act && k >= ilo && k <= ihi) {
                        const bool eq = ((Q.chunk((uint32_t)(k < 0 ? -k : 0)) ^ Tq.chunk((uint32_t)(k > 0 ? k : 0))) & 3u) == 0u;
                        if (eq ? (si == 0) : (si == SLIM_XG)) cc.M = (uint32_t)((k > 0 ? k : 0) + 1);
                    }
                };
                seed(c0, k0, true); seed(c1, k0 + 1, act1);
            }
            /* extend: a present cell has v >= 1, so it applies iff ext = min(n - v, m - h) > 0.  The first 16
             * bases are compared whatever the cell (an absent cell or one at the end of a sequence compares
             * clamped positions and advances by 0): the common case is straight-line code */
            auto extend = [&](const uint32_t M, const uint32_t j, const int k) -> uint32_t {
                const int ext = (int)min(j + 1u, um) - (int)M;                  /* min(n + k, m) - M */
                const uint32_t v = min((uint32_t)((int)M - k), (uint32_t)nm1);
                const uint32_t xx = Q.chunk(v) ^ Tq.chunk(M);
                int l = matched_bases(xx);                            /* 16 when all 16 bases agree */
                if (l >= 16 && ext > 16) {
                    do {
                        const uint32_t x2 = Q.chunk(v + (uint32_t)l) ^ Tq.chunk(M + (uint32_t)l);
                        if (x2) { l += matched_bases(x2); break; }
                        l += 16;
                    } while (l < ext);
                }
                const uint32_t Mn = M ? M + (uint32_t)max(min(l, ext), 0) : 0u;
                if (M != 0u && l >= ext) {
                    /* The cell has reached the end of a sequence -- the only cells that can pass the end test on
                     * diagonal m - n (wfa.go:235-239: offset >= m) or count in the start-cell search (wfa.go:306-323 /
                     * :341-358); they post to the row's flags in shared memory, nobody else pays for either test. */
                    if (k == Ak && Mn >= um) sts_u32(sF + 8u, 1u);
                    if (SEMI) {
                        const int h = (int)Mn, vv = h - k;
                        int cls = 0;
                        if (vv <= 0 || vv > n || h > m) cls = 1;
                        else if ((vv == n && h >= n) || (h == m && vv >= m)) cls = 2;
                        if (cls) {
                            const int key = (k + n) * 2 + (cls == 2);
                            if (k <= Ak) asm volatile("red.shared.max.s32 [%0], %1;" :: "r"(sF), "r"(key) : "memory");
                            else asm volatile("red.shared.min.s32 [%0], %1;" :: "r"(sF + 4u), "r"(key) : "memory");
                        }
                    }
                }
                return Mn;
            };
            c0.M = extend(c0.M, j0, k0);
            c1.M = extend(c1.M, j0 + 1u, k0 + 1);
            const uint32_t Mw = c0.M | c1.M << 16;
            sts_u32(bMc + wa, Mw);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(bEc + 2u * wa), "r"(Iw), "r"(Dw) : "memory");
            /* arena: two 8-byte cells {M | I << 16, D}, one 128-bit store */
            grow[c] = make_uint4(__byte_perm(Mw, Iw, 0x5410), Dw & 0xffffu, __byte_perm(Mw, Iw, 0x7632), Dw >> 16);
            if (Mw) { pmin = min(pmin, c); pmax = max(pmax, c); }
        };
        auto run = [&](auto initc) {
            int c = c_first + (int)tid;
            if (c <= c_last) { cellpair(c, initc); c += (int)T; }
            if (pending) {
                /* ---- the previous row: its barrier has had a column pair's worth of time to complete */
                cluster_wait();
                pending = false;
                const uint32_t mbase = sMail + (uint32_t)((si - 1) & 1) * 256u;
                int g0 = INT_MAX, g1 = INT_MIN, g3 = INT_MIN, g4 = INT_MAX; uint32_t g2 = 0;
                for (uint32_t r = 0; r < C; r++) {
                    const uint32_t mb = mbase + r * 32u;
                    g0 = min(g0, (int)lds_u32(mb)); g1 = max(g1, (int)lds_u32(mb + 4u)); g2 |= lds_u32(mb + 8u);
                    if (SEMI) { g3 = max(g3, (int)lds_u32(mb + 12u)); g4 = min(g4, (int)lds_u32(mb + 16u)); }
                }
                const bool exists = g0 <= g1, endhit = g2 != 0u;
                bool hit = false; int hitK = Ak;
                if (SEMI) {
                    /* backtraceStartPosistion (wfa.go:270-375) for this one score: scan (a) runs from the end
                     * diagonal downwards, scan (b) upwards from the one above it; (b) overrides (a) */
                    if (g3 != INT_MIN && (g3 & 1)) { hit = true; hitK = (g3 >> 1) - n; }
                    if (g4 != INT_MAX && (g4 & 1)) { hit = true; hitK = (g4 >> 1) - n; }
                }
                if (keeper) {
                    /* header of the row, work counters, where the backtrace starts */
                    const int wlo = g0 - nm1, whi = g1 - nm1;
                    int4 hc = make_int4(0, 1, 0, 0);
                    K->c_written += K->p_aw;
                    if (exists) {
                        K->c_steps++; K->c_cells += (unsigned long long)(whi - wlo + 1);
                        hc = make_int4((int)K->p_ja - nm1, wlo, whi, (int)K->p_off);
                        K->top_final = K->p_off;
                    }
                    *reinterpret_cast<int4 *>(reinterpret_cast<SlimHdr *>(slot) + (si - 1)) = hc;
                    K->si_final = si - 1;
                    if (exists && (endhit || hit)) { K->minS = (uint32_t)(si - 1) * P.g; K->lastK = SEMI && hit ? hitK : Ak; }
                }
                if (exists && (endhit || (SEMI && hit))) return true;
            }
            /* the column pairs next to the neighbouring segments, now that their halo cells are here; their
             * outermost cells go on to the neighbour that reads them as k + 1 / k - 1 */
            if (wid == nw - 1u && lane < 2u && (lane == 0 ? doL : doR)) {
                /* (both in one pass: the warp with the fewest column pairs of its own takes one more) */
                cellpair(lane == 0 ? 0 : (int)HALF - 1, initc);
                if (lane == 0) {
                    const uint32_t left = rank - 1u;
                    st_cluster_u16(cluster_map(bMc + 4u * (HALF + 1u), left), lds_u16(bMc + 4u));             /* M -> column SEG of the left CTA */
                    st_cluster_u16(cluster_map(bEc + 8u * (HALF + 1u) + 4u, left), lds_u16(bEc + 8u + 4u));   /* D */
                } else {
                    const uint32_t right = rank + 1u;
                    st_cluster_u16(cluster_map(bMc + 2u, right), lds_u16(bMc + 4u * HALF + 2u));              /* M -> column -1 of the right CTA */
                    st_cluster_u16(cluster_map(bEc + 2u, right), lds_u16(bEc + 8u * HALF + 2u));              /* I */
                }
                cluster_fence();
            }
            for (; c <= c_last; c += (int)T) cellpair(c, initc);
            return false;
        };
        const bool done = has_init ? run(std::true_type{}) : run(std::false_type{});
        if (done) break;

        /* ---- this row's reductions: exact first / last present column from the thread's own pair words; the end
         * test and the start-cell keys are in the row's flags already */
        int cmin = INT_MAX, cmax = INT_MIN;
        if (pmin <= pmax) {
            const uint32_t w0 = lds_u32(bMc + 4u * (uint32_t)pmin + 4u), w1 = lds_u32(bMc + 4u * (uint32_t)pmax + 4u);
            cmin = (int)base_j + 2 * pmin + ((w0 & 0xffffu) ? 0 : 1);
            cmax = (int)base_j + 2 * pmax + ((w1 >> 16) ? 1 : 0);
        }
        cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
        if (lane == 0) { sts_u32(sRed + wid * 4u, (uint32_t)cmin); sts_u32(sRed + 128u + wid * 4u, (uint32_t)cmax); }
        __syncthreads();
        if (wid == 0) {
            int a0 = lane < nw ? (int)lds_u32(sRed + lane * 4u) : INT_MAX, a1 = lane < nw ? (int)lds_u32(sRed + 128u + lane * 4u) : INT_MIN;
            a0 = __reduce_min_sync(0xffffffffu, a0); a1 = __reduce_max_sync(0xffffffffu, a1);
            const uint32_t a3 = lds_u32(sF), a4 = lds_u32(sF + 4u), a2 = lds_u32(sF + 8u);
            __syncwarp();
            if (lane == 0) { sts_u32(sF, (uint32_t)INT_MIN); sts_u32(sF + 4u, (uint32_t)INT_MAX); sts_u32(sF + 8u, 0u); }      /* free for the row after the next */
            if (lane < C) {
                const uint32_t dst = cluster_map(sMail + (uint32_t)(si & 1) * 256u + rank * 32u, lane);
                st_cluster_u32(dst, (uint32_t)a0); st_cluster_u32(dst + 4u, (uint32_t)a1); st_cluster_u32(dst + 8u, a2);
                st_cluster_u32(dst + 12u, a3); st_cluster_u32(dst + 16u, a4);
                cluster_fence();
            }
        }
        /* arrive now, wait in the middle of the next row: halos (their writers' fences) and mailboxes are
         * in place everywhere once the barrier completes */
        cluster_arrive_relaxed();
        pending = true;
        if (keeper) { K->p_off = off; K->p_ja = ja; K->p_aw = aw; }
        top = off;
    }

    f.status = status;
    if (keeper) {
        f.minS = K->minS; f.lastK = K->lastK; f.si = K->si_final; f.top = (uint64_t)K->top_final;
        f.c_cells = K->c_cells; f.c_written = K->c_written; f.c_steps = K->c_steps;
    }
    return f;
}

/* Persistent clusters: item i of the launch is aligned into slot i of the arena (the slots are
 * read by wide_finish_kernel afterwards), its FwdOut goes to P.wide_rec[i]. */
template <bool SEMI>
__global__ void __launch_bounds__(1024, 1)
wide_kernel(const KParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t rank = cluster_ctarank(), C = cluster_nctarank();
    const uint32_t sItem = sbase + 1152u;
    for (;;) {
        if (rank == 0 && threadIdx.x == 0) {
            const uint32_t it = (uint32_t)atomicAdd(&P.ctr->work_next, 1ull);
            for (uint32_t r = 0; r < C; r++) st_cluster_u32(cluster_map(sItem, r), it);
        }
        cluster_sync_all();
        const uint32_t item = lds_u32(sItem);
        if (item >= P.n_work) break;
        const uint32_t pair = P.work ? P.work[item] : item;
        FwdOut f;
        if (P.pflags[pair] & 1) {
            f.status = ST_NEED8; f.minS = 0; f.lastK = 0; f.si = 0; f.n = f.m = 0; f.top = 0;
            f.c_cells = f.c_written = f.c_steps = 0; f.first_eq = false;
        } else f = forward_wide<SEMI>(P, pair, sbase, P.arena + (uint64_t)item * P.slot_bytes, P.slot_bytes, rank, C);
        if (rank == 0 && threadIdx.x == 0) P.wide_rec[item] = f;
        cluster_sync_all();                                     /* the item word and the rings are free again */
    }
}

/* Component.Get on a WIDE slot of a SEMI-GLOBAL pair: every row spans the same columns (all n + m - 1 diagonals,
 * padded to whole column pairs) and every computed row has its cells in the slot (absent ones as zeros), so the cell of
 * (score index, diagonal) sits at an address that is arithmetic in both -- the backtrace's dependent chain is one arena
 * read per step instead of header + cell.  Otherwise SlimView<3>. */
struct WideSemiView {
    typedef SlimCell<3> SC;
    const uint64_t *cells;
    uint32_t slot_cells, aw; int alo;
    int si_last, n, m;
    int c_si, c_k; uint64_t c_w[5];
    __device__ __forceinline__ uint64_t word(int si, int k) const
    {
        const uint32_t col = (uint32_t)(k - alo);
        if (si < 0 || si > si_last || col >= aw) return 0;
        return __ldg(cells + (slot_cells - (uint32_t)(si + 1) * aw + col));
    }
    __device__ __forceinline__ uint64_t cached_word(int si, int k) const
    {
        const int dk = k - c_k, ds = c_si - si;
        if (c_si >= 0) {
            if (dk == -1) { if (ds == SLIM_OEG) return c_w[0]; if (ds == SLIM_EG) return c_w[1]; }
            else if (dk == 1) { if (ds == SLIM_OEG) return c_w[2]; if (ds == SLIM_EG) return c_w[3]; }
            else if (dk == 0 && ds == SLIM_XG) return c_w[4];
        }
        return word(si, k);
    }
    __device__ __forceinline__ uint32_t get(int comp, int si, int k) const { return SC::get(cached_word(si, k), comp) << T_BITS; }
    __device__ __forceinline__ uint32_t get_typed(int comp, int si, int k)
    {
        const uint32_t o = SC::get(cached_word(si, k), comp);
        if (o == 0) return 0;
        const uint64_t wl = word(si - SLIM_OEG, k - 1), el = word(si - SLIM_EG, k - 1);
        const uint64_t wr = word(si - SLIM_OEG, k + 1), er = word(si - SLIM_EG, k + 1);
        const uint64_t wx = word(si - SLIM_XG, k);
        c_si = si; c_k = k; c_w[0] = wl; c_w[1] = el; c_w[2] = wr; c_w[3] = er; c_w[4] = wx;
        const CellO c = next_off(SC::get(wl, 0), SC::get(el, 1), SC::get(wr, 0), SC::get(er, 2), SC::get(wx, 0),
                                 (uint32_t)m, (uint32_t)(n + k));
        uint32_t code;
        if (comp == 1) code = T_INS_OPEN + ((c.code >> 3) & 1u);
        else if (comp == 2) code = T_DEL_OPEN + ((c.code >> 4) & 1u);
        else code = c.M ? (c.code & 7u) : (si == 0 ? T_MATCH : T_MISMATCH);          /* init cell (wfa.go:160-183) */
        return o << T_BITS | code;
    }
};

__device__ __noinline__ void finish_group_wide_semi(const KParams &P, const bool have, const uint32_t pair, const FwdOut &f, uint8_t *slot, const uint64_t slot_bytes)
{
    uint32_t *words = reinterpret_cast<uint32_t *>(slot);
    const uint64_t slot_words = slot_bytes >> 2, top_w = f.top * 2;
    const uint64_t scratch_w = (((uint64_t)(f.si + 1) * sizeof(SlimHdr) + 7) / 8) * 2;
    uint64_t *scratch = reinterpret_cast<uint64_t *>(words + scratch_w);
    int status = have ? f.status : ST_PENDING;
    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;
    uint32_t n_ops = 0;
    __syncwarp();
    if (status == ST_OK) {
        WideSemiView A; A.cells = reinterpret_cast<const uint64_t *>(slot);
        A.slot_cells = (uint32_t)min(slot_bytes / 8, (uint64_t)0xfffffff0u) & ~1u;
        A.alo = -(f.n - 1); A.aw = (((uint32_t)(f.n + f.m - 2)) | 1u) + 1u;          /* columns 0 .. (W - 1) | 1, as the forward pass laid them out */
        A.si_last = f.si; A.n = f.n; A.m = f.m; A.c_si = -1; A.c_k = 0;
        A.c_w[0] = A.c_w[1] = A.c_w[2] = A.c_w[3] = A.c_w[4] = 0;
        OpSink sink; sink.buf = scratch; sink.cap = (uint32_t)min((uint64_t)0x7fffffff, top_w > scratch_w ? (top_w - scratch_w) / 2 : 0);
        sink.n = 0; sink.cur_op = 0; sink.cur_n = 0; sink.overflow = false; sink.stride = 1;
        back_trace_inl(A, P, f.n, f.m, f.minS, f.lastK, res, sink);
        n_ops = sink.n;
        if (sink.overflow) { status = ST_ARENA; n_ops = 0; }
    }
    __syncwarp();
    group_emit(P, have, pair, status, res, n_ops, ScratchOps{scratch, 1u},
               (unsigned long long)((slot_words - top_w + scratch_w) * 4 + 8ull * n_ops), f.c_cells, f.c_written, f.c_steps);
}

/* Backtraces (wfa.go:703-983) and results of the items of a WIDE launch, lane-parallel. */
__global__ void __launch_bounds__(128)
wide_finish_kernel(const KParams P)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t item = warp * 32u + lane;
    if (warp * 32u >= P.n_work) return;
    const bool have = item < P.n_work;
    FwdOut f;
    f.status = ST_PENDING; f.minS = 0; f.lastK = 0; f.si = 0; f.n = f.m = 0; f.top = 0; f.c_cells = f.c_written = f.c_steps = 0; f.first_eq = false;
    uint32_t pair = 0;
    if (have) { f = P.wide_rec[item]; pair = P.work ? P.work[item] : item; }
    uint8_t *slot = P.arena + (uint64_t)(have ? item : 0u) * P.slot_bytes;
    if (!P.global_aln) finish_group_wide_semi(P, have, pair, f, slot, P.slot_bytes);
    else finish_group_slim<3>(P, have, pair, f, slot, P.slot_bytes);
}

} /* namespace wfak */
