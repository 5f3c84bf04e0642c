/*
 * wfa_slim.cuh -- SLIM worker of libwfacuda.so (sm_100a): one warp per pair, for global
 * alignments under penalties of the default shape x : o+e : e = 2 : 4 : 1 (4/6/2 and its
 * multiples) whose wavefronts stay below 32 MAXP diagonals (wf-adaptive reduction, or few
 * errors).  Configs 3 and 5.
 *
 * Why (profiles/r1_cfg3_warp.md): the WARP worker of wfa_kernels.cuh spends ~590 warp
 * instructions per score step of a ~30-cell row -- provenance codes for every cell, raw
 * 12-byte {M,I,D} triples in ring and arena, row headers in shared memory, mode branches,
 * per-row passes for Lo/Hi and reduce.  This worker keeps its shape (lane = diagonal k - lo,
 * ceil(width / 32) passes per row, source rows in a per-warp shared-memory ring) and removes
 * the rest:
 *   - only offsets are computed (as in the LANE class): the 3-bit provenance code is a
 *     function of the five source offsets and is re-derived by the backtrace for the cells it
 *     visits (SlimView::get_typed), not computed and packed for every cell;
 *   - the arena holds ONE word per (score, diagonal): M | I << 10 | D << 20 (targets up to
 *     1022 bases) or M | I << 21 | D << 42 in 64 bits, plus a 16-byte header per score;
 *   - the kernel is specialised on the penalty shape, on wf-adaptive and on the cell word, so
 *     ring depths and source rows are compile-time constants and the ranges of the four live
 *     rows sit in registers; passes are unrolled (MAXP) and keep the lane's cells in
 *     registers, so Lo/Hi, the end test and `reduce` never look at the row again;
 *   - both sequences are read through a 2 KB shared-memory window per warp that follows the
 *     front (one LDS.64 + one funnel shift per sequence and 16-base compare; the window tests
 *     are hoisted out of the compare loop), so `extend` of a 100 kbp pair stays out of L2.
 * A pair whose row outgrows 32 MAXP diagonals is reported as ST_RING and re-queued on a wider
 * instantiation or on the WARP worker.  Semantics follow the reference at /root/reference
 * (cited as wfa.go:LINE); the recurrences themselves are next_off3 / next_off of wfa_lane.cuh.
 */
#pragma once
#include "wfa_kernels.cuh"
#include "wfa_lane.cuh"

namespace wfak {

constexpr int SLIM_XG = 2, SLIM_OEG = 4, SLIM_EG = 1;   /* x, o+e, e in units of g */
constexpr int SLIM_WIN = 128;                            /* sequence window: 128 entries of 16 bases per sequence */
constexpr uint32_t SLIM_MAX_M10 = 1022;                  /* offsets up to m+1 must fit 10 bits ... */
constexpr uint32_t SLIM_MAX_SHORT = 16 * SLIM_WIN;       /* ... the whole sequence fits the window ... */
constexpr uint32_t SLIM_MAX_M21 = (1u << 21) - 2;        /* ... or offsets fit 21 bits */
constexpr int SLIM_NONE_LO = 1 << 30, SLIM_NONE_HI = -(1 << 30);

/* One score's row in a SLIM slot: cells of diagonals [alo, alo + aw) start at cell index `off`;
 * [lo, hi] = M WaveFront.Lo/Hi after reduce (outside: absent, wfa.go:526-537), lo > hi: no such score. */
struct SlimHdr { int32_t alo, lo, hi; uint32_t off; };

/* Size class of a launch: 0 = targets up to 1022 bases (32-bit cells), 1 = sequences up to 2048
 * bases (64-bit cells, whole sequences in the window), 2 = longer (64-bit cells, moving window) */
template <int SZ> struct SlimCell {
    static constexpr bool WIDE = SZ != 0;
    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type T;
    static constexpr int BITS = SZ == 3 ? 16 : WIDE ? 21 : 10;            /* SZ 3: the WIDE worker's cells (16-bit offsets: M | I << 16 in one half, D in the other) */
    __device__ static __forceinline__ T pack(uint32_t M, uint32_t I, uint32_t D)
    {
        if (SZ == 3) return (uint64_t)(M | I << 16) | (uint64_t)D << 32;
        if (WIDE) return (uint64_t)(M | I << 21) | (uint64_t)(I >> 11 | D << 10) << 32;
        return (T)(M | I << 10 | D << 20);
    }
    __device__ static __forceinline__ uint32_t get(T w, int comp)
    {
        return (uint32_t)(w >> (BITS * comp)) & ((1u << BITS) - 1u);
    }
};

/* Shared memory of a warp: M ring (5 rows), I|D ring (2 rows of two planes), the two sequence windows.
 * A row is stored from its own first diagonal on (column 0 = alo), so a source cell k +- 1 sits at a
 * fixed distance from the lane's base address whatever the pass; reads that fall outside a source
 * row's range are discarded and may stray up to one row before and two rows behind it -- into a
 * neighbouring row, the windows behind the rings, or the pad in front of the block's first warp. */
__host__ __device__ inline size_t slim_smem_bytes(int maxp) { return 9 * (size_t)maxp * 128 + 2 * (size_t)SLIM_WIN * 8; }   /* per warp */
__host__ __device__ inline size_t slim_smem_pad(int maxp) { return (size_t)maxp * 128 + 16; }                                /* per block, in front */

/* A sequence seen through the warp's shared-memory window: entry j of the ring holds the 2-bit
 * words j and j+1 of the sequence, for j in [wbase, wend).  A compare that would leave the
 * window is cut short (or not started) and finished from the packed pool by the caller. */
struct SeqWin {
    const uint32_t *g;
    uint32_t sa;            /* shared-window byte address of ring entry 0 */
    uint32_t wbase, wend, nwords;
    __device__ __forceinline__ void fill(uint32_t from, uint32_t to, int lane)
    {
        for (uint32_t j = from + (uint32_t)lane; j < to; j += 32) {
            const uint32_t a = __ldg(g + j), b = __ldg(g + j + 1);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sa + (j & (SLIM_WIN - 1)) * 8u), "r"(a), "r"(b) : "memory");
        }
    }
    __device__ __forceinline__ void init(const uint32_t *words, uint32_t n_sym, uint32_t ring_sa, int lane)
    {
        g = words; sa = ring_sa; nwords = (n_sym + 15u) >> 4;
        wbase = 0; wend = min(nwords, (uint32_t)SLIM_WIN);
        fill(0, wend, lane);
    }
    /* 16 bases from base `pos` on (base pos in the low bits); pos / 16 must be inside the window */
    __device__ __forceinline__ uint32_t chunk(uint32_t pos) const
    {
        uint32_t a, b;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sa + ((pos >> 1) & ((SLIM_WIN - 1) * 8u))));
        return __funnelshift_r(a, b, pos * 2u);              /* the shift count wraps at 32: (pos % 16) * 2 */
    }
    __device__ __forceinline__ uint32_t chunk_global(uint32_t pos) const
    {
        const uint32_t wi = pos >> 4;
        return __funnelshift_r(__ldg(g + wi), __ldg(g + wi + 1), pos * 2u);
    }
    /* warp-uniform: move the window on so that it ends 32..64 entries past word `w` (its oldest entries are dropped) */
    __device__ __forceinline__ void advance_to(uint32_t w, int lane)
    {
        const uint32_t to = min(nwords, (w + 64u) & ~31u);
        if (to <= wend) return;
        const uint32_t base = to > (uint32_t)SLIM_WIN ? to - (uint32_t)SLIM_WIN : 0u;
        __syncwarp();
        fill(max(wend, base), to, lane);
        wend = to; wbase = base;
        __syncwarp();
    }
};


/* f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>): a loop whose index is a constant
 * expression in the body (offsets of the shared-memory accesses become immediates) */
template <class F, int... Is>
__device__ __forceinline__ void static_for(F &&f, std::integer_sequence<int, Is...>) { (f(std::integral_constant<int, Is>{}), ...); }

/* Forward pass of one pair: wfa.go:228-251 with next + extend fused per cell. */
template <int MAXP, int SZ, bool ADAPT>
__device__ __forceinline__ FwdOut forward_slim(const KParams &P, const uint32_t pair, const uint32_t smem_sa, uint8_t *slot, const uint64_t slot_bytes)
{
    constexpr int WR = 32 * MAXP;                  /* cells per ring row */
    constexpr uint32_t ROWB = WR * 4;
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr bool LONGSEQ = SZ == 2;
    typedef SlimCell<SZ> SC;
    typedef typename SC::T CellT;
    constexpr uint32_t HDR_CELLS = sizeof(SlimHdr) / sizeof(CellT);
    int lane = threadIdx.x & 31;
    keep(lane);                                                         /* (not to be re-read from the special register in the loop) */
    const PairDesc pd = P.pairs[pair];
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const int maxdiff = P.max_dist_diff, min_wf_len = P.min_wf_len;

    FwdOut f;
    f.status = ST_OK; f.minS = 0; f.lastK = Ak; f.si = 0; f.n = n; f.m = m; f.top = 0;
    f.c_cells = f.c_written = f.c_steps = 0; f.first_eq = false;
    /* offsets must fit the cell word, sequences of the short classes the window */
    if ((uint32_t)m > (SZ == 0 ? SLIM_MAX_M10 : SLIM_MAX_M21) || (SZ < 2 && (uint32_t)max(n, m) > SLIM_MAX_SHORT)) { f.status = ST_RING; return f; }

    /* shared memory of the warp: M ring (5 rows: s-4 .. s), I|D ring (2 rows, I plane then D plane), the two sequence windows */
    uint32_t rM = smem_sa;
    keep(rM);
    const uint32_t rE = rM + 5 * ROWB;
    SeqWin Q, T;
    __syncwarp();
    Q.init(P.packed + pd.q_word, (uint32_t)n, rE + 4 * ROWB, lane);
    T.init(P.packed + pd.t_word, (uint32_t)m, rE + 4 * ROWB + (uint32_t)SLIM_WIN * 8u, lane);
    __syncwarp();

    SlimHdr *hdrs = reinterpret_cast<SlimHdr *>(slot);                  /* grows up, index s/g */
    CellT   *cells = reinterpret_cast<CellT *>(slot);                   /* rows grow down from the end */
    CellT   *cells_lane = cells + lane;
    keep_ptr(cells_lane);
    const uint32_t slot_cells = (uint32_t)min(slot_bytes / sizeof(CellT), (uint64_t)0xfffffff0u);
    uint32_t top = slot_cells;
    /* cells still free between the headers (growing up) and the rows (growing down), after the
     * headers of the first rows, one spare header and some slack */
    int room = (int)min(slot_cells, 0x7fffffffu) - (int)(3 * HDR_CELLS + 8);

    /* extend (wfa.go:394-455) of a present cell, offset h = M on diagonal k, inside the window;
     * `slow` is raised when the compare has to be finished outside the window */
    auto extend = [&](uint32_t M, int k, bool &slow) -> uint32_t {
        const int h = (int)M, v = h - k;
        /* a present cell has v >= 1, so "v > 0, v < n, h < m" is "min(n-v, m-h) > 0" */
        const int ext = min(n - v, m - h);
        if (!LONGSEQ) {
            /* whole sequences in the window: the first 16 bases are compared whatever the cell (the window
             * addresses wrap inside the ring; an absent cell or one at the end of a sequence advances by 0),
             * so the common case is straight-line code for all lanes */
            const uint32_t xx = Q.chunk((uint32_t)v) ^ T.chunk((uint32_t)h);
            int l = matched_bases(xx);                                /* 16 when all 16 bases agree */
            if (l >= 16 && ext > 16) {
                do {
                    const uint32_t x2 = Q.chunk((uint32_t)(v + l)) ^ T.chunk((uint32_t)(h + l));
                    if (x2) { l += matched_bases(x2); break; }
                    l += 16;
                } while (l < ext);
            }
            return M ? M + (uint32_t)max(min(l, ext), 0) : 0u;
        }
        /* moving window: the first 16 bases are compared before anything is known about the cell (the window
         * addresses wrap inside the ring), the tests that decide what the compare is worth follow */
        const uint32_t xx = Q.chunk((uint32_t)v) ^ T.chunk((uint32_t)h);
        if (M == 0u || ext <= 0) return M;
        if ((uint32_t)v < Q.wbase * 16u || (uint32_t)h < T.wbase * 16u) { slow = true; return M; }
        const int lim = min(ext, min((int)(Q.wend * 16u) - v, (int)(T.wend * 16u) - h));       /* > 0 only if both chunks start inside */
        if (lim <= 0) { slow = true; return M; }
        int l = matched_bases(xx);                                    /* 16 when all 16 bases agree */
        if (l >= 16 && lim > 16) {
            do {
                const uint32_t x2 = Q.chunk((uint32_t)(v + l)) ^ T.chunk((uint32_t)(h + l));
                if (x2) { l += matched_bases(x2); break; }
                l += 16;
            } while (l < lim);
        }
        if (l >= lim && lim < ext) slow = true;
        return M + (uint32_t)min(l, lim);
    };
    /* the same from the packed pool (idempotent on an already extended cell) */
    auto extend_global = [&](uint32_t M, int k) -> uint32_t {
        const int h = (int)M, v = h - k;
        const int ext = min(n - v, m - h);
        if (M == 0u || ext <= 0) return M;
        int l = 0;
        do {
            const uint32_t xx = Q.chunk_global((uint32_t)(v + l)) ^ T.chunk_global((uint32_t)(h + l));
            if (xx) { l += matched_bases(xx); break; }
            l += 16;
        } while (l < ext);
        return M + (uint32_t)min(l, ext);
    };

    /* ranges [lo, hi] (after reduce) of the rows s-1 .. s-4; absent: (NONE_LO, NONE_HI) */
    int lo1 = SLIM_NONE_LO, hi1 = SLIM_NONE_HI, lo2 = SLIM_NONE_LO, hi2 = SLIM_NONE_HI;
    int lo3 = SLIM_NONE_LO, hi3 = SLIM_NONE_HI, lo4 = SLIM_NONE_LO, hi4 = SLIM_NONE_HI;
    /* where their cell of diagonal 0 would be: ring row address - 4 * (first diagonal stored);
     * an absent row points at a ring row too, whatever is read there is discarded */
    uint32_t zM1 = rM, zM2 = rM, zM3 = rM, zM4 = rM, zE1 = rE;
    const uint32_t lane4 = (uint32_t)lane * 4u;

    uint32_t c_cells = 0, c_steps = 0;
    int status = ST_OK, si = 0;
    uint32_t minS = 0;
    bool finished = false;

    /* initComponents (wfa.go:155-158): M[0][0] = 1 if q[0] == t[0], else M[x][0] = 1.  `next` has no
     * source before that row, so the rows before it do not exist and the row itself is this one cell. */
    const bool first_eq = ((Q.chunk(0u) ^ T.chunk(0u)) & 3u) == 0u;
    const int si_first = first_eq ? 0 : SLIM_XG;
    uint32_t recM = rM + (uint32_t)(si_first % 5) * ROWB, recE = rE + (uint32_t)(si_first & 1) * 2u * ROWB;    /* ring rows of score index si */
    {
        const SlimHdr none = {0, 1, 0, 0u};
        for (int j = lane; j < si_first; j += 32) hdrs[j] = none;
        bool slow = false;
        uint32_t Mx = extend(1u, 0, slow);                              /* the same on every lane */
        if (LONGSEQ && slow) Mx = extend_global(1u, 0);
        top -= 1;
        if (lane == 0) {
            sts_u32(recM, Mx); sts_u32(recE, 0u); sts_u32(recE + ROWB, 0u);
            cells[top] = SC::pack(Mx, 0u, 0u);
            const SlimHdr h0 = {0, 0, 0, top};
            hdrs[si_first] = h0;
        }
        __syncwarp();
        si = si_first; lo1 = hi1 = 0; zM1 = recM; zE1 = recE;
        c_steps = 1; c_cells = 1;
        room -= (int)(HDR_CELLS * (uint32_t)(si_first + 1)) + 1;
        if (Ak == 0 && (int)Mx >= m) { finished = true; minS = (uint32_t)si * P.g; }      /* wfa.go:235-239 */
    }

    /* One row of NP passes of 32 diagonals (GUARD: up to NP passes, the row's width decides).  The
     * body is instantiated per pass count, so that the common narrow rows run straight-line code:
     * next + extend + stores per pass, with the lane's Lo/Hi, end test and distance-to-end taken
     * on the way, then the warp reductions. */
    int lo = 0, hi = 0, aw = 0, elo = 0, ehi = 0, width = 0;
    uint32_t off = 0;
    bool endhit = false, row_exists = false;
    auto row = [&](auto npc, auto guardc) {
        constexpr int NP = decltype(npc)::value;
        constexpr bool GUARD = decltype(guardc)::value;
        /* this lane's addresses in the source rows M[s-o-e], M[s-x], I|D[s-e] and in the destination
         * rows, for pass 0: pass p is 128 p bytes on, neighbours are +- 4 bytes */
        const int k0 = lo + lane;
        const uint32_t kb = (uint32_t)k0 * 4u;
        const uint32_t aO = zM4 + kb, aX = zM2 + kb, aE = zE1 + kb, aC = recM + lane4, aF = recE + lane4;
        const uint32_t cntO = (uint32_t)max(hi4 - lo4 + 1, 0), cntE = (uint32_t)max(hi1 - lo1 + 1, 0), cntX = (uint32_t)max(hi2 - lo2 + 1, 0);
        CellT *grow = cells_lane + off;
        uint32_t Mn[NP]; uint32_t slowmask = 0;
        uint32_t du[NP];                                 /* distance to the end of the lane's cells, 0xffffffff = not counted */
        int pmin = INT_MAX, pmax = INT_MIN; uint32_t dmin = 0xffffffffu;
        bool hit = false;
        auto pass = [&](auto pc) {
            constexpr int p = decltype(pc)::value;
            Mn[p] = 0u; du[p] = 0xffffffffu;
            if (!GUARD || p * 32 < aw) {
                const int k = k0 + 32 * p;
                /* five sources; a row is only valid inside its range (Get, wfa_wavefront.go:153-159) */
                const int rO = k - lo4, rI = k - lo1, rX = k - lo2;
                uint32_t mo_l = lds32<128 * p - 4>(aO), mo_r = lds32<128 * p + 4>(aO);
                uint32_t ie_l = lds32<128 * p - 4>(aE), de_r = lds32<128 * p + 4 + (int)ROWB>(aE), mx = lds32<128 * p>(aX);
                mo_l = (uint32_t)(rO - 1) < cntO ? mo_l : 0u; mo_r = (uint32_t)(rO + 1) < cntO ? mo_r : 0u;
                ie_l = (uint32_t)(rI - 1) < cntE ? ie_l : 0u; de_r = (uint32_t)(rI + 1) < cntE ? de_r : 0u;
                mx = (uint32_t)rX < cntX ? mx : 0u;
                /* only the row's last pass reaches beyond hi */
                const bool act = (GUARD || p == NP - 1) ? k <= hi : true;
                const uint32_t ubk = (uint32_t)(n + k);
                Cell3O c = next_off3(mo_l, ie_l, mo_r, de_r, mx, act ? (uint32_t)m : 0u, act ? ubk : 0u);
                bool slow = false;
                c.M = extend(c.M, k, slow);
                if (LONGSEQ && slow) slowmask |= 1u << p;
                Mn[p] = c.M;
                sts32<128 * p>(aC, c.M); sts32<128 * p>(aF, c.I); sts32<128 * p + (int)ROWB>(aF, c.D);
                if (act) grow[32 * p] = SC::pack(c.M, c.I, c.D);
                {
                    /* M WaveFront.Lo/Hi (first / last present cell), end test on diagonal m-n (wfa.go:235-239) */
                    const bool present = c.M != 0u;
                    pmin = present ? min(pmin, k) : pmin; pmax = present ? k : pmax;
                    hit = hit || (k == Ak && c.M >= (uint32_t)m);
                    if (ADAPT) {
                        /* distance to the end for reduce (wfa.go:474-494): counted iff present, v < n, h < m, i.e.
                         * 1 <= M < min(n + k, m) (a present cell has v >= 1); then max(m - h, n - v) = max(m, n + k) - M */
                        const bool counted = c.M - 1u < min(ubk, (uint32_t)m) - 1u;
                        du[p] = counted ? max(ubk, (uint32_t)m) - c.M : 0xffffffffu;
                        dmin = min(dmin, du[p]);
                    }
                }
            }
        };
        static_for(pass, std::make_integer_sequence<int, NP>{});
        if (LONGSEQ && __any_sync(FULL, slowmask != 0u)) {
            /* some compare left the window: finish those cells from the packed pool (and what was taken
             * from them on the way), then move the windows on */
            uint32_t wq = 0, wt = 0;
            auto fix = [&](auto pc) {
                constexpr int p = decltype(pc)::value;
                if (slowmask >> p & 1u) {
                    const int k = k0 + 32 * p;
                    const uint32_t M = extend_global(Mn[p], k), ubk = (uint32_t)(n + k);
                    Mn[p] = M;
                    sts32<128 * p>(aC, M);
                    grow[32 * p] = SC::pack(M, lds32<128 * p>(aF), lds32<128 * p + (int)ROWB>(aF));
                    wq = max(wq, (uint32_t)((int)M - k) >> 4); wt = max(wt, M >> 4);
                    hit = hit || (k == Ak && M >= (uint32_t)m);
                    if (ADAPT) du[p] = M - 1u < min(ubk, (uint32_t)m) - 1u ? max(ubk, (uint32_t)m) - M : 0xffffffffu;
                }
            };
            static_for(fix, std::make_integer_sequence<int, NP>{});
            if (ADAPT) {
                dmin = 0xffffffffu;
#pragma unroll
                for (int p = 0; p < NP; p++) dmin = min(dmin, du[p]);
            }
            wq = __reduce_max_sync(FULL, wq); wt = __reduce_max_sync(FULL, wt);
            Q.advance_to(wq, lane); T.advance_to(wt, lane);
        }
        const int wlo = __reduce_min_sync(FULL, pmin), whi = __reduce_max_sync(FULL, pmax);
        const uint32_t mind = ADAPT ? __reduce_min_sync(FULL, dmin) : 0u;
        endhit = __any_sync(FULL, hit);
        __syncwarp();                                               /* the row is in the ring */
        row_exists = wlo <= whi;
        elo = wlo; ehi = whi; width = whi - wlo + 1;                 /* C counts the row before reduce */
        if (ADAPT && row_exists && !endhit && whi - wlo + 1 >= min_wf_len) {
            /* reduce (wfa.go:461-540) as reductions over the lanes' cells: a cell is
             * near iff it counts and d - min <= MaxDistDiff (an uncounted one wraps to a huge value) */
            bool anyfar = false; int fk = INT_MAX, Lk = INT_MIN;
#pragma unroll
            for (int p = 0; p < NP; p++) {
                const uint32_t t = du[p] - mind;
                if (t <= (uint32_t)maxdiff) { fk = min(fk, k0 + 32 * p); Lk = k0 + 32 * p; }
                else if (du[p] != 0xffffffffu) anyfar = true;
            }
            if (__any_sync(FULL, anyfar)) {
                const int fmin = __reduce_min_sync(FULL, fk);
                ehi = __reduce_max_sync(FULL, Lk);
                int lf = INT_MIN;
#pragma unroll
                for (int p = 0; p < NP; p++) if (du[p] != 0xffffffffu && k0 + 32 * p < fmin) lf = k0 + 32 * p;
                lf = __reduce_max_sync(FULL, lf);
                if (lf != INT_MIN) elo = lf + 1;
            }
        }
    };

    while (!finished) {
        si++;
        recM = recM == rM + 4 * ROWB ? rM : recM + ROWB; recE = recE == rE ? rE + 2 * ROWB : rE;
        /* loop range of next (wfa.go:557-563): hull of the source rows +- 1, clamped */
        lo = max(min(min(lo1, lo2), lo4) - 1, -(n - 1)); hi = min(max(max(hi1, hi2), hi4) + 1, m - 1);
        int4 hc = make_int4(0, 1, 0, 0);
        endhit = false; row_exists = false;
        room -= (int)HDR_CELLS;
        if (lo <= hi) {
            aw = hi - lo + 1;
            if (aw > WR) { status = ST_RING; break; }
            room -= aw;
            if (room < 0) { status = ST_ARENA; break; }
            off = top - (uint32_t)aw;
            const int np = (aw + 31) >> 5;
            if (MAXP > 4 && np > 4) row(std::integral_constant<int, MAXP>{}, std::true_type{});
            else if (np == 1) row(std::integral_constant<int, 1>{}, std::false_type{});
            else if (np == 2) row(std::integral_constant<int, 2>{}, std::false_type{});
            else if (np == 3) row(std::integral_constant<int, 3>{}, std::false_type{});
            else row(std::integral_constant<int, 4>{}, std::false_type{});
            if (row_exists) {
                top = off;
                c_steps++; c_cells += (uint32_t)width;
                hc = make_int4(lo, elo, ehi, (int)off);
            } else room += aw;
        }
        if (room < 0) { status = ST_ARENA; break; }
        if (lane == 0) *reinterpret_cast<int4 *>(hdrs + si) = hc;
        lo4 = lo3; hi4 = hi3; lo3 = lo2; hi3 = hi2; lo2 = lo1; hi2 = hi1;
        lo1 = row_exists ? elo : SLIM_NONE_LO; hi1 = row_exists ? ehi : SLIM_NONE_HI;
        zM4 = zM3; zM3 = zM2; zM2 = zM1;
        zM1 = recM - (row_exists ? (uint32_t)lo * 4u : 0u); zE1 = recE - (row_exists ? (uint32_t)lo * 4u : 0u);
        if (endhit) { minS = (uint32_t)si * P.g; finished = true; }
    }

    f.status = status; f.minS = minS; f.lastK = Ak; f.si = si; f.top = (uint64_t)top;
    f.c_cells = c_cells; f.c_written = slot_cells - top; f.c_steps = c_steps;     /* every row stays in the slot */
    f.first_eq = first_eq;
    return f;
}

/* Component.Get (wfa_component.go:142-155) on a SLIM slot; like LaneView it re-derives the
 * provenance code of the cell the backtrace stands on from the cell's five sources, exactly as
 * `next` chose it (wfa.go:579-698), and remembers those five words: the next cell of the walk
 * and the offsets the reference re-derives there (wfa.go:766-817) are always among them. */
template <int SZ> struct SlimView {
    typedef SlimCell<SZ> SC;
    typedef typename SC::T CellT;
    const SlimHdr *hdr; const CellT *cells;
    int si_last, n, m;
    bool first_eq;
    int c_si, c_k; CellT c_w[5];
    __device__ __forceinline__ CellT word(int si, int k) const
    {
        if (si < 0 || si > si_last) return 0;
        const int4 h = *reinterpret_cast<const int4 *>(hdr + si);
        if (k < h.y || k > h.z) return 0;
        return cells[(uint32_t)h.w + (uint32_t)(k - h.x)];
    }
    __device__ __forceinline__ CellT cached_word(int si, int k) const
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
        /* the walk goes on 1, 2 or 4 rows down: have the headers it will need next on their way */
        if (si >= 12) asm volatile("prefetch.global.L1 [%0];" :: "l"(hdr + si - 12));
        const CellT wl = word(si - SLIM_OEG, k - 1), el = word(si - SLIM_EG, k - 1);
        const CellT wr = word(si - SLIM_OEG, k + 1), er = word(si - SLIM_EG, k + 1);
        const CellT wx = word(si - SLIM_XG, k);
        c_si = si; c_k = k; c_w[0] = wl; c_w[1] = el; c_w[2] = wr; c_w[3] = er; c_w[4] = wx;
        const CellO c = next_off(SC::get(wl, 0), SC::get(el, 1), SC::get(wr, 0), SC::get(er, 2), SC::get(wx, 0),
                                 (uint32_t)m, (uint32_t)(n + k));
        uint32_t code;
        if (comp == 1) code = T_INS_OPEN + ((c.code >> 3) & 1u);
        else if (comp == 2) code = T_DEL_OPEN + ((c.code >> 4) & 1u);
        else code = c.M ? (c.code & 7u) : (si == 0 ? T_MATCH : T_MISMATCH);      /* init cell: M[0] holds the matching start cells, M[x] the others (wfa.go:155-183) */
        return o << T_BITS | code;
    }
};

/* backtraces (wfa.go:703-983) of a group's pairs, lane-parallel (lane j owns pair j and its sub-slot) */
template <int SZ>
__device__ __noinline__ void finish_group_slim(const KParams &P, const bool have, const uint32_t pair, const FwdOut &f, uint8_t *slot, const uint64_t slot_bytes)
{
    typedef typename SlimCell<SZ>::T CellT;
    uint32_t *words = reinterpret_cast<uint32_t *>(slot);
    const uint64_t slot_words = slot_bytes >> 2, top_w = f.top * (sizeof(CellT) / 4);
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
        SlimView<SZ> A; A.hdr = reinterpret_cast<const SlimHdr *>(slot); A.cells = reinterpret_cast<const CellT *>(slot);
        A.si_last = f.si; A.n = f.n; A.m = f.m; A.first_eq = f.first_eq; A.c_si = -1; A.c_k = 0;
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

#ifndef WFA_SLIM_MINB
#define WFA_SLIM_MINB 7
#endif
template <int MAXP, int SZ, bool ADAPT>
__global__ void __launch_bounds__(128, SZ == 2 ? 4 : WFA_SLIM_MINB)        /* long pairs come in small numbers: registers before residency */
slim_kernel(const KParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wib = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)slim_smem_pad(MAXP) + (uint32_t)wib * (uint32_t)slim_smem_bytes(MAXP);
    const uint64_t worker = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    uint8_t *slot = P.arena + worker * P.slot_bytes;
    const uint32_t G = (uint32_t)P.group;
    const uint64_t sub_bytes = P.slot_bytes / G;
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = (uint32_t)atomicAdd(&P.ctr->work_next, (unsigned long long)G);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= P.n_work) break;
        const uint32_t cnt = min(G, P.n_work - first);
        FwdOut mine; mine.status = ST_PENDING; mine.minS = 0; mine.lastK = 0; mine.si = 0; mine.n = mine.m = 0; mine.top = 0;
        mine.c_cells = mine.c_written = mine.c_steps = 0; mine.first_eq = false;
        bool have = false; uint32_t my_pair = 0;
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t pair = P.work ? P.work[first + j] : first + j;
            if (P.pflags[pair] & 1) {
                if (lane == 0) {
                    const unsigned long long r = atomicAdd(P.retry_ctr, 1ull);
                    P.retry[r] = (uint64_t)ST_NEED8 << 32 | pair;
                }
                continue;
            }
            const FwdOut f = forward_slim<MAXP, SZ, ADAPT>(P, pair, smem_sa, slot + (uint64_t)j * sub_bytes, sub_bytes);
            if (lane == (int)j) { mine = f; have = true; my_pair = pair; }
        }
        finish_group_slim<SZ>(P, have, my_pair, mine, slot + (uint64_t)lane * sub_bytes, sub_bytes);
    }
}

} /* namespace wfak */
