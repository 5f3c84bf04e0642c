/*
 * wfa_reg.cuh -- REG worker of libwfacuda.so (sm_100a): one warp per pair with the live
 * wavefront rows in REGISTERS, for global alignments whose wavefronts stay narrow (wf-adaptive
 * reduction, or few errors) under penalties of the default shape x : o+e : e = 2 : 4 : 1
 * (4/6/2 and its multiples).  Configs 3 and 5.
 *
 * Why (profiles/r1_cfg3_warp.md): the WARP worker of wfa_kernels.cuh spends ~590 warp
 * instructions per score step of a ~30-cell row, half of them bookkeeping around a
 * shared-memory ring (ring slots, three row headers, range tests on every source load,
 * provenance codes for every cell).  Here
 *   - diagonal k lives in column c = k mod W (W = 32 S) for the whole pair, column c in lane
 *     c / S, register c % S: a row of up to W diagonals is S cells per lane, and the rows
 *     `next` reads (M[s-x], M[s-o-e], I[s-e], D[s-e], wfa.go:579-651) are register arrays.  A
 *     cell's neighbours k-1 / k+1 are the lane's own neighbouring registers, except at the
 *     lane's edge: FOUR shuffles per score step bring those in, whatever S is;
 *   - a register that holds no live cell is 0 = absent, and every live source cell of a row
 *     lies inside the row's loop range (wfa.go:557-563: hull of the sources +- 1), which is at
 *     most W wide -- so a column identifies its diagonal and no source needs a range test;
 *   - only offsets are computed (as in the LANE class): the provenance code is a function of
 *     the five source offsets and is re-derived by the backtrace for the cells it visits;
 *   - the arena holds ONE word per (score, diagonal): M | I << 11 | D << 22 (targets up to
 *     2046 bases) or M | I << 21 | D << 42 in 64 bits, plus a 16-byte header per score;
 *   - the four row ranges `next` needs are kept in registers and rotated;
 *   - both sequences are read through a 2 KB shared-memory window per warp that follows the
 *     front (one LDS.64 + one funnel shift per 16-base compare), so `extend` of a 100 kbp pair
 *     does not go to L2.
 * A pair whose row outgrows W is reported as ST_RING and re-queued with a larger S or on the
 * WARP worker.  Semantics follow the reference at /root/reference (cited as wfa.go:LINE).
 */
#pragma once
#include "wfa_kernels.cuh"
#include "wfa_lane.cuh"

namespace wfak {

constexpr int REG_XG = 2, REG_OEG = 4, REG_EG = 1;     /* x, o+e, e in units of g */
constexpr int REG_WIN = 128;                           /* sequence window: 128 entries of 16 bases per sequence */
constexpr uint32_t REG_MAX_M32 = 2046;                 /* offsets up to m+1 must fit 11 bits */
constexpr uint32_t REG_MAX_M64 = (1u << 21) - 2;       /* ... or 21 bits */
constexpr int REG_NONE_LO = 1 << 30, REG_NONE_HI = -(1 << 30);

/* One score's row in a REG slot: cells of diagonals [alo, alo + aw) start at cell index `off`;
 * [lo, hi] = M WaveFront.Lo/Hi after reduce (outside: absent, wfa.go:526-537), lo > hi: no such score. */
struct RegHdr { int32_t alo, lo, hi; uint32_t off; };

template <bool WIDE> struct RegCell {
    typedef typename std::conditional<WIDE, uint64_t, uint32_t>::type T;
    static constexpr int BITS = WIDE ? 21 : 11;
    __device__ static __forceinline__ T pack(uint32_t M, uint32_t I, uint32_t D)
    {
        if (WIDE) return (uint64_t)M | (uint64_t)I << 21 | (uint64_t)D << 42;
        return (T)(M | I << 11 | D << 22);
    }
    __device__ static __forceinline__ uint32_t get(T w, int comp)
    {
        return (uint32_t)(w >> (BITS * comp)) & ((1u << BITS) - 1u);
    }
};

__host__ __device__ inline size_t reg_smem_bytes() { return 2 * (size_t)REG_WIN * 8; }   /* per warp */

/* A sequence seen through the warp's shared-memory window: entry j of the ring holds the 2-bit
 * words j and j+1 of the sequence, for j in [wbase, wend); anything else is read from the packed
 * pool (correct, only slower).  The window moves forward when a lane ran past its end. */
struct SeqWin {
    const uint32_t *g;
    uint32_t sa;            /* shared-window byte address of ring entry 0 */
    uint32_t wbase, wend, nwords;
    __device__ __forceinline__ void fill(uint32_t from, uint32_t to, int lane)
    {
        for (uint32_t j = from + (uint32_t)lane; j < to; j += 32) {
            const uint32_t a = __ldg(g + j), b = __ldg(g + j + 1);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(sa + (j & (REG_WIN - 1)) * 8u), "r"(a), "r"(b) : "memory");
        }
    }
    __device__ __forceinline__ void init(const uint32_t *words, uint32_t n_sym, uint32_t ring_sa, int lane)
    {
        g = words; sa = ring_sa; nwords = (n_sym + 15u) >> 4;
        wbase = 0; wend = min(nwords, (uint32_t)REG_WIN);
        fill(0, wend, lane);
    }
    /* 16 bases from base `pos` on (base pos in the low bits); `ahead` is raised when the read ran past the window */
    __device__ __forceinline__ uint32_t chunk(uint32_t pos, bool &ahead) const
    {
        const uint32_t wi = pos >> 4;
        uint32_t a, b;
        if (wi - wbase < wend - wbase) {
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sa + (wi & (REG_WIN - 1)) * 8u));
        } else {
            a = __ldg(g + wi); b = __ldg(g + wi + 1);
            ahead = ahead || wi >= wend;
        }
        return __funnelshift_r(a, b, pos * 2u);              /* the shift count wraps at 32: (pos % 16) * 2 */
    }
    /* warp-uniform: move the window 32 entries on (its oldest 32 are dropped) */
    __device__ __forceinline__ void advance(int lane)
    {
        if (wend >= nwords) return;
        const uint32_t to = min(nwords, wend + 32u);
        __syncwarp();
        fill(wend, to, lane);
        wend = to; wbase = wend > (uint32_t)REG_WIN ? wend - (uint32_t)REG_WIN : 0u;
        __syncwarp();
    }
};

/* Forward pass of one pair: wfa.go:228-251 with next + extend fused per cell. */
template <int S, bool WIDE, bool ADAPT>
__device__ __forceinline__ FwdOut forward_reg(const KParams &P, const uint32_t pair, const uint32_t smem_sa, uint8_t *slot, const uint64_t slot_bytes)
{
    constexpr int W = 32 * S;
    constexpr uint32_t FULL = 0xffffffffu;
    typedef RegCell<WIDE> RC;
    typedef typename RC::T CellT;
    constexpr uint32_t HDR_CELLS = sizeof(RegHdr) / sizeof(CellT);
    const int lane = threadIdx.x & 31;
    const PairDesc pd = P.pairs[pair];
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const int maxdiff = P.max_dist_diff, min_wf_len = P.min_wf_len;

    FwdOut f;
    f.status = ST_OK; f.minS = 0; f.lastK = Ak; f.si = 0; f.n = n; f.m = m; f.top = 0;
    f.c_cells = f.c_written = f.c_steps = 0;
    if ((uint32_t)m > (WIDE ? REG_MAX_M64 : REG_MAX_M32)) { f.status = ST_RING; return f; }   /* offsets would not fit the cell word */

    SeqWin Q, T;
    __syncwarp();
    Q.init(P.packed + pd.q_word, (uint32_t)n, smem_sa, lane);
    T.init(P.packed + pd.t_word, (uint32_t)m, smem_sa + (uint32_t)REG_WIN * 8u, lane);
    __syncwarp();

    RegHdr *hdrs = reinterpret_cast<RegHdr *>(slot);                    /* grows up, index s/g */
    CellT  *cells = reinterpret_cast<CellT *>(slot);                    /* rows grow down from the end */
    const uint32_t slot_cells = (uint32_t)min(slot_bytes / sizeof(CellT), (uint64_t)0xfffffff0u);
    uint32_t top = slot_cells;
    uint32_t hdr_limit = 3 * HDR_CELLS + 8;                             /* cells covered by headers incl. the next one + slack */

    /* extend (wfa.go:394-455) of a present cell: offset h on diagonal k */
    bool aheadQ = false, aheadT = false;
    auto extend = [&](uint32_t M, int k) -> uint32_t {
        const int h = (int)M, v = h - k;
        /* a present cell has v >= 1, so "v > 0, v < n, h < m" is "min(n-v, m-h) > 0" */
        const int ext = min(n - v, m - h);
        if (M == 0u || ext <= 0) return M;
        int l = 0;
        do {
            const uint32_t xx = Q.chunk((uint32_t)(v + l), aheadQ) ^ T.chunk((uint32_t)(h + l), aheadT);
            if (xx) { l += __clz((int)__brev(xx)) >> 1; break; }
            l += 16;
        } while (l < ext);
        return M + (uint32_t)min(l, ext);
    };

    /* rows s-1 .. s-4 of M, row s-1 of I and D; this lane's columns are lane * S + i */
    uint32_t M1[S], M2[S], M3[S], M4[S], I1[S], D1[S];
#pragma unroll
    for (int i = 0; i < S; i++) M1[i] = M2[i] = M3[i] = M4[i] = I1[i] = D1[i] = 0u;
    int lo1 = REG_NONE_LO, hi1 = REG_NONE_HI, lo2 = REG_NONE_LO, hi2 = REG_NONE_HI;
    int lo3 = REG_NONE_LO, hi3 = REG_NONE_HI, lo4 = REG_NONE_LO, hi4 = REG_NONE_HI;
    const int c0 = lane * S;

    uint32_t c_cells = 0, c_written = 0, c_steps = 0;
    int status = ST_OK, si = 0;
    uint32_t minS = 0;
    bool finished = false;

    /* initComponents (wfa.go:155-158): M[0][0] or M[x][0] = 1 */
    bool dummy = false;
    const bool first_eq = ((Q.chunk(0u, dummy) ^ T.chunk(0u, dummy)) & 3u) == 0u;
    {
        RegHdr h0; h0.alo = 0; h0.lo = 1; h0.hi = 0; h0.off = 0;
        if (first_eq) {
            const uint32_t Mx = extend(1u, 0);                         /* the same on every lane */
            top -= 1;
            if (lane == 0) { M1[0] = Mx; cells[top] = RC::pack(Mx, 0u, 0u); }
            h0.lo = 0; h0.hi = 0; h0.off = top;
            lo1 = hi1 = 0;
            c_steps = 1; c_cells = 1; c_written = 1;
            if (Ak == 0 && (int)Mx >= m) finished = true;               /* wfa.go:235-239 */
        }
        if (lane == 0) hdrs[0] = h0;
        hdr_limit += HDR_CELLS;
    }

    while (!finished) {
        si++;
        /* loop range of next (wfa.go:557-563): hull of the source rows +- 1, clamped */
        int lo = min(min(lo1, lo2), lo4) - 1, hi = max(max(hi1, hi2), hi4) + 1;
        lo = max(lo, -(n - 1)); hi = min(hi, m - 1);
        const bool init = si == REG_XG && !first_eq;                    /* the row of score x starts with M[x][0] */
        if (init) { lo = min(lo, 0); hi = max(hi, 0); }
        uint32_t Mn[S], In[S], Dn[S];
#pragma unroll
        for (int i = 0; i < S; i++) Mn[i] = In[i] = Dn[i] = 0u;
        int elo = REG_NONE_LO, ehi = REG_NONE_HI;
        RegHdr hc; hc.alo = 0; hc.lo = 1; hc.hi = 0; hc.off = 0;
        bool endhit = false;
        if (lo <= hi) {
            const int aw = hi - lo + 1;
            if (aw > W) { status = ST_RING; break; }
            if (top < hdr_limit || top - hdr_limit < (uint32_t)aw) { status = ST_ARENA; break; }
            const uint32_t off = top - (uint32_t)aw;
            /* this lane's diagonals: the one of [lo, lo + W) in each of its columns */
            int lom = lo % W; lom += lom < 0 ? W : 0;
            int k[S];
#pragma unroll
            for (int i = 0; i < S; i++) { int t = c0 + i - lom; t += t < 0 ? W : 0; k[i] = lo + t; }
            /* sources across the lane's edges */
            const uint32_t moL = __shfl_sync(FULL, M4[S - 1], (lane + 31) & 31), moR = __shfl_sync(FULL, M4[0], (lane + 1) & 31);
            const uint32_t ieL = __shfl_sync(FULL, I1[S - 1], (lane + 31) & 31), deR = __shfl_sync(FULL, D1[0], (lane + 1) & 31);
            int pmin = INT_MAX, pmax = INT_MIN;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const bool act = k[i] <= hi;
                const uint32_t um = act ? (uint32_t)m : 0u, ubk = act ? (uint32_t)(n + k[i]) : 0u;
                Cell3O c = next_off3(i ? M4[i - 1] : moL, i ? I1[i - 1] : ieL, i < S - 1 ? M4[i + 1] : moR, i < S - 1 ? D1[i + 1] : deR,
                                     M2[i], um, ubk);
                if (init && k[i] == 0 && c.M == 0u) c.M = 1u;            /* unless next's Set wrote the cell (wfa_wavefront.go:93) */
                c.M = extend(c.M, k[i]);
                Mn[i] = c.M; In[i] = c.I; Dn[i] = c.D;
                if (act) cells[off + (uint32_t)(k[i] - lo)] = RC::pack(c.M, c.I, c.D);
                if (c.M) { pmin = min(pmin, k[i]); pmax = max(pmax, k[i]); if (k[i] == Ak && (int)c.M >= m) endhit = true; }
            }
            const int wlo = __reduce_min_sync(FULL, pmin), whi = __reduce_max_sync(FULL, pmax);
            endhit = __any_sync(FULL, endhit);
            if (wlo <= whi) {
                top = off;
                c_steps++; c_cells += (uint32_t)(whi - wlo + 1); c_written += (uint32_t)aw;
                elo = wlo; ehi = whi;
                if (ADAPT && !endhit && whi - wlo + 1 >= min_wf_len) {
                    /* reduce (wfa.go:461-540) as reductions over the lanes' cells */
                    int d[S], dmin = INT_MAX;
#pragma unroll
                    for (int i = 0; i < S; i++) {
                        const int h = (int)Mn[i], v = h - k[i];
                        d[i] = (Mn[i] != 0u && v < n && h < m) ? max(m - h, n - v) : -1;      /* v >= 1 for a present cell */
                        if (d[i] >= 0) dmin = min(dmin, d[i]);
                    }
                    const int mind = __reduce_min_sync(FULL, dmin);
                    bool anyfar = false; int fk = INT_MAX, Lk = INT_MIN;
#pragma unroll
                    for (int i = 0; i < S; i++) if (d[i] >= 0) {
                        if (d[i] - mind > maxdiff) anyfar = true;
                        else { fk = min(fk, k[i]); Lk = max(Lk, k[i]); }
                    }
                    if (__any_sync(FULL, anyfar)) {
                        const int fmin = __reduce_min_sync(FULL, fk);
                        ehi = __reduce_max_sync(FULL, Lk);
                        int lf = INT_MIN;
#pragma unroll
                        for (int i = 0; i < S; i++) if (d[i] >= 0 && k[i] < fmin) lf = max(lf, k[i]);
                        lf = __reduce_max_sync(FULL, lf);
                        if (lf != INT_MIN) elo = lf + 1;
#pragma unroll
                        for (int i = 0; i < S; i++) if (k[i] < elo || k[i] > ehi) Mn[i] = In[i] = Dn[i] = 0u;   /* Delete, wfa.go:526-535 */
                    }
                }
                hc.alo = lo; hc.lo = elo; hc.hi = ehi; hc.off = off;
            }
        }
        if (top < hdr_limit) { status = ST_ARENA; break; }
        if (lane == 0) *reinterpret_cast<int4 *>(hdrs + si) = make_int4(hc.alo, hc.lo, hc.hi, (int)hc.off);
        hdr_limit += HDR_CELLS;
        /* rows move on by one score */
#pragma unroll
        for (int i = 0; i < S; i++) { M4[i] = M3[i]; M3[i] = M2[i]; M2[i] = M1[i]; M1[i] = Mn[i]; I1[i] = In[i]; D1[i] = Dn[i]; }
        lo4 = lo3; hi4 = hi3; lo3 = lo2; hi3 = hi2; lo2 = lo1; hi2 = hi1; lo1 = elo; hi1 = ehi;
        if (endhit) { minS = (uint32_t)si * P.g; finished = true; }
        /* sequence windows follow the front */
        if (__any_sync(FULL, aheadQ)) { Q.advance(lane); aheadQ = false; }
        if (__any_sync(FULL, aheadT)) { T.advance(lane); aheadT = false; }
    }

    f.status = status; f.minS = minS; f.lastK = Ak; f.si = si; f.top = (uint64_t)top;
    f.c_cells = c_cells; f.c_written = c_written; f.c_steps = c_steps;
    /* what the backtrace needs besides the arena */
    f.n = n; f.m = m;
    f.first_eq = first_eq;
    return f;
}

/* Component.Get (wfa_component.go:142-155) on a REG slot; like LaneView it re-derives the
 * provenance code of the cell the backtrace stands on from the cell's five sources, exactly as
 * `next` chose it (wfa.go:579-698), and remembers those five words: the next cell of the walk
 * and the offsets the reference re-derives there (wfa.go:766-817) are always among them. */
template <bool WIDE> struct RegView {
    typedef RegCell<WIDE> RC;
    typedef typename RC::T CellT;
    const RegHdr *hdr; const CellT *cells;
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
            if (dk == -1) { if (ds == REG_OEG) return c_w[0]; if (ds == REG_EG) return c_w[1]; }
            else if (dk == 1) { if (ds == REG_OEG) return c_w[2]; if (ds == REG_EG) return c_w[3]; }
            else if (dk == 0 && ds == REG_XG) return c_w[4];
        }
        return word(si, k);
    }
    __device__ __forceinline__ uint32_t get(int comp, int si, int k) const { return RC::get(cached_word(si, k), comp) << T_BITS; }
    __device__ __forceinline__ uint32_t get_typed(int comp, int si, int k)
    {
        const uint32_t o = RC::get(cached_word(si, k), comp);
        if (o == 0) return 0;
        const CellT wl = word(si - REG_OEG, k - 1), el = word(si - REG_EG, k - 1);
        const CellT wr = word(si - REG_OEG, k + 1), er = word(si - REG_EG, k + 1);
        const CellT wx = word(si - REG_XG, k);
        c_si = si; c_k = k; c_w[0] = wl; c_w[1] = el; c_w[2] = wr; c_w[3] = er; c_w[4] = wx;
        const CellO c = next_off(RC::get(wl, 0), RC::get(el, 1), RC::get(wr, 0), RC::get(er, 2), RC::get(wx, 0),
                                 (uint32_t)m, (uint32_t)(n + k));
        uint32_t code;
        if (comp == 1) code = T_INS_OPEN + ((c.code >> 3) & 1u);
        else if (comp == 2) code = T_DEL_OPEN + ((c.code >> 4) & 1u);
        else code = c.M ? (c.code & 7u) : (first_eq ? T_MATCH : T_MISMATCH);
        return o << T_BITS | code;
    }
};

/* backtraces (wfa.go:703-983) of a group's pairs, lane-parallel (lane j owns pair j and its sub-slot) */
template <bool WIDE>
__device__ __noinline__ void finish_group_reg(const KParams &P, const bool have, const uint32_t pair, const FwdOut &f, uint8_t *slot, const uint64_t slot_bytes)
{
    typedef typename RegCell<WIDE>::T CellT;
    uint32_t *words = reinterpret_cast<uint32_t *>(slot);
    const uint64_t slot_words = slot_bytes >> 2, top_w = f.top * (sizeof(CellT) / 4);
    const uint64_t scratch_w = (((uint64_t)(f.si + 1) * sizeof(RegHdr) + 7) / 8) * 2;
    uint64_t *scratch = reinterpret_cast<uint64_t *>(words + scratch_w);
    int status = have ? f.status : ST_PENDING;

    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;
    uint32_t n_ops = 0;
    __syncwarp();
    if (status == ST_OK) {
        RegView<WIDE> A; A.hdr = reinterpret_cast<const RegHdr *>(slot); A.cells = reinterpret_cast<const CellT *>(slot);
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

#ifndef WFA_REG_MINB
#define WFA_REG_MINB 4
#endif
template <int S, bool WIDE, bool ADAPT>
__global__ void __launch_bounds__(128, WFA_REG_MINB)
reg_kernel(const KParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wib = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)wib * (uint32_t)reg_smem_bytes();
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
            const FwdOut f = forward_reg<S, WIDE, ADAPT>(P, pair, smem_sa, slot + (uint64_t)j * sub_bytes, sub_bytes);
            if (lane == (int)j) { mine = f; have = true; my_pair = pair; }
        }
        finish_group_reg<WIDE>(P, have, my_pair, mine, slot + (uint64_t)lane * sub_bytes, sub_bytes);
    }
}

} /* namespace wfak */
