/*
 * wfa_lane.cuh -- LANE worker of libwfacuda.so (sm_100a): one pair per lane, 32 pairs per warp
 * in lockstep, for short global alignments without heuristic (config 2: 150 bp reads).
 *
 * Why: a warp that spreads ONE short pair's diagonals over its lanes (WARP worker) spends
 * ~500 warp instructions per score on bookkeeping that is uniform over the wavefront, with
 * half of the lanes idle in the cell loop (ncu: 12.1 k warp instructions per 150 bp pair, 83 %
 * issue-slot utilisation -- the kernel is issue bound, not memory bound).  Here every lane owns
 * a whole pair and the warp walks (score, diagonal) in lockstep: without heuristic the loop
 * range of `next` (wfa.go:557-563) depends only on which scores exist, so the 32 ranges nearly
 * coincide and their union is used -- outside its own range a lane has no source cell and
 * computes "absent", exactly what the reference's narrower loop leaves behind.  The per-score
 * bookkeeping is paid once per 32 pairs and there is no cross-lane traffic at all.
 *
 * Per-lane private data (shared memory, lane-interleaved so that every access is conflict
 * free whatever the lane's own index):
 *   seqQ, seqT   u32 [17][32]          2-bit packed sequences (<= 254 bases)
 *   ringM        u8  [dM][W][32]       offsets of the last max(x,o+e)/g+1 scores of M
 *   ringI, ringD u8  [dE][W][32]       offsets of the last e/g+1 scores of I and D
 * Offsets fit a byte (<= m+1 <= 255); the 3-bit provenance codes are not needed by `next`.
 * Diagonal k lives at column k + W/2 of every row.
 *
 * Backtrace arena (HBM, one slot per warp, reused group after group):
 *   hdr   int4 {lo, hi, off, aw} per score index, shared by the 32 pairs
 *   cell  u32 [aw][32] per score: M | I<<8 | D<<16 | codeM<<24 | extI<<27 | extD<<28
 * i.e. 4 bytes per (score, diagonal, pair) instead of the 12 of three raw words; `LaneView::get`
 * rebuilds the reference's raw word offset<<3|code for the unchanged literal backtrace.
 *
 * Semantics follow the reference at /root/reference (cited as wfa.go:LINE).
 */
#pragma once
#include "wfa_kernels.cuh"

namespace wfak {

constexpr int LANE_SEQ_WORDS = 17;          /* 16 words = 256 bases + 1 for the funnel shift */
constexpr int LANE_MAX_LEN = 254;           /* offsets up to m+1 must fit a byte */
#ifndef WFA_LANE_WARPS
#define WFA_LANE_WARPS 2
#endif

__host__ __device__ inline size_t lane_smem_bytes(int dM, int dE, int W)
{
    size_t b = (((size_t)dM * 8) + 15) & ~(size_t)15;                 /* meta int2[dM] */
    b += 2 * (size_t)LANE_SEQ_WORDS * 128;                             /* seqQ, seqT */
    b += ((size_t)dM + 2 * (size_t)dE) * (size_t)W * 32;               /* rings */
    return (b + 127) & ~(size_t)127;
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

/* Component.Get on the packed group arena (see the header comment) */
struct LaneView {
    const int4     *hdr;
    const uint32_t *cells;     /* already offset by the lane */
    int             si_last;
    __device__ __forceinline__ uint32_t get(int comp, int si, int k) const
    {
        if (si < 0 || si > si_last) return 0;
        const int4 h = hdr[si];
        if (k < h.x || k > h.y) return 0;
        const uint32_t w = cells[(uint32_t)h.z + (uint32_t)(k - h.x) * 32u];
        if (comp == 0) { const uint32_t o = w & 255u; return o ? (o << T_BITS | ((w >> 24) & 7u)) : 0u; }
        if (comp == 1) { const uint32_t o = (w >> 8) & 255u; return o ? (o << T_BITS | (T_INS_OPEN + ((w >> 27) & 1u))) : 0u; }
        const uint32_t o = (w >> 16) & 255u;
        return o ? (o << T_BITS | (T_DEL_OPEN + ((w >> 28) & 1u))) : 0u;
    }
};

/* One diagonal of `next` (wfa.go:572-699) on bare offsets (0 = absent); same candidate packing
 * as next_cell: value<<p | priority, one max() picks offset and provenance.
 *   um = m (offset <= m, :581,:585,:651), ubk = n + k (offset - k <= n, :616,:620,:651) */
struct CellO { uint32_t M, I, D, code; };   /* code = codeM | extI<<3 | extD<<4 */

__device__ __forceinline__ CellO next_off(uint32_t mo_l, uint32_t ie_l, uint32_t mo_r, uint32_t de_r, uint32_t mx,
                                          uint32_t um, uint32_t ubk)
{
    CellO r;
    uint32_t ca = (mo_l - 1u) < um ? (mo_l << 1 | 1u) : 0u;
    uint32_t cb = (ie_l - 1u) < um ? (ie_l << 1) : 0u;
    uint32_t best = max(ca, cb);
    r.I = best ? (best >> 1) + 1u : 0u;
    const uint32_t extI = (best & 1u) ^ 1u;                        /* 0 = InsertOpen, 1 = InsertExt */
    ca = (mo_r - 1u) < ubk ? (mo_r << 1 | 1u) : 0u;
    cb = (de_r - 1u) < ubk ? (de_r << 1) : 0u;
    best = max(ca, cb);
    r.D = best >> 1;
    const uint32_t extD = (best & 1u) ^ 1u;
    const uint32_t cx = (mx - 1u) < min(um, ubk) ? ((mx + 1u) << 2 | 2u) : 0u;
    const uint32_t ci = r.I ? (r.I << 2 | 1u) : 0u;
    const uint32_t cd = r.D << 2;
    const uint32_t bestM = max(max(cx, ci), cd);
    const uint32_t src = bestM & 3u;
    const uint32_t tM = src == 2u ? T_MISMATCH : (src == 1u ? T_INS_OPEN + extI : T_DEL_OPEN + extD);
    r.M = bestM >> 2;
    r.code = tM | extI << 3 | extD << 4;
    return r;
}

/* 16 bases starting at base `pos` of a lane's sequence in shared memory (base pos in the low bits) */
__device__ __forceinline__ uint32_t lane_chunk(uint32_t seq_sa, int pos)
{
    const uint32_t a = seq_sa + ((uint32_t)pos >> 4) * 128u;
    return __funnelshift_r(lds_u32(a), lds_u32(a + 128u), ((uint32_t)pos & 15u) * 2u);
}

/* Forward pass + backtrace of one group of up to 32 pairs. */
__device__ __noinline__ void lane_group(const KParams &P, const bool have, const uint32_t pair, unsigned char *smem,
                                        uint8_t *slot, const uint64_t slot_bytes)
{
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int dM = P.dM, dE = P.dE, W = P.ring_cap, KC = W >> 1;
    const int xg = P.xg, oeg = P.oeg, eg = P.eg, x = (int)P.x;

    int2 *meta = reinterpret_cast<int2 *>(smem);
    unsigned char *p = smem + ((((size_t)dM * 8) + 15) & ~(size_t)15);
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(p) + (uint32_t)lane * 4u;
    const uint32_t sT = sQ + LANE_SEQ_WORDS * 128u;
    const uint32_t rM = sQ - (uint32_t)lane * 4u + 2u * LANE_SEQ_WORDS * 128u + (uint32_t)lane;   /* column of this lane */
    const uint32_t rowB = (uint32_t)W * 32u;
    const uint32_t rI = rM + (uint32_t)dM * rowB, rD = rI + (uint32_t)dE * rowB;

    int status = have ? ST_OK : ST_PENDING;
    PairDesc pd; pd.q_byte = pd.t_byte = pd.q_word = pd.t_word = 0; pd.n = pd.m = 0;
    if (have) {
        pd = P.pairs[pair];
        if (P.pflags[pair] & 1) status = ST_NEED8;
    }
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    bool act = status == ST_OK;

    /* sequences -> lane-private shared memory columns */
    {
        const uint32_t *gq = P.packed + pd.q_word, *gt = P.packed + pd.t_word;
        const int wq = act ? (n + 15) >> 4 : 0, wt = act ? (m + 15) >> 4 : 0;
#pragma unroll 1
        for (int w = 0; w < LANE_SEQ_WORDS; w++) {
            sts_u32(sQ + (uint32_t)w * 128u, w < wq ? __ldg(gq + w) : 0u);
            sts_u32(sT + (uint32_t)w * 128u, w < wt ? __ldg(gt + w) : 0u);
        }
    }
    const bool first_eq = ((lds_u32(sQ) ^ lds_u32(sT)) & 3u) == 0u;       /* q[0] == t[0], wfa.go:155-158 */
    /* a lane works on diagonals [klo, klo + kspan] = [-(n-1), m-1] (wfa.go:562-563); an idle or
     * finished lane gets an empty range and computes nothing but "absent" */
    int klo = act ? -(n - 1) : 0x3fffffff;
    const uint32_t kspan = act ? (uint32_t)(n + m - 2) : 0u;
    const int ulo = -(__reduce_max_sync(FULL, act ? n : 1) - 1), uhi = __reduce_max_sync(FULL, act ? m : 1) - 1;

    int4     *hdrs  = reinterpret_cast<int4 *>(slot);
    uint32_t *cells = reinterpret_cast<uint32_t *>(slot);
    const uint32_t slot_words = (uint32_t)min((uint64_t)0xfffffff0u, slot_bytes >> 2);
    uint32_t top = slot_words;
    uint32_t hdr_limit = 2 * 4 + 64 * 32;              /* header words incl. the next one + a minimum of op scratch */

    for (int i = lane; i < dM; i += 32) meta[i] = make_int2(1, 0);
    __syncwarp();

    uint32_t s = 0; int si = 0, cur = 0, curE = 0;
    bool done = false; uint32_t minS = 0; int my_si = 0;
    uint32_t c_cells = 0, c_written = 0, c_steps = 0;
    int group_fail = 0;

    if (__any_sync(FULL, act)) for (;;) {
        int slX = cur - xg, slO = cur - oeg, slE = cur - eg;
        slX += slX < 0 ? dM : 0; slO += slO < 0 ? dM : 0; slE += slE < 0 ? dM : 0;
        const int2 EMPTY = make_int2(1, 0);
        const int2 hX = si >= xg ? meta[slX] : EMPTY;
        const int2 hO = si >= oeg ? meta[slO] : EMPTY;
        const int2 hE = si >= eg ? meta[slE] : EMPTY;
        int slEe = curE - eg; slEe += slEe < 0 ? dE : 0;
        /* union loop range (wfa.go:557-563), clamped with the longest sequences of the group */
        int lo = INT_MAX, hi = INT_MIN;
        if (hX.x <= hX.y) { lo = min(lo, hX.x); hi = max(hi, hX.y); }
        if (hO.x <= hO.y) { lo = min(lo, hO.x); hi = max(hi, hO.y); }
        if (hE.x <= hE.y) { lo = min(lo, hE.x); hi = max(hi, hE.y); }
        if (lo <= hi) { lo = max(lo - 1, ulo); hi = min(hi + 1, uhi); }
        const bool has_init = (s == 0) || (s == (uint32_t)x);          /* global: the one cell k = 0 */
        if (has_init) { lo = min(lo, 0); hi = max(hi, 0); }

        bool seen = false; int wlo = 0, whi = 0;
        int aw = 0; uint32_t off = 0;
        if (lo <= hi) {
            aw = hi - lo + 1;
            if (lo < -KC || hi > KC - 1) { group_fail = ST_RING; break; }
            const uint32_t need = (uint32_t)aw * 32u;
            if (top < hdr_limit || top - hdr_limit < need) { group_fail = ST_ARENA; break; }
            off = top - need;
            /* lane-private byte columns; cell k of a row is 32*k bytes further */
            const uint32_t bO = rM + (uint32_t)(slO * W + KC) * 32u, bX = rM + (uint32_t)(slX * W + KC) * 32u;
            const uint32_t bI = rI + (uint32_t)(slEe * W + KC) * 32u, bD = rD + (uint32_t)(slEe * W + KC) * 32u;
            const uint32_t bCM = rM + (uint32_t)(cur * W + KC) * 32u;
            const uint32_t bCI = rI + (uint32_t)(curE * W + KC) * 32u, bCD = rD + (uint32_t)(curE * W + KC) * 32u;
            uint32_t *gC = cells + off + lane - lo * 32;
            const uint32_t um = (uint32_t)m;
            auto inO = [&](int k) { return k >= hO.x && k <= hO.y; };
            auto inE = [&](int k) { return k >= hE.x && k <= hE.y; };
            auto inX = [&](int k) { return k >= hX.x && k <= hX.y; };
            auto at = [](uint32_t base, int k) { return base + (uint32_t)(k * 32); };

            /* phase A of one cell: sources -> next -> clamp -> (init) -> first 16-base compare */
            struct Pend { CellO c; int k, ext; uint32_t xr; };
            auto cell_a = [&](auto chk, const int k, const uint32_t mo_l, const uint32_t mo_r) -> Pend {
                constexpr bool CHK = decltype(chk)::value;
                uint32_t ie_l, de_r, mx;
                if (CHK) {
                    ie_l = inE(k - 1) ? lds_u8(at(bI, k - 1)) : 0u;
                    de_r = inE(k + 1) ? lds_u8(at(bD, k + 1)) : 0u;
                    mx = inX(k) ? lds_u8(at(bX, k)) : 0u;
                } else {
                    ie_l = lds_u8(at(bI, k - 1)); de_r = lds_u8(at(bD, k + 1)); mx = lds_u8(at(bX, k));
                }
                Pend q; q.k = k;
                q.c = next_off(mo_l, ie_l, mo_r, de_r, mx, um, (uint32_t)(n + k));
                if ((uint32_t)(k - klo) > kspan) { q.c.M = 0; q.c.I = 0; q.c.D = 0; }
                if (CHK) {
                    if (has_init && k == 0 && q.c.M == 0 && act && !done && (first_eq ? (s == 0) : (s == (uint32_t)x))) {
                        /* initComponents (wfa.go:155-158); next's Set wins when both write */
                        q.c.M = 1u; q.c.code = (q.c.code & ~7u) | (first_eq ? T_MATCH : T_MISMATCH);
                    }
                }
                const int h = (int)q.c.M, v = h - k;
                const bool ex = q.c.M != 0 && v > 0 && v < n && h < m;                /* extend applies (wfa.go:404) */
                q.ext = ex ? min(n - v, m - h) : 0;
                q.xr = lane_chunk(sQ, ex ? v : 0) ^ lane_chunk(sT, ex ? h : 0);
                return q;
            };
            /* phase B: finish extend (wfa.go:411-454), store ring + arena, bookkeeping */
            auto cell_b = [&](Pend &q) {
                const int k = q.k;
                if (q.ext) {
                    int l = q.xr ? (__ffs((int)q.xr) - 1) >> 1 : 16;
                    if (q.xr == 0 && q.ext > 16) {
                        const int h = (int)q.c.M, v = h - k;
                        while (l < q.ext) {
                            const uint32_t xx = lane_chunk(sQ, v + l) ^ lane_chunk(sT, h + l);
                            if (xx) { l += (__ffs((int)xx) - 1) >> 1; break; }
                            l += 16;
                        }
                    }
                    q.c.M += (uint32_t)min(l, q.ext);
                }
                sts_u8(at(bCM, k), q.c.M); sts_u8(at(bCI, k), q.c.I); sts_u8(at(bCD, k), q.c.D);
                gC[k * 32] = q.c.M | q.c.I << 8 | q.c.D << 16 | q.c.code << 24;
                if (q.c.M) { if (!seen) { seen = true; wlo = k; } whi = k; }
            };
            using T_ = std::true_type; using F_ = std::false_type;
            /* interior: every source cell lies inside its row's written range */
            int ia = INT_MAX, ib = INT_MIN;
            if (hX.x <= hX.y && hO.x <= hO.y && hE.x <= hE.y && !has_init) {
                ia = max(max(hX.x, hO.x + 1), max(hE.x + 1, lo));
                ib = min(min(hX.y, hO.y - 1), min(hE.y - 1, hi));
            }
            int k = lo;
            uint32_t o_m1 = inO(k - 1) ? lds_u8(at(bO, k - 1)) : 0u, o_0 = inO(k) ? lds_u8(at(bO, k)) : 0u;
            for (; k <= hi && k < ia; k++) {
                const uint32_t o_p1 = inO(k + 1) ? lds_u8(at(bO, k + 1)) : 0u;
                Pend q = cell_a(T_{}, k, o_m1, o_p1);
                cell_b(q);
                o_m1 = o_0; o_0 = o_p1;
            }
            for (; k + 1 <= ib; k += 2) {
                const uint32_t o_p1 = lds_u8(at(bO, k + 1)), o_p2 = lds_u8(at(bO, k + 2));
                Pend q0 = cell_a(F_{}, k, o_m1, o_p1);
                Pend q1 = cell_a(F_{}, k + 1, o_0, o_p2);
                cell_b(q0); cell_b(q1);
                o_m1 = o_p1; o_0 = o_p2;
            }
            for (; k <= hi; k++) {
                const uint32_t o_p1 = inO(k + 1) ? lds_u8(at(bO, k + 1)) : 0u;
                Pend q = cell_a(T_{}, k, o_m1, o_p1);
                cell_b(q);
                o_m1 = o_0; o_0 = o_p1;
            }
        }
        const bool any = __any_sync(FULL, seen);
        if (any) {
            top = off;
            if (seen) { c_steps++; c_cells += (uint32_t)(whi - wlo + 1); c_written += (uint32_t)aw; }
            /* end test on diagonal m-n (wfa.go:235-239); the lane reads back its own column */
            if (seen && Ak >= lo && Ak <= hi) {
                const uint32_t hM = lds_u8(rM + (uint32_t)(cur * W + KC + Ak) * 32u);
                if ((int)hM >= m) { done = true; minS = s; my_si = si; klo = 0x3fffffff; }
            }
        }
        __syncwarp();
        meta[cur] = any ? make_int2(lo, hi) : make_int2(1, 0);            /* same value from every lane */
        if (lane == 0) hdrs[si] = any ? make_int4(lo, hi, (int)off, aw) : make_int4(1, 0, 0, 0);
        __syncwarp();
        if (__all_sync(FULL, !act || done)) break;
        s += P.g; si++; hdr_limit += 4;
        cur = cur + 1 == dM ? 0 : cur + 1;
        curE = curE + 1 == dE ? 0 : curE + 1;
    }
    if (act && !done) status = group_fail ? group_fail : ST_ARENA;

    /* ---------------- backtrace (wfa.go:703-983), lane-parallel, then the group's results */
    const uint32_t scratch_w = ((((uint32_t)(si + 1) * 16u) + 7u) / 8u) * 2u;
    uint64_t *scratch = reinterpret_cast<uint64_t *>(cells + scratch_w) + lane;
    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;
    uint32_t n_ops = 0;
    __syncwarp();
    __threadfence_block();
    if (status == ST_OK) {
        LaneView A; A.hdr = hdrs; A.cells = cells + lane; A.si_last = my_si;
        OpSink sink; sink.buf = scratch; sink.stride = 32;
        sink.cap = top > scratch_w ? (top - scratch_w) / 64u : 0u;
        sink.n = 0; sink.cur_op = 0; sink.cur_n = 0; sink.overflow = false;
        back_trace(A, P, n, m, minS, Ak, res, sink);
        n_ops = sink.n;
        if (sink.overflow) { status = ST_ARENA; n_ops = 0; }
    }
    __syncwarp();
    const uint32_t max_ops = __reduce_max_sync(FULL, n_ops);
    group_emit(P, have, pair, status, res, n_ops, scratch, 32u,
               (unsigned long long)(slot_words - top + scratch_w) * 4ull + 256ull * max_ops, c_cells, c_written, c_steps);
}

__global__ void __launch_bounds__(32 * WFA_LANE_WARPS)
lane_kernel(const KParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wib = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    unsigned char *smem = smem_raw + (size_t)wib * lane_smem_bytes(P.dM, P.dE, P.ring_cap);
    const uint64_t worker = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    uint8_t *slot = P.arena + worker * P.slot_bytes;
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = (uint32_t)atomicAdd(&P.ctr->work_next, 32ull);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= P.n_work) break;
        const bool have = first + lane < P.n_work;
        const uint32_t pair = have ? P.work[first + lane] : 0u;
        lane_group(P, have, pair, smem, slot, P.slot_bytes);
    }
}

} /* namespace wfak */
