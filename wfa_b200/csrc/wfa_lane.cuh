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
 *   seqQ, seqT   u32 [SW][32]          2-bit packed sequences (<= 254 bases; SW = words + 1)
 *   ringM        u8  [max(x,o+e)/g][W][32]   offsets of the scores of M still needed
 *   ringI, ringD u8  [e/g][W][32]            the same for I and D
 * Offsets fit a byte (<= m+1 <= 255); the 3-bit provenance codes are not needed by `next`.
 * Diagonal k lives at column k + W/2 of every row.  A row is replaced in place by the row
 * max(x,o+e) (resp. e) scores later, and everything outside a row's written range is kept
 * "absent" (0), so the diagonal loop reads its five sources without any range test.
 *
 * Backtrace arena (HBM, one slot per warp, reused group after group):
 *   (row headers {lo, hi, first word} per score index, shared by the 32 pairs, stay in shared memory)
 *   cell  u32 [aw][32] per score: M | I<<8 | D<<16 (offsets)
 * i.e. 4 bytes per (score, diagonal, pair) instead of the 12 of three raw words.  The 3-bit
 * provenance codes are not stored: `next` picks them as a function of the five source offsets,
 * which are all in the arena, so `LaneView::get_typed` re-derives the code of the ~20 cells the
 * backtrace stands on (same validity tests, same tie order) instead of the forward pass
 * computing and packing it for all ~530 cells of a pair.  The backtrace itself is the shared
 * literal one (back_trace in wfa_kernels.cuh).
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

constexpr int LANE_HDR_ROWS = 64;            /* scores (in units of g) a group may reach; beyond: WARP worker */
constexpr int LANE_FINISH_WARPS = 4;
constexpr uint32_t LANE_SOPS = 32;           /* ops per pair kept in shared memory by the finish kernel (16 bits each); more go to the slot */

/* Op sink of the finish kernel: the first LANE_SOPS run-merged ops of a pair stay in shared
 * memory as letter index << 13 | count (counts <= n + m < 8192 in the LANE class), the rest go
 * to the group's slot like OpSink's. */
struct LaneSink {
    uint16_t *sbuf;                /* shared memory, [op][lane], already offset by the lane */
    uint64_t *buf; uint32_t cap, n; uint32_t cur_op, cur_n; bool overflow;
    __device__ __forceinline__ void add(uint32_t op, uint32_t cnt)
    {
        if (op == cur_op) { cur_n += cnt; return; }
        flush();
        cur_op = op; cur_n = cnt;
    }
    __device__ __forceinline__ void flush()
    {
        if (cur_op == 0) return;
        if (n < LANE_SOPS) {
            /* 'D' 44, 'H' 48, 'I' 49, 'M' 4d, 'X' 58 -> 0..4 */
            const uint32_t c = cur_op == 'M' ? 3u : (cur_op == 'X' ? 4u : (cur_op == 'I' ? 2u : (cur_op == 'H' ? 1u : 0u)));
            sbuf[n * 32u] = (uint16_t)(c << 13 | cur_n);
        } else if (n - LANE_SOPS < cap) buf[(size_t)(n - LANE_SOPS) * 32u] = (uint64_t)cur_op << 32 | cur_n;
        else overflow = true;
        n++;
        cur_op = 0;
    }
};
struct LaneOps {
    const uint16_t *sbuf; const uint64_t *buf;
    __device__ __forceinline__ uint64_t operator()(uint32_t j) const
    {
        if (j < LANE_SOPS) {
            const uint32_t w = sbuf[j * 32u];
            const uint32_t letter = __byte_perm(0x4d494844u, 0x00000058u, w >> 13) & 0xffu;
            return (uint64_t)letter << 32 | (w & 0x1fffu);
        }
        return buf[(size_t)(j - LANE_SOPS) * 32u];
    }
};

/* dM, dE as in KParams (max(x,o+e)/g+1, e/g+1); W ring columns; SW words per sequence */
__host__ __device__ inline size_t lane_smem_bytes(int dM, int dE, int W, int SW)
{
    size_t b = 2 * (size_t)SW * 128;                                   /* seqQ, seqT */
    b += ((size_t)(dM - 1) + 2 * (size_t)(dE - 1)) * (size_t)W * 32;   /* rings (in place: one row less than the WARP worker) */
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
template <int OFF> __device__ __forceinline__ uint32_t lds_u8o(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ void sts_u8o(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0+%1], %2;" :: "r"(addr), "n"(OFF), "r"(v) : "memory");
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

/* One diagonal of `next` (wfa.go:572-699) on bare offsets (0 = absent); same candidate packing
 * as next_cell: value<<p | priority, one max() picks offset and provenance.
 *   um = m (offset <= m, :581,:585,:651), ubk = n + k (offset - k <= n, :616,:620,:651)
 * Used by the backtrace to re-derive the provenance codes the forward pass does not store. */
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

/* The same recurrence, offsets only (forward pass): I = max(valid sources) + 1, D = max(valid
 * sources), M = max(I, D, valid M[s-x][k] + 1).  Which source won -- the provenance code --
 * is a function of the same five stored offsets and is re-derived by the backtrace for the
 * few cells it visits (LaneView::get_typed) instead of being computed for every cell here. */
struct Cell3O { uint32_t M, I, D; };
__device__ __forceinline__ Cell3O next_off3(uint32_t mo_l, uint32_t ie_l, uint32_t mo_r, uint32_t de_r, uint32_t mx,
                                            uint32_t um, uint32_t ubk)
{
    Cell3O r;
    /* a source is valid iff 1 <= offset <= bound; an absent source (0) may pass the test, it selects 0 either way */
    const uint32_t a = mo_l <= um ? mo_l : 0u, b = ie_l <= um ? ie_l : 0u;
    const uint32_t mi = max(a, b);
    r.I = mi + (mi != 0u);
    const uint32_t c = mo_r <= ubk ? mo_r : 0u, d = de_r <= ubk ? de_r : 0u;
    r.D = max(c, d);
    const uint32_t e = (mx - 1u) < min(um, ubk) ? mx + 1u : 0u;
    r.M = max(max(e, r.I), r.D);
    return r;
}

/* 16 bases starting at base `pos` of a lane's sequence in shared memory (base pos in the low bits) */
__device__ __forceinline__ uint32_t lane_chunk(uint32_t seq_sa, int pos)
{
    const uint32_t a = seq_sa + ((uint32_t)pos >> 4) * 128u;
    return __funnelshift_r(lds_u32(a), lds_u32(a + 128u), (uint32_t)pos * 2u);     /* the shift count wraps at 32: (pos % 16) * 2 */
}

/* extend (wfa.go:394-455) of one M offset on diagonal k: returns the extended offset */
__device__ __forceinline__ uint32_t lane_extend(uint32_t sQ, uint32_t sT, uint32_t M, int k, int n, int m)
{
    const int h = (int)M, v = h - k;
    if (M == 0 || v <= 0 || v >= n || h >= m) return M;
    const int ext = min(n - v, m - h);
    int l = 0;
    while (l < ext) {
        const uint32_t xx = lane_chunk(sQ, v + l) ^ lane_chunk(sT, h + l);
        if (xx) { l += (__ffs((int)xx) - 1) >> 1; break; }
        l += 16;
    }
    return M + (uint32_t)min(l, ext);
}

/* One stage of the forward pass of one group of up to 32 pairs: the rows of score indices
 * [si0, si1] go to the group's slot of this stage.  A pair that reaches the end cell leaves its
 * record for the finish kernel; a pair still running at the end of the stage saves its rings and
 * is queued for the next stage, where it is grouped with other pairs that are still running --
 * lanes of a warp work in lockstep, so a group costs as many rows as its slowest pair needs, and
 * regrouping the survivors keeps finished lanes from idling through the long tail. */
constexpr uint32_t LANE_REC_WORDS = 12;      /* 0 status | first_eq << 8 | final score index << 16 | stages << 24, 1 final score,
                                              * 2 C, 3 cells written, 4 score steps, 5.. group << 5 | lane of every stage */

__device__ __forceinline__ void lane_stage(const KParams &P, const int stg, const bool have, const uint32_t ticket, const uint32_t pair,
                                           unsigned char *smem, uint32_t *cells, const uint32_t group)
{
    const uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    /* rings are written in place: the row of score s replaces the oldest row still needed,
     * M[s - max(x,o+e)] resp. I/D[s - e]; sources ahead of the write position are read first and
     * the ones behind it are carried in registers.  A row's range contains the range of the row
     * it replaces (the geometry only grows), so what lies outside a row's range stays 0 = absent. */
    const int RM = P.dM - 1, RE = P.dE - 1, W = P.ring_cap, KC = W >> 1, SW = P.group;
    const int xg = P.xg, oeg = P.oeg, x = (int)P.x;
    const LaneGeom &G = P.lg;

    unsigned char *p = smem;
    const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(p) + (uint32_t)lane * 4u;
    const uint32_t sT = sQ + (uint32_t)SW * 128u;
    const uint32_t rowB = (uint32_t)W * 32u;
    const uint32_t ringL = sQ - (uint32_t)lane * 4u + 2u * (uint32_t)SW * 128u + (uint32_t)lane;          /* this lane's byte of column 0, ring row 0 */
    const uint32_t rM = ringL + (uint32_t)KC * 32u;                                                        /* diagonal 0 */
    const uint32_t rI = rM + (uint32_t)RM * rowB, rD = rI + (uint32_t)RE * rowB;

    int status = have ? ST_OK : ST_PENDING;
    PairDesc pd; pd.q_byte = pd.t_byte = pd.q_word = pd.t_word = 0; pd.n = pd.m = 0;
    if (have) {
        pd = P.pairs[pair];
        if (stg == 0 && (P.pflags[pair] & 1)) status = ST_NEED8;
    }
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const bool act = status == ST_OK;

    /* sequences -> lane-private shared memory columns; rings start all "absent" */
    {
        const uint32_t *gq = P.packed + pd.q_word, *gt = P.packed + pd.t_word;
        const int wq = act ? (n + 15) >> 4 : 0, wt = act ? (m + 15) >> 4 : 0;
#pragma unroll 1
        for (int w0 = 0; w0 < SW; w0 += 4) {
            uint32_t a[4], c[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { a[j] = w0 + j < wq ? __ldg(gq + w0 + j) : 0u; c[j] = w0 + j < wt ? __ldg(gt + w0 + j) : 0u; }
#pragma unroll
            for (int j = 0; j < 4; j++) if (w0 + j < SW) { sts_u32(sQ + (uint32_t)(w0 + j) * 128u, a[j]); sts_u32(sT + (uint32_t)(w0 + j) * 128u, c[j]); }
        }
        const uint32_t ring0 = sQ + 2u * (uint32_t)SW * 128u;           /* rings are a multiple of 128 bytes per 4 columns */
        const int ring_words = (RM + 2 * RE) * W * 8;
#pragma unroll 1
        for (int w = 0; w < ring_words; w += 32) sts_u32(ring0 + (uint32_t)w * 4u, 0u);
    }
    __syncwarp();
    const bool first_eq = ((lds_u32(sQ) ^ lds_u32(sT)) & 3u) == 0u;       /* q[0] == t[0], wfa.go:155-158 */
    /* a lane works on diagonals [klo, klo + kspan] = [-(n-1), m-1] (wfa.go:562-563); an idle or
     * finished lane gets an empty range and computes nothing but "absent" */
    const int klo = -(n - 1);
    const uint32_t kspan = (uint32_t)(n + m - 2);
    /* rows inside [clamp_lo, clamp_hi] need no per-pair clamp.  Idle and finished lanes are not
     * masked at all: what they compute is never read (their counters, end test and results are
     * guarded, their arena column past the final score is not visited by the backtrace). */
    const int clamp_lo = -(__reduce_min_sync(FULL, act ? n : INT_MAX) - 1), clamp_hi = __reduce_min_sync(FULL, act ? m : INT_MAX) - 1;

    const int si0 = stg == 0 ? 0 : G.stage_end[stg - 1] + 1, si1 = G.stage_end[stg];
    /* saved state of a pair: ring_rows x 16 words of 4 columns each, in the frame of a 64-column
     * ring (diagonal k in column k + 32) whatever this stage's own ring width is -- a stage whose
     * rows are narrow runs with fewer columns (more warps per SM) and shifts by wshift words */
    const int ring_rows = RM + 2 * RE, wpr = W >> 2, wshift = (32 - KC) >> 2;
    bool done = false; uint32_t minS = 0; int my_si = 0;
    uint32_t c_cells = 0, c_written = 0, c_steps = 0;
    /* ring columns (in words of 4) that hold anything after row sb: those of the last row that exists up to sb */
    auto ring_span = [&](int sb, int &c0, int &c1) {
        while (sb > 0 && G.lo[sb] > G.hi[sb]) sb--;
        c0 = (G.lo[sb] + 32) >> 2; c1 = (G.hi[sb] + 32) >> 2;
    };
    if (stg > 0 && act) {
        /* rings as the previous stage left them; only the columns its last row could reach.  A
         * ring row is wpr = 16 words of four columns: its four 16-byte quarters are requested
         * together (each lane reads its own pair's record, so every load is a DRAM round trip) */
        const uint4 *st = reinterpret_cast<const uint4 *>(P.la.state + (size_t)ticket * G.state_words);
        int cw0, cw1; ring_span(si0 - 1, cw0, cw1);
        const int q0 = cw0 >> 2, q1 = cw1 >> 2, qpr = 4;
#pragma unroll 1
        for (int r = 0; r < ring_rows; r++) {
            uint4 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = (q >= q0 && q <= q1) ? st[r * qpr + q] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                for (int j = 0; j < 4; j++) if (w4[j] && (uint32_t)(q * 4 + j - wshift) < (uint32_t)wpr) {
                    const uint32_t a = ringL + (uint32_t)r * rowB + (uint32_t)(q * 4 + j - wshift) * 128u;
                    sts_u8o<0>(a, w4[j] & 255u); sts_u8o<32>(a, (w4[j] >> 8) & 255u); sts_u8o<64>(a, (w4[j] >> 16) & 255u); sts_u8o<96>(a, w4[j] >> 24);
                }
            }
        }
        const uint4 cn = st[ring_rows * qpr];
        c_cells = cn.x; c_written = cn.y; c_steps = cn.z;
    }

    uint32_t s = (uint32_t)si0 * P.g; int si = si0, cur = si0 % RM, curE = si0 % RE;
    if (__any_sync(FULL, act)) for (;;) {
        const int lo = G.lo[si], hi = G.hi[si];
        if (lo <= hi) {
            int slX = cur - xg, slO = cur - oeg;
            slX += slX < 0 ? RM : 0; slO += slO < 0 ? RM : 0;
            const bool has_init = (s == 0) || (s == (uint32_t)x);          /* global: the one cell k = 0 */
            const int aw = hi - lo + 1;
            const uint32_t off = (uint32_t)G.off[si] * 32u;
            const uint32_t bCM = rM + (uint32_t)cur * rowB, bCI = rI + (uint32_t)curE * rowB, bCD = rD + (uint32_t)curE * rowB;
            /* running byte addresses of cell k in the lane's columns; a row's cell k+1 is 32 bytes on */
            uint32_t pO = rM + (uint32_t)slO * rowB + (uint32_t)(lo * 32), pX = rM + (uint32_t)slX * rowB + (uint32_t)(lo * 32);
            uint32_t pM = bCM + (uint32_t)(lo * 32), pI = bCI + (uint32_t)(lo * 32), pD = bCD + (uint32_t)(lo * 32);
            uint32_t *gp = cells + off + lane;
            keep(pO); keep(pX); keep(pM); keep(pI); keep(pD); keep_ptr(gp);
            const uint32_t um = (uint32_t)m;
            const uint32_t pos_max = (uint32_t)(SW - 1) * 16u - 1u;      /* last base whose 16-base chunk lies inside the SW words of a sequence */

            struct Pend { Cell3O c; int ext; uint32_t xr; };
            /* phase A of one cell: next (wfa.go:572-699) -> first 16-base compare of extend.  CLAMP
             * rows reach beyond some pair's own diagonals [-(n-1), m-1] (wfa.go:562-563). */
            auto cell_a = [&](auto clamp, const int k, const uint32_t mo_l, const uint32_t ie_l, const uint32_t mo_r, const uint32_t de_r, const uint32_t mx) -> Pend {
                Pend q;
                if (decltype(clamp)::value) {
                    const bool inb = (uint32_t)(k - klo) <= kspan;
                    q.c = next_off3(mo_l, ie_l, mo_r, de_r, mx, inb ? um : 0u, inb ? (uint32_t)(n + k) : 0u);
                } else q.c = next_off3(mo_l, ie_l, mo_r, de_r, mx, um, (uint32_t)(n + k));
                /* extend applies iff the cell exists, v > 0, v < n and h < m (wfa.go:404).  A present
                 * cell always has v >= 1: the start cell is (1, 1) and every step of next / extend
                 * keeps or raises v -- so the test is "exists and min(n - v, m - h) > 0". */
                const int h = (int)q.c.M, v = h - k;
                q.ext = q.c.M != 0u ? max(min(n - v, m - h), 0) : 0;
                /* positions are only meaningful when ext > 0; clamped so that the loads stay inside the lane's own
                 * sequence words (an absent cell or one at the end of a sequence compares something and advances by 0) */
                q.xr = lane_chunk(sQ, (int)min((uint32_t)v, pos_max)) ^ lane_chunk(sT, (int)min((uint32_t)h, pos_max));
                return q;
            };
            /* phase B: rest of extend (wfa.go:411-454) */
            auto cell_b = [&](Pend &q, const int k) {
                /* bases matched by the first compare: index of the lowest differing bit pair, 16 when
                 * there is none (brev(0) = 0 has 32 leading zeros).  ext = 0 adds nothing. */
                int l = matched_bases(q.xr);
                if (l == 16 && q.ext > 16) {
                    const int h = (int)q.c.M, v = h - k;
                    while (l < q.ext) {
                        const uint32_t xx = lane_chunk(sQ, v + l) ^ lane_chunk(sT, h + l);
                        if (xx) { l += matched_bases(xx); break; }
                        l += 16;
                    }
                }
                q.c.M += (uint32_t)min(l, q.ext);
            };
            auto row = [&](auto clamp) {
                int k = lo;
                uint32_t o_m1 = lds_u8o<-32>(pO), o_0 = lds_u8o<0>(pO), i_m1 = lds_u8o<-32>(pI);
                auto word_of = [](const Pend &q) { return q.c.M | q.c.I << 8 | q.c.D << 16; };
                /* four diagonals per pass: their dependency chains interleave (the warps per SM
                 * are few -- shared memory -- so the instruction-level parallelism has to come from here) */
                for (; k + 3 <= hi; k += 4) {
                    const uint32_t o_p1 = lds_u8o<32>(pO), o_p2 = lds_u8o<64>(pO), o_p3 = lds_u8o<96>(pO), o_p4 = lds_u8o<128>(pO);
                    const uint32_t i_0 = lds_u8o<0>(pI), i_p1 = lds_u8o<32>(pI), i_p2 = lds_u8o<64>(pI), i_p3 = lds_u8o<96>(pI);
                    const uint32_t d_p1 = lds_u8o<32>(pD), d_p2 = lds_u8o<64>(pD), d_p3 = lds_u8o<96>(pD), d_p4 = lds_u8o<128>(pD);
                    const uint32_t x_0 = lds_u8o<0>(pX), x_p1 = lds_u8o<32>(pX), x_p2 = lds_u8o<64>(pX), x_p3 = lds_u8o<96>(pX);
                    Pend q0 = cell_a(clamp, k, o_m1, i_m1, o_p1, d_p1, x_0);
                    Pend q1 = cell_a(clamp, k + 1, o_0, i_0, o_p2, d_p2, x_p1);
                    Pend q2 = cell_a(clamp, k + 2, o_p1, i_p1, o_p3, d_p3, x_p2);
                    Pend q3 = cell_a(clamp, k + 3, o_p2, i_p2, o_p4, d_p4, x_p3);
                    cell_b(q0, k); cell_b(q1, k + 1); cell_b(q2, k + 2); cell_b(q3, k + 3);
                    sts_u8o<0>(pM, q0.c.M); sts_u8o<0>(pI, q0.c.I); sts_u8o<0>(pD, q0.c.D);
                    sts_u8o<32>(pM, q1.c.M); sts_u8o<32>(pI, q1.c.I); sts_u8o<32>(pD, q1.c.D);
                    sts_u8o<64>(pM, q2.c.M); sts_u8o<64>(pI, q2.c.I); sts_u8o<64>(pD, q2.c.D);
                    sts_u8o<96>(pM, q3.c.M); sts_u8o<96>(pI, q3.c.I); sts_u8o<96>(pD, q3.c.D);
                    gp[0] = word_of(q0); gp[32] = word_of(q1); gp[64] = word_of(q2); gp[96] = word_of(q3);
                    o_m1 = o_p3; o_0 = o_p4; i_m1 = i_p3;
                    pO += 128; pX += 128; pM += 128; pI += 128; pD += 128; gp += 128;
                }
                for (; k <= hi; k++) {
                    const uint32_t o_p1 = lds_u8o<32>(pO), i_0 = lds_u8o<0>(pI), d_p1 = lds_u8o<32>(pD), x_0 = lds_u8o<0>(pX);
                    Pend q0 = cell_a(clamp, k, o_m1, i_m1, o_p1, d_p1, x_0);
                    cell_b(q0, k);
                    sts_u8o<0>(pM, q0.c.M); sts_u8o<0>(pI, q0.c.I); sts_u8o<0>(pD, q0.c.D);
                    gp[0] = word_of(q0);
                    o_m1 = o_0; o_0 = o_p1; i_m1 = i_0;
                    pO += 32; pX += 32; pM += 32; pI += 32; pD += 32; gp += 32;
                }
            };
            if (lo < clamp_lo || hi > clamp_hi) row(std::true_type{}); else row(std::false_type{});
            if (has_init && act && !done && (first_eq ? (s == 0) : (s == (uint32_t)x)) && lds_u8(bCM) == 0u) {
                /* initComponents (wfa.go:155-158): cell k = 0, unless next's Set already wrote it
                 * (wfa_wavefront.go:93); then its extend (the row loop saw an absent cell) */
                const uint32_t M = lane_extend(sQ, sT, 1u, 0, n, m);
                sts_u8(bCM, M);
                uint32_t *g0 = cells + off + lane - lo * 32;
                *g0 = (*g0 & 0xffffff00u) | M;
            }
            if (act && !done) {
                /* M WaveFront.Lo/Hi of this pair = first and last present cell (only the work counter C needs them) */
                int a = lo, b = hi;
                while (a <= hi && lds_u8(bCM + (uint32_t)(a * 32)) == 0u) a++;
                while (b > a && lds_u8(bCM + (uint32_t)(b * 32)) == 0u) b--;
                if (a <= hi) {
                    c_steps++; c_cells += (uint32_t)(b - a + 1); c_written += (uint32_t)aw;
                    /* end test on diagonal m-n (wfa.go:235-239); the lane reads back its own column */
                    if (Ak >= lo && Ak <= hi) {
                        const uint32_t hM = lds_u8(bCM + (uint32_t)(Ak * 32));
                        if ((int)hM >= m) { done = true; minS = s; my_si = si; }
                    }
                }
            }
        }
        if (si == si1 || __all_sync(FULL, !act || done)) break;
        s += P.g; si++;
        cur = cur + 1 == RM ? 0 : cur + 1;
        curE = curE + 1 == RE ? 0 : curE + 1;
    }
    const bool last_stage = stg + 1 >= G.n_stages;
    bool cont = act && !done;                                              /* still running at the end of the stage */
    if (cont && last_stage) { status = ST_RING; cont = false; }          /* beyond the byte rings' reach: WARP worker */

    if (!last_stage) {
        const unsigned cm = __ballot_sync(FULL, cont);
        if (cm) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.ctr->lane_count[stg + 1], (unsigned long long)__popc(cm));
            base = __shfl_sync(FULL, base, 0);
            if (cont) {
                const unsigned long long pos = base + (unsigned)__popc(cm & ((1u << lane) - 1u));
                if (pos < (unsigned long long)P.la.cap[stg + 1] * 32ull) {
                    P.la.list[stg + 1][pos] = ticket;
                    uint4 *st = reinterpret_cast<uint4 *>(P.la.state + (size_t)ticket * G.state_words);
                    int cw0, cw1; ring_span(si1, cw0, cw1);
                    const int q0 = cw0 >> 2, q1 = cw1 >> 2, qpr = 4;
#pragma unroll 1
                    for (int r = 0; r < ring_rows; r++)
#pragma unroll 1
                        for (int q = q0; q <= q1; q++) {
                            uint32_t w4[4];
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                w4[j] = 0u;
                                if ((uint32_t)(q * 4 + j - wshift) < (uint32_t)wpr) {
                                    const uint32_t a = ringL + (uint32_t)r * rowB + (uint32_t)(q * 4 + j - wshift) * 128u;
                                    w4[j] = lds_u8o<0>(a) | lds_u8o<32>(a) << 8 | lds_u8o<64>(a) << 16 | lds_u8o<96>(a) << 24;
                                }
                            }
                            st[r * qpr + q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    st[ring_rows * qpr] = make_uint4(c_cells, c_written, c_steps, 0u);
                } else { status = ST_ARENA; cont = false; }                /* the next stage's arena is full: the host re-queues the pair */
            }
        }
    }
    if (have) {
        uint32_t *rec = P.la.rec + (size_t)ticket * LANE_REC_WORDS;
        rec[5 + stg] = group << 5 | (uint32_t)lane;
        if (!cont) {
            rec[0] = (uint32_t)status | (first_eq ? 0x100u : 0u) | (uint32_t)my_si << 16 | (uint32_t)(stg + 1) << 24;
            rec[1] = minS; rec[2] = c_cells; rec[3] = c_written; rec[4] = c_steps;
        }
    }
}

/* Component.Get on the staged group arenas.  The rows of score index si of every group of a
 * stage have the same geometry (LaneGeom, copied to shared memory); the pair's own column in
 * each stage is `seg[stage]` = its group's slot + its lane.  Like before, the view remembers
 * the five source words it fetched for the cell the backtrace stands on: the next cell of the
 * walk and the offsets the reference re-derives there (wfa.go:766-817) are always among them,
 * so one step of the backtrace costs one round of five independent global loads. */
struct LaneGeomS { int8_t lo[64], hi[64]; uint16_t off[64]; };
struct LaneView {
    const LaneGeomS *G;        /* shared memory */
    const uint32_t *const *seg;/* shared memory: seg[stage * 32] = this lane's column in the slot of its group of that stage */
    int             e0, e1, e2;/* last row of stages 0..2 (INT_MAX where there is no later stage) */
    int             si_last;
    int             n, m, xg, oeg, eg;
    bool            first_eq;
    int             c_si, c_k;                 /* cell whose sources are cached (c_si < 0: none) */
    uint32_t        c_w[5];                    /* words at (si-oeg,k-1) (si-eg,k-1) (si-oeg,k+1) (si-eg,k+1) (si-xg,k) */
    __device__ __forceinline__ uint32_t word(int si, int k) const
    {
        if (si < 0 || si > si_last) return 0;
        const int lo = G->lo[si], hi = G->hi[si];
        if (k < lo || k > hi) return 0;
        const int j = (si > e0) + (si > e1) + (si > e2);
        return seg[j * 32][((uint32_t)G->off[si] + (uint32_t)(k - lo)) * 32u];
    }
    __device__ __forceinline__ uint32_t cached_word(int si, int k) const
    {
        const int dk = k - c_k, ds = c_si - si;
        if (c_si >= 0) {
            if (dk == -1) { if (ds == oeg) return c_w[0]; if (ds == eg) return c_w[1]; }
            else if (dk == 1) { if (ds == oeg) return c_w[2]; if (ds == eg) return c_w[3]; }
            else if (dk == 0 && ds == xg) return c_w[4];
        }
        return word(si, k);
    }
    /* offset << 3, code bits zero: all the backtrace needs of a source cell */
    __device__ __forceinline__ uint32_t get(int comp, int si, int k) const
    {
        return ((cached_word(si, k) >> (8 * comp)) & 255u) << T_BITS;
    }
    /* raw word of the cell the backtrace stands on: offset << 3 | provenance code, the code
     * re-derived from the cell's five sources exactly as `next` chose it (wfa.go:579-698), or
     * the init code when `next` wrote nothing there (wfa.go:155-158) */
    __device__ __forceinline__ uint32_t get_typed(int comp, int si, int k)
    {
        const uint32_t o = (cached_word(si, k) >> (8 * comp)) & 255u;
        if (o == 0) return 0;
        const uint32_t wl = word(si - oeg, k - 1), el = word(si - eg, k - 1);
        const uint32_t wr = word(si - oeg, k + 1), er = word(si - eg, k + 1);
        const uint32_t wx = word(si - xg, k);
        c_si = si; c_k = k; c_w[0] = wl; c_w[1] = el; c_w[2] = wr; c_w[3] = er; c_w[4] = wx;
        const CellO c = next_off(wl & 255u, (el >> 8) & 255u, wr & 255u, (er >> 16) & 255u, wx & 255u,
                                 (uint32_t)m, (uint32_t)(n + k));
        uint32_t code;
        if (comp == 1) code = T_INS_OPEN + ((c.code >> 3) & 1u);
        else if (comp == 2) code = T_DEL_OPEN + ((c.code >> 4) & 1u);
        else code = c.M ? (c.code & 7u) : (first_eq ? T_MATCH : T_MISMATCH);
        return o << T_BITS | code;
    }
};

/* Backtraces (wfa.go:703-983) of 32 pairs, lane-parallel, then their results. */
__device__ __forceinline__ void lane_finish(const KParams &P, const bool have, const uint32_t ticket, const uint32_t pair,
                                            const LaneGeomS *Gs, const uint32_t **seg, uint16_t *sops, const bool sample, WorkAcc *acc)
{
    const int lane = threadIdx.x & 31;
    const LaneGeom &G = P.lg;
    int status = ST_PENDING; bool first_eq = false; int my_si = 0, nseg = 0;
    uint32_t minS = 0, c_cells = 0, c_written = 0, c_steps = 0;
    uint32_t *slot_last = nullptr; uint32_t lane_last = 0;
    if (have) {
        const uint32_t *rec = P.la.rec + (size_t)ticket * LANE_REC_WORDS;
        const uint32_t r0 = rec[0];
        status = (int)(r0 & 255u); first_eq = (r0 & 0x100u) != 0u; my_si = (int)((r0 >> 16) & 255u); nseg = (int)(r0 >> 24);
        minS = rec[1]; c_cells = rec[2]; c_written = rec[3]; c_steps = rec[4];
        for (int j = 0; j < nseg; j++) {
            const uint32_t sg = rec[5 + j];
            uint32_t *slot = reinterpret_cast<uint32_t *>(P.la.arena[j]) + (size_t)(sg >> 5) * G.slot_words[j];
            seg[j * 32 + lane] = slot + (sg & 31u);
            slot_last = slot; lane_last = sg & 31u;
        }
    }
    int n = 0, m = 0;
    if (status == ST_OK) { const PairDesc pd = P.pairs[pair]; n = (int)pd.n; m = (int)pd.m; }
    const int Ak = m - n;

    Result res;
    res.score = 0; res.tbegin = res.tend = res.qbegin = res.qend = 0;
    res.align_len = res.matches = res.gaps = res.gap_regions = 0; res.n_ops = 0;
    res.status = (uint8_t)status; res.pad_[0] = res.pad_[1] = res.pad_[2] = 0;
    uint32_t n_ops = 0;
    /* ops beyond the shared-memory tier: the pair's column of the op scratch of its last slot */
    uint64_t *scratch = reinterpret_cast<uint64_t *>(slot_last) + lane_last;
    __syncwarp();
    if (status == ST_OK) {
        LaneView A; A.G = Gs; A.seg = seg + lane; A.si_last = my_si; A.c_si = -1; A.c_k = 0;
        A.e0 = G.n_stages > 1 ? G.stage_end[0] : INT_MAX; A.e1 = G.n_stages > 2 ? G.stage_end[1] : INT_MAX; A.e2 = G.n_stages > 3 ? G.stage_end[2] : INT_MAX;
        A.c_w[0] = A.c_w[1] = A.c_w[2] = A.c_w[3] = A.c_w[4] = 0;
        A.n = n; A.m = m; A.xg = P.xg; A.oeg = P.oeg; A.eg = P.eg; A.first_eq = first_eq;
        LaneSink sink; sink.sbuf = sops + lane; sink.buf = scratch;
        sink.cap = G.scratch_words / 64u;
        sink.n = 0; sink.cur_op = 0; sink.cur_n = 0; sink.overflow = false;
        back_trace_inl(A, P, n, m, minS, Ak, res, sink);
        n_ops = sink.n;
        if (sink.overflow) { status = ST_ARENA; n_ops = 0; }
        else if (sample) atomicAdd(&P.ctr->lane_hist[my_si & 63], 1u);
    }
    __syncwarp();
    group_emit(P, have, pair, status, res, n_ops, LaneOps{sops + lane, scratch},
               (unsigned long long)G.slot_words[0] * 4ull, c_cells, c_written, c_steps, acc);
}

__global__ void __launch_bounds__(32 * WFA_LANE_WARPS)
lane_kernel(const KParams P, const int stg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wib = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    unsigned char *smem = smem_raw + (size_t)wib * lane_smem_bytes(P.dM, P.dE, P.ring_cap, P.group);
    if (stg == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(&P.ctr->t_first, t); }
    /* pairs entering this stage: the whole work list, or what the previous stage queued */
    uint32_t n_in = P.n_work;
    if (stg > 0) n_in = (uint32_t)min(P.ctr->lane_count[stg], (unsigned long long)P.la.cap[stg] * 32ull);
    for (;;) {
        uint32_t first = 0;
        if (lane == 0) first = (uint32_t)atomicAdd(&P.ctr->lane_count[4 + stg], 32ull);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n_in) break;
        const bool have = first + lane < n_in;
        const uint32_t ticket = have ? (stg == 0 ? first + lane : P.la.list[stg][first + lane]) : 0u;
        const uint32_t pair = have ? (P.work ? P.work[ticket] : P.pair_base + ticket) : 0u;
        const uint32_t group = first >> 5;                                   /* one slot per group and stage */
        lane_stage(P, stg, have, ticket, pair, smem, reinterpret_cast<uint32_t *>(P.la.arena[stg]) + (size_t)group * P.lg.slot_words[stg], group);
    }
    if (lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(&P.ctr->t_last, t); }
}

/* One warp per 32 pairs, many warps per SM: the backtrace is a chain of dependent arena reads, so
 * it runs here at high occupancy instead of inside the shared-memory-limited forward kernel. */
__global__ void __launch_bounds__(32 * LANE_FINISH_WARPS, 8)
lane_finish_kernel(const KParams P)
{
    __shared__ LaneGeomS geom_s;
    __shared__ const uint32_t *seg_s[LANE_FINISH_WARPS][LANE_MAX_STAGES * 32];
    __shared__ uint16_t ops_s[LANE_FINISH_WARPS][LANE_SOPS * 32];
    const int wib = (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { geom_s.lo[i] = P.lg.lo[i]; geom_s.hi[i] = P.lg.hi[i]; geom_s.off[i] = P.lg.off[i]; }
    __syncthreads();
    const uint32_t n_groups = (P.n_work + 31u) >> 5;
    WorkAcc acc; acc.cells = acc.written = acc.steps = acc.ops = acc.arena_max = 0;
    for (uint32_t g = blockIdx.x * LANE_FINISH_WARPS + wib; g < n_groups; g += gridDim.x * LANE_FINISH_WARPS) {
        const uint32_t ticket = (g << 5) + lane;
        const bool have = ticket < P.n_work;
        const uint32_t pair = have ? (P.work ? P.work[ticket] : P.pair_base + ticket) : 0u;
        __syncwarp();
        lane_finish(P, have, ticket, pair, &geom_s, seg_s[wib], ops_s[wib], (g & 15u) == 0u, &acc);
    }
    if (lane == 0) {
        atomicAdd(&P.ctr->cells, acc.cells); atomicAdd(&P.ctr->cells_written, acc.written);
        atomicAdd(&P.ctr->steps, acc.steps); atomicAdd(&P.ctr->ops, acc.ops);
        if (acc.arena_max) atomicMax(&P.ctr->arena_used_max, acc.arena_max);
    }
    if (lane == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMax(&P.ctr->t_last, t); }
}

} /* namespace wfak */
