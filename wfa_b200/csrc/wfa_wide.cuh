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
 *     arriving, so all CTAs take the same decision without a second round;
 *   - a thread handles two neighbouring diagonals per step (one 32-bit shared-memory word per row),
 *     only offsets are computed (codes re-derived by the backtrace, as in the LANE / SLIM classes),
 *     both sequences are read through shared-memory windows (one LDS.64 + one funnel shift per
 *     16-base compare), and the arena gets ONE 64-bit word per cell (M | I << 21 | D << 42, the
 *     SLIM worker's wide cell): 8 instead of 12 bytes per cell, written once with 128-bit stores,
 *     never read by the forward pass;
 *   - Lo/Hi, the end test (wfa.go:235-239) and the semi-global start-cell test (wfa.go:270-375,
 *     early-stop form) are taken on the way; there is no second pass over a row.
 * The backtraces run afterwards in their own kernel (wide_finish_kernel: lane-parallel over the
 * pairs of the launch, SlimView on the slots the forward pass left), so no cluster idles while one
 * thread chases pointers.
 *
 * Semantics follow the reference at /root/reference (cited as wfa.go:LINE); the recurrences are
 * next_off3 / next_off of wfa_lane.cuh.
 */
#pragma once
#include "wfa_slim.cuh"

namespace wfak {

constexpr uint32_t WIDE_MAX_M = 65534;            /* offsets up to m + 1 must fit 16 bits */
constexpr uint32_t WIDE_HEAD_BYTES = 1280;        /* mailboxes 512 + reduction scratch 640 + work item 16, padded */
constexpr int WIDE_MAX_CLUSTER = 8;

/* shared memory of one CTA: head, 9 ring rows of seg + 4 16-bit columns (two halo columns on either
 * side keep the rows word aligned), the two sequence windows (8 bytes per 16 bases + 1 entry each) */
__host__ __device__ inline size_t wide_smem_bytes(uint32_t seg, uint32_t seq_entries)
{
    return WIDE_HEAD_BYTES + 9 * (size_t)(seg + 4) * 2 + 8 + (size_t)seq_entries * 8;
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
 * Every thread of every CTA keeps the same row bookkeeping in registers (the reductions are combined
 * identically everywhere).  Returns what the finish kernel needs (valid in all threads). */
template <bool SEMI>
__device__ __forceinline__ FwdOut forward_wide(const KParams &P, const uint32_t pair, const uint32_t sbase, uint8_t *slot, const uint64_t slot_bytes,
                                               const uint32_t rank, const uint32_t C)
{
    typedef SlimCell<1> SC;
    typedef uint64_t CellT;
    constexpr uint32_t HDR_CELLS = sizeof(SlimHdr) / sizeof(CellT);
    const uint32_t tid = threadIdx.x, T = blockDim.x, lane = tid & 31u, wid = tid >> 5, nw = T >> 5;
    const uint32_t SEG = (uint32_t)P.wide_seg, RS = (SEG + 4u) * 2u, HALF = SEG >> 1;
    const uint32_t sMail = sbase, sRed = sbase + 512u;
    const uint32_t sRing = sbase + WIDE_HEAD_BYTES;
    const uint32_t sSeq = (sRing + 9u * RS + 7u) & ~7u;

    const PairDesc pd = P.pairs[pair];
    const int n = (int)pd.n, m = (int)pd.m, Ak = m - n;
    const uint32_t W = (uint32_t)(n + m - 1);                  /* columns: j = k + n - 1 in [0, W) */
    const uint32_t qent = (pd.n + 15u) >> 4, tent = (pd.m + 15u) >> 4;

    FwdOut f;
    f.status = ST_OK; f.minS = 0; f.lastK = Ak; f.si = 0; f.n = n; f.m = m; f.top = 0;
    f.c_cells = f.c_written = f.c_steps = 0; f.first_eq = false;
    if (pd.m > WIDE_MAX_M || W > C * SEG || qent + tent + 2u > P.wide_seq_cap) { f.status = ST_RING; return f; }

    /* ---- per-pair set-up: ring zeroed (absent everywhere), the sequences in their windows */
    for (uint32_t a = tid * 4u; a < 9u * RS; a += T * 4u) sts_u32(sRing + a, 0u);
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
    cluster_sync_all();                                        /* nobody pushes a halo cell into a ring that is still being cleared */

    SlimHdr *hdrs = reinterpret_cast<SlimHdr *>(slot);         /* grows up, index s/g */
    CellT   *cells = reinterpret_cast<CellT *>(slot);          /* rows grow down from the end */
    const uint32_t slot_cells = (uint32_t)min(slot_bytes / sizeof(CellT), (uint64_t)0xfffffff0u) & ~1u;
    uint32_t top = slot_cells;
    long long room = (long long)slot_cells - (long long)(3 * HDR_CELLS + 8);

    const int nm1 = n - 1;
    const uint32_t base_j = rank * SEG;                        /* first column of this CTA's segment */
    const int ilo = SEMI ? -nm1 : 0, ihi = SEMI ? m - 1 : 0;   /* init cells, wfa.go:155-183 */

    /* present ranges [lo, hi] of the rows s-1 .. s-5 (s-5: what the M slot being overwritten still holds) */
    int lo1 = SLIM_NONE_LO, hi1 = SLIM_NONE_HI, lo2 = SLIM_NONE_LO, hi2 = SLIM_NONE_HI, lo3 = SLIM_NONE_LO, hi3 = SLIM_NONE_HI;
    int lo4 = SLIM_NONE_LO, hi4 = SLIM_NONE_HI, lo5 = SLIM_NONE_LO, hi5 = SLIM_NONE_HI;
    unsigned long long c_cells = 0, c_written = 0, c_steps = 0;
    int status = ST_OK, si = -1, lastK = Ak;
    uint32_t minS = 0, parity = 0;
    uint32_t slotM = 4;                                        /* ring slot of row si (si mod 5), advanced incrementally */

    for (;;) {
        si++;
        slotM = slotM == 4u ? 0u : slotM + 1u;
        const bool has_init = si == 0 || si == SLIM_XG;
        int lo = min(min(lo1, lo2), lo4), hi = max(max(hi1, hi2), hi4);
        if (lo <= hi) { lo = max(lo - 1, -nm1); hi = min(hi + 1, m - 1); }       /* wfa.go:557-563 */
        if (has_init) { lo = min(lo, ilo); hi = max(hi, ihi); }
        lo = min(lo, lo5); hi = max(hi, hi5);                                     /* cells the destination slot still holds are recomputed (to absent) */
        int4 hc = make_int4(0, 1, 0, 0);
        room -= (long long)HDR_CELLS;
        bool exists = false, endhit = false, hit = false;
        int wlo = SLIM_NONE_LO, whi = SLIM_NONE_HI, hitK = Ak;
        if (lo <= hi) {
            const uint32_t ja = (uint32_t)(lo + nm1) & ~1u, jb = (uint32_t)(hi + nm1) | 1u;   /* even / odd: whole column pairs */
            const uint32_t aw = jb - ja + 1u;
            room -= (long long)aw;
            if (room < 0) { status = ST_ARENA; break; }
            const uint32_t off = top - aw;
            /* this CTA's share: column pairs c (columns base_j + 2c, + 1), c in [cl, ch] */
            const int gl = (int)(ja >> 1) - (int)(base_j >> 1), gh = (int)(jb >> 1) - (int)(base_j >> 1);
            const int cl = max(gl, 0), ch = min(gh, (int)HALF - 1);
            /* source and destination rows */
            uint32_t s4 = slotM + 1u; s4 = s4 >= 5u ? s4 - 5u : s4;             /* row si-4 = slot (si+1) mod 5 */
            uint32_t s2 = slotM + 3u; s2 = s2 >= 5u ? s2 - 5u : s2;             /* row si-2 */
            const uint32_t pe = (uint32_t)(si & 1);
            const uint32_t bM4 = sRing + s4 * RS, bM2 = sRing + s2 * RS, bI1 = sRing + (5u + (pe ^ 1u)) * RS, bD1 = sRing + (7u + (pe ^ 1u)) * RS;
            const uint32_t bMc = sRing + slotM * RS, bIc = sRing + (5u + pe) * RS, bDc = sRing + (7u + pe) * RS;
            CellT *grow = cells + off - ja;                                     /* cell of column j at grow[j] */
            int pmin = INT_MAX, pmax = INT_MIN, ka = INT_MIN, kb = INT_MAX;
            auto body = [&](auto initc) {
                constexpr bool INIT = decltype(initc)::value;
                for (int c = cl + (int)tid; c <= ch; c += (int)T) {
                    const uint32_t wa = 4u * (uint32_t)c + 4u;
                    const uint32_t a = lds_u32(bM4 + wa - 4u), b = lds_u32(bM4 + wa), d = lds_u32(bM4 + wa + 4u);
                    const uint32_t ia = lds_u32(bI1 + wa - 4u), ib = lds_u32(bI1 + wa);
                    const uint32_t db = lds_u32(bD1 + wa), dd = lds_u32(bD1 + wa + 4u);
                    const uint32_t xm = lds_u32(bM2 + wa);
                    const uint32_t j0 = base_j + 2u * (uint32_t)c;              /* columns j0, j0 + 1; n + k = j + 1 */
                    const bool act1 = j0 + 1u < W;                              /* the padding column past the last diagonal stays absent */
                    const uint32_t um = (uint32_t)m;
                    Cell3O c0 = next_off3(a >> 16, ia >> 16, b >> 16, db >> 16, xm & 0xffffu, um, j0 + 1u);
                    Cell3O c1 = next_off3(b & 0xffffu, ib & 0xffffu, d & 0xffffu, dd & 0xffffu, xm >> 16, act1 ? um : 0u, act1 ? j0 + 2u : 0u);
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
                    /* extend (wfa.go:394-455): a present cell has v >= 1, so it applies iff min(n - v, m - h) > 0 */
                    auto extend = [&](uint32_t M, const uint32_t j, const int k) -> uint32_t {
                        const int ext = (int)min(j + 1u, um) - (int)M;          /* min(n + k, m) - M */
                        if (M == 0u) return M;
                        if (ext > 0) {
                            const uint32_t v = (uint32_t)((int)M - k);
                            int l = 0;
                            do {
                                const uint32_t xx = Q.chunk(v + (uint32_t)l) ^ Tq.chunk(M + (uint32_t)l);
                                if (xx) { l += __clz((int)__brev(xx)) >> 1; break; }
                                l += 16;
                            } while (l < ext);
                            M += (uint32_t)min(l, ext);
                            if (l < ext) return M;
                        }
                        if (SEMI) {
                            /* a cell at the end of a sequence: start-cell classification (wfa.go:306-323 / :341-358) */
                            const int h = (int)M, v = h - k;
                            int cls = 0;
                            if (v <= 0 || v > n || h > m) cls = 1;
                            else if ((v == n && h >= n) || (h == m && v >= m)) cls = 2;
                            if (cls) {
                                const int key = (k + n) * 2 + (cls == 2);
                                if (k <= Ak) ka = max(ka, key); else kb = min(kb, key);
                            }
                        }
                        return M;
                    };
                    c0.M = extend(c0.M, j0, k0);
                    c1.M = extend(c1.M, j0 + 1u, k0 + 1);
                    sts_u32(bMc + wa, c0.M | c1.M << 16); sts_u32(bIc + wa, c0.I | c1.I << 16); sts_u32(bDc + wa, c0.D | c1.D << 16);
                    const CellT w0 = SC::pack(c0.M, c0.I, c0.D), w1 = SC::pack(c1.M, c1.I, c1.D);
                    *reinterpret_cast<ulonglong2 *>(grow + j0) = make_ulonglong2(w0, w1);
                    /* M WaveFront.Lo/Hi: first / last present cell */
                    if (c0.M) { pmin = min(pmin, (int)j0); pmax = max(pmax, (int)j0); }
                    if (c1.M) { pmin = min(pmin, (int)j0 + 1); pmax = max(pmax, (int)j0 + 1); }
                }
            };
            if (has_init) body(std::true_type{}); else body(std::false_type{});
            /* halo: the first / last column of the segment goes to the neighbour that reads it as k + 1 / k - 1 */
            if (cl <= ch) {
                if (rank > 0 && cl == 0 && tid == 0) {
                    const uint32_t left = rank - 1u;
                    st_cluster_u16(cluster_map(bMc + 4u + SEG * 2u, left), lds_u16(bMc + 4u));          /* M -> column SEG of the left CTA */
                    st_cluster_u16(cluster_map(bDc + 4u + SEG * 2u, left), lds_u16(bDc + 4u));
                }
                if (rank + 1u < C && ch == (int)HALF - 1 && tid == (uint32_t)(ch - cl) % T) {
                    const uint32_t right = rank + 1u;
                    st_cluster_u16(cluster_map(bMc + 2u, right), lds_u16(bMc + 2u + SEG * 2u));          /* M -> column -1 of the right CTA */
                    st_cluster_u16(cluster_map(bIc + 2u, right), lds_u16(bIc + 2u + SEG * 2u));
                }
                /* end test on diagonal m - n = column m - 1 (wfa.go:235-239), by the thread that wrote it */
                const int cA = (int)(((uint32_t)m - 1u) >> 1) - (int)(base_j >> 1);
                if (cA >= cl && cA <= ch && tid == (uint32_t)(cA - cl) % T)
                    endhit = lds_u16(bMc + 4u + ((uint32_t)m - 1u - base_j) * 2u) >= (uint32_t)m;
            }
            /* reductions: warp, block, then every CTA's partial result into every CTA's mailbox */
            pmin = __reduce_min_sync(0xffffffffu, pmin); pmax = __reduce_max_sync(0xffffffffu, pmax);
            const uint32_t eh = __any_sync(0xffffffffu, endhit) ? 1u : 0u;
            if (SEMI) { ka = __reduce_max_sync(0xffffffffu, ka); kb = __reduce_min_sync(0xffffffffu, kb); }
            if (lane == 0) {
                sts_u32(sRed + wid * 4u, (uint32_t)pmin); sts_u32(sRed + 128u + wid * 4u, (uint32_t)pmax); sts_u32(sRed + 256u + wid * 4u, eh);
                if (SEMI) { sts_u32(sRed + 384u + wid * 4u, (uint32_t)ka); sts_u32(sRed + 512u + wid * 4u, (uint32_t)kb); }
            }
            __syncthreads();
            if (wid == 0) {
                int a0 = lane < nw ? (int)lds_u32(sRed + lane * 4u) : INT_MAX, a1 = lane < nw ? (int)lds_u32(sRed + 128u + lane * 4u) : INT_MIN;
                uint32_t a2 = lane < nw ? lds_u32(sRed + 256u + lane * 4u) : 0u;
                int a3 = INT_MIN, a4 = INT_MAX;
                if (SEMI) { a3 = lane < nw ? (int)lds_u32(sRed + 384u + lane * 4u) : INT_MIN; a4 = lane < nw ? (int)lds_u32(sRed + 512u + lane * 4u) : INT_MAX; }
                a0 = __reduce_min_sync(0xffffffffu, a0); a1 = __reduce_max_sync(0xffffffffu, a1); a2 = __reduce_or_sync(0xffffffffu, a2);
                if (SEMI) { a3 = __reduce_max_sync(0xffffffffu, a3); a4 = __reduce_min_sync(0xffffffffu, a4); }
                if (lane < C) {
                    const uint32_t dst = cluster_map(sMail + parity * 256u + rank * 32u, lane);
                    st_cluster_u32(dst, (uint32_t)a0); st_cluster_u32(dst + 4u, (uint32_t)a1); st_cluster_u32(dst + 8u, a2);
                    st_cluster_u32(dst + 12u, (uint32_t)a3); st_cluster_u32(dst + 16u, (uint32_t)a4);
                }
            }
            cluster_sync_all();                                 /* row, halos and mailboxes are in place everywhere */
            {
                int g0 = INT_MAX, g1 = INT_MIN, g3 = INT_MIN, g4 = INT_MAX; uint32_t g2 = 0;
                for (uint32_t r = 0; r < C; r++) {
                    const uint32_t mb = sMail + parity * 256u + r * 32u;
                    g0 = min(g0, (int)lds_u32(mb)); g1 = max(g1, (int)lds_u32(mb + 4u)); g2 |= lds_u32(mb + 8u);
                    if (SEMI) { g3 = max(g3, (int)lds_u32(mb + 12u)); g4 = min(g4, (int)lds_u32(mb + 16u)); }
                }
                parity ^= 1u;
                exists = g0 <= g1;
                wlo = g0 - nm1; whi = g1 - nm1; endhit = g2 != 0u;
                if (SEMI) {
                    /* backtraceStartPosistion (wfa.go:270-375) for this one score: scan (a) runs from the end
                     * diagonal downwards, scan (b) upwards from the one above it; (b) overrides (a) */
                    if (g3 != INT_MIN && (g3 & 1)) { hit = true; hitK = (g3 >> 1) - n; }
                    if (g4 != INT_MAX && (g4 & 1)) { hit = true; hitK = (g4 >> 1) - n; }
                }
            }
            if (exists) {
                top = off;
                c_steps++; c_cells += (unsigned long long)(whi - wlo + 1); c_written += aw;
                hc = make_int4((int)ja - nm1, wlo, whi, (int)off);
            } else room += (long long)aw;
        }
        if (room < 0) { status = ST_ARENA; break; }
        if (rank == 0 && tid == 0) *reinterpret_cast<int4 *>(hdrs + si) = hc;
        lo5 = lo4; hi5 = hi4; lo4 = lo3; hi4 = hi3; lo3 = lo2; hi3 = hi2; lo2 = lo1; hi2 = hi1;
        lo1 = exists ? wlo : SLIM_NONE_LO; hi1 = exists ? whi : SLIM_NONE_HI;
        if (exists && endhit) { minS = (uint32_t)si * P.g; lastK = SEMI && hit ? hitK : Ak; break; }
        if (SEMI && exists && hit) { minS = (uint32_t)si * P.g; lastK = hitK; break; }
    }

    f.status = status; f.minS = minS; f.lastK = lastK; f.si = si; f.top = (uint64_t)top;
    f.c_cells = c_cells; f.c_written = c_written; f.c_steps = c_steps;
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
    finish_group_slim<1>(P, have, pair, f, P.arena + (uint64_t)(have ? item : 0u) * P.slot_bytes, P.slot_bytes);
}

} /* namespace wfak */
