/*
 * wfa_oracle.c -- literal CPU restatement of the reference's wavefront path.
 *
 * TEST INFRASTRUCTURE ONLY (see wfa_oracle.h).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).  The
 * storage layout is our own (dense per-score rows), the *semantics* of
 * Set/Get/Delete/Lo/Hi are the reference's, quirks included.
 */
#include "wfa_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* wfa_backtrace_types.go:23-37 */
enum { T_BITS = 3, T_MASK = 7 };
enum { T_INS_OPEN = 1, T_INS_EXT, T_DEL_OPEN, T_DEL_EXT, T_MISMATCH, T_MATCH };
static const char OPS_OF_TYPE[8] = {'.', 'I', 'I', 'D', 'D', 'X', 'M', 'H'};

/* ---------------------------------------------------------------- WaveFront
 * wfa_wavefront.go:45-183.  Lo/Hi start at +inf/-inf, widen on Set, shrink
 * only when the edge cell is deleted.  raw = offset<<3|type, 0 = absent. */
typedef struct {
    int       lo, hi;
    int       base;      /* k of a[0] */
    int       cap;
    uint32_t *a;
} wf_t;

static void wf_reset(wf_t *w) { w->lo = INT_MAX; w->hi = INT_MIN; if (w->cap) memset(w->a, 0, (size_t)w->cap * 4); }

static void wf_reserve(wf_t *w, int k)
{
    if (w->cap == 0) {
        w->cap = 64; w->base = k - 32; w->a = (uint32_t *)calloc((size_t)w->cap, 4);
        return;
    }
    if (k >= w->base && k < w->base + w->cap) return;
    int nlo = k < w->base ? k : w->base, nhi = k >= w->base + w->cap ? k + 1 : w->base + w->cap;
    int ncap = (nhi - nlo) * 2;
    int nbase = nlo - (ncap - (nhi - nlo)) / 2;
    uint32_t *na = (uint32_t *)calloc((size_t)ncap, 4);
    memcpy(na + (w->base - nbase), w->a, (size_t)w->cap * 4);
    free(w->a);
    w->a = na; w->cap = ncap; w->base = nbase;
}

/* wfa_wavefront.go:85-104 (Set) */
static void wf_set(wf_t *w, int k, uint32_t offset, uint32_t type)
{
    wf_reserve(w, k);
    w->a[k - w->base] = offset << T_BITS | type;
    if (k < w->lo) w->lo = k;
    if (k > w->hi) w->hi = k;
}
/* wfa_wavefront.go:130-150 (Increase) */
static void wf_increase(wf_t *w, int k, uint32_t delta)
{
    wf_reserve(w, k);
    w->a[k - w->base] += delta << T_BITS;
    if (k < w->lo) w->lo = k;
    if (k > w->hi) w->hi = k;
}
/* wfa_wavefront.go:162-168 (GetRaw); Get = GetRaw split (:153-159) */
static int wf_get_raw(const wf_t *w, int k, uint32_t *raw)
{
    if (k < w->lo || k > w->hi) { *raw = 0; return 0; }
    *raw = w->a[k - w->base];
    return *raw > 0;
}
/* wfa_wavefront.go:171-183 (Delete) */
static void wf_delete(wf_t *w, int k)
{
    if (k < w->lo || k > w->hi) return;
    w->a[k - w->base] = 0;
    if (k == w->hi) w->hi--;
    else if (k == w->lo) w->lo++;
}

/* ---------------------------------------------------------------- Component
 * wfa_component.go:37-187: wavefronts indexed by score, NULL = absent. */
typedef struct {
    wf_t   **wfs;
    uint32_t len;        /* len(WaveFronts) */
    uint32_t high;       /* 1 + highest score ever set since reset */
} comp_t;

typedef struct { wf_t **v; size_t n, cap; } wfpool_t;

struct oracle_aligner {
    oracle_config cfg;
    comp_t M, I, D;
    wfpool_t pool;
    uint64_t *ops; size_t nops, ops_cap;
    int *dist; size_t dist_cap;
    oracle_counters ctr;
    uint32_t last_score;
};

static wf_t *pool_get(wfpool_t *p)
{
    wf_t *w;
    if (p->n) w = p->v[--p->n];
    else w = (wf_t *)calloc(1, sizeof(wf_t));
    wf_reset(w);
    return w;
}
static void pool_put(wfpool_t *p, wf_t *w)
{
    if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 256; p->v = (wf_t **)realloc(p->v, p->cap * sizeof(wf_t *)); }
    p->v[p->n++] = w;
}

/* wfa_component.go:57-64 (Reset) */
static void comp_reset(comp_t *c, wfpool_t *p)
{
    for (uint32_t i = 0; i < c->high; i++)
        if (c->wfs[i]) { pool_put(p, c->wfs[i]); c->wfs[i] = NULL; }
    c->high = 0;
}
/* wfa_component.go:81-86 */
static int comp_has_score(const comp_t *c, uint32_t s) { return s < c->len && c->wfs[s] != NULL; }
/* wfa_component.go:91-101 (KRange): missing => (0,0) */
static void comp_krange(const comp_t *c, uint32_t s, uint32_t diff, int *lo, int *hi)
{
    *lo = 0; *hi = 0;
    if (diff > s) return;
    s -= diff;
    if (s >= c->len || !c->wfs[s]) return;
    *lo = c->wfs[s]->lo; *hi = c->wfs[s]->hi;
}
/* wfa_component.go:104-115 (Set) */
static void comp_set(comp_t *c, wfpool_t *p, uint32_t s, int k, uint32_t offset, uint32_t type)
{
    if (s >= c->len) {
        uint32_t nl = c->len ? c->len : 2048;
        while (s >= nl) nl *= 2;
        c->wfs = (wf_t **)realloc(c->wfs, (size_t)nl * sizeof(wf_t *));
        memset(c->wfs + c->len, 0, (size_t)(nl - c->len) * sizeof(wf_t *));
        c->len = nl;
    }
    if (!c->wfs[s]) c->wfs[s] = pool_get(p);
    if (s + 1 > c->high) c->high = s + 1;
    wf_set(c->wfs[s], k, offset, type);
}
/* wfa_component.go:142-147 (Get): plain score, used by backTrace with wrapped uint32 */
static int comp_get(const comp_t *c, uint32_t s, int k, uint32_t *offset, uint32_t *type)
{
    uint32_t raw;
    *offset = 0; if (type) *type = 0;
    if (s >= c->len || !c->wfs[s]) return 0;
    int ok = wf_get_raw(c->wfs[s], k, &raw);
    *offset = raw >> T_BITS; if (type) *type = raw & T_MASK;
    return ok;
}
/* wfa_component.go:150-155 (GetRaw) */
static int comp_get_raw(const comp_t *c, uint32_t s, int k, uint32_t *raw)
{
    *raw = 0;
    if (s >= c->len || !c->wfs[s]) return 0;
    return wf_get_raw(c->wfs[s], k, raw);
}
/* wfa_component.go:158-167 (GetAfterDiff): diff > s => absent */
static int comp_get_diff(const comp_t *c, uint32_t s, uint32_t diff, int k, uint32_t *offset)
{
    *offset = 0;
    if (diff > s) return 0;
    return comp_get(c, s - diff, k, offset, NULL);
}
/* wfa_component.go:182-187 (Delete) */
static void comp_delete(comp_t *c, uint32_t s, int k)
{
    if (s >= c->len || !c->wfs[s]) return;
    wf_delete(c->wfs[s], k);
}

/* ---------------------------------------------------------------- result ops
 * wfa_cigar.go:118-124 (AddN) */
static void ops_add(oracle_aligner *a, char op, uint32_t n)
{
    if (a->nops == a->ops_cap) { a->ops_cap = a->ops_cap ? a->ops_cap * 2 : 1024; a->ops = (uint64_t *)realloc(a->ops, a->ops_cap * 8); }
    a->ops[a->nops++] = (uint64_t)(uint8_t)op << 32 | (uint64_t)n;
}

/* wfa_cigar.go:136-214 (process): reverse, merge equal neighbours, stats
 * between the first and the last 'M' op (begin/end default to 0). */
static void ops_process(oracle_aligner *a, oracle_result *r)
{
    uint64_t *s = a->ops; size_t len = a->nops;
    if (len == 0) { r->n_ops = 0; return; }   /* unreachable: backTrace always emits >= 1 op */
    for (size_t i = 0, j = len - 1; i < j; i++, j--) { uint64_t t = s[i]; s[i] = s[j]; s[j] = t; }
    size_t j = 0; uint64_t pre = s[0];
    for (size_t i = 1; i < len; i++) {
        uint64_t op = s[i];
        if (op >> 32 == pre >> 32) { pre += op & 0xffffffffu; s[j] = pre; continue; }
        j++;
        if (i != j) s[j] = s[i];
        pre = op;
    }
    len = j + 1; a->nops = len;
    size_t begin = 0, end = 0;
    for (size_t i = 0; i < len; i++) if (s[i] >> 32 == 'M') { begin = i; break; }
    for (size_t i = len; i-- > 0;) if (s[i] >> 32 == 'M') { end = i; break; }
    uint32_t alen = 0, matches = 0, gaps = 0, regions = 0;
    for (size_t i = begin; i <= end; i++) {
        uint32_t n = (uint32_t)(s[i] & 0xffffffffu);
        alen += n;
        switch (s[i] >> 32) {
        case 'M': matches += n; break;
        case 'I': case 'D': gaps += n; regions++; break;
        }
    }
    r->align_len = alen; r->matches = matches; r->gaps = gaps; r->gap_regions = regions;
    r->n_ops = (uint32_t)len;
}

/* ---------------------------------------------------------------- hot path */

/* wfa.go:143-184 (initComponents) */
static void init_components(oracle_aligner *a, const uint8_t *q, int n, const uint8_t *t, int m)
{
    comp_reset(&a->M, &a->pool); comp_reset(&a->I, &a->pool); comp_reset(&a->D, &a->pool);
    const uint32_t x = a->cfg.mismatch;
    if (q[0] == t[0]) comp_set(&a->M, &a->pool, 0, 0, 1, T_MATCH);
    else              comp_set(&a->M, &a->pool, x, 0, 1, T_MISMATCH);
    if (!a->cfg.global_alignment) {
        for (int k = 1; k < m; k++) {                 /* first row */
            if (q[0] == t[k]) comp_set(&a->M, &a->pool, 0, k, (uint32_t)(k + 1), T_MATCH);
            else              comp_set(&a->M, &a->pool, x, k, (uint32_t)(k + 1), T_MISMATCH);
        }
        for (int k = 1; k < n; k++) {                 /* first column */
            if (q[k] == t[0]) comp_set(&a->M, &a->pool, 0, -k, 1, T_MATCH);
            else              comp_set(&a->M, &a->pool, x, -k, 1, T_MISMATCH);
        }
    }
}

/* wfa.go:381-458 (extend).  The 8-byte block loop plus the byte loop compute
 * exactly LCP(q[v:], t[h:]); we restate it as a byte LCP.  Returns Lo/Hi as
 * read on entry. */
static void extend(oracle_aligner *a, const uint8_t *q, int n, const uint8_t *t, int m, uint32_t s, int *plo, int *phi)
{
    wf_t *wf = a->M.wfs[s];
    int lo = wf->lo, hi = wf->hi;
    for (int k = hi; k >= lo; k--) {
        uint32_t raw;
        if (!wf_get_raw(wf, k, &raw)) continue;
        a->ctr.visits++;
        int h = (int)(raw >> T_BITS), v = h - k;
        if (v <= 0 || v >= n || h >= m) continue;      /* wfa.go:404 */
        int N = 0;
        while (q[v] == t[h]) { v++; h++; N++; if (v == n || h == m) break; }
        a->ctr.words += (uint64_t)(N + 1 + 15) / 16;
        if (N == 0) continue;
        wf_increase(wf, k, (uint32_t)N);
    }
    *plo = lo; *phi = hi;
}

/* wfa.go:461-540 (reduce) */
static void reduce(oracle_aligner *a, int n, int m, uint32_t s)
{
    wf_t *wf = a->M.wfs[s];
    int lo = wf->lo, hi = wf->hi;
    size_t w = (size_t)(hi - lo + 1);
    if (w > a->dist_cap) { a->dist_cap = w * 2; a->dist = (int *)realloc(a->dist, a->dist_cap * sizeof(int)); }
    int *ds = a->dist;
    int min_dist = INT_MAX;
    for (int k = lo; k <= hi; k++) {
        uint32_t raw;
        if (!wf_get_raw(wf, k, &raw)) { ds[k - lo] = -1; continue; }
        int h = (int)(raw >> T_BITS), v = h - k;
        if (v < 0 || v >= n || h >= m) { ds[k - lo] = -1; continue; }   /* wfa.go:483 */
        int d = (m - h) > (n - v) ? (m - h) : (n - v);
        ds[k - lo] = d;
        if (d < min_dist) min_dist = d;
    }
    int _lo = lo, _hi = hi;
    int max_diff = (int)a->cfg.max_dist_diff;
    int update_lo = 1, found = 0;
    for (size_t i = 0; i < w; i++) {
        int d = ds[i];
        if (d < 0) continue;
        if (d - min_dist > max_diff) {
            found = 1;
            if (update_lo) _lo = lo + (int)i + 1;
            ds[i] = -1;
        } else update_lo = 0;
    }
    if (found)
        for (size_t i = w; i-- > 0;)
            if (ds[i] >= 0) { _hi = lo + (int)i; break; }
    for (int k = lo; k < _lo; k++)      { wf_delete(wf, k); comp_delete(&a->I, s, k); comp_delete(&a->D, s, k); }
    for (int k = _hi + 1; k <= hi; k++) { wf_delete(wf, k); comp_delete(&a->I, s, k); comp_delete(&a->D, s, k); }
    wf->lo = _lo; wf->hi = _hi;
}

static inline uint32_t max2(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* wfa.go:549-700 (next) */
static void next(oracle_aligner *a, int n, int m, uint32_t s)
{
    comp_t *M = &a->M, *I = &a->I, *D = &a->D; wfpool_t *P = &a->pool;
    const uint32_t x = a->cfg.mismatch, oe = a->cfg.gap_open + a->cfg.gap_ext, e = a->cfg.gap_ext;
    int l1, h1, l2, h2, l3, h3, l4, h4;
    comp_krange(M, s, x, &l1, &h1);
    comp_krange(M, s, oe, &l2, &h2);
    comp_krange(I, s, e, &l3, &h3);
    comp_krange(D, s, e, &l4, &h4);
    int hi = imin(m - 1, imax(imax(h1, h2), imax(h3, h4)) + 1);
    int lo = imax(-(n - 1), imin(imin(l1, l2), imin(l3, l4)) - 1);
    for (int k = lo; k <= hi; k++) {
        uint32_t v1, v2, Isk, Dsk, Msk, tI = 0, tD = 0, tM = 0;
        int fromM, fromI, fromD, updI = 0, updD = 0;
        /* insertion, wfa.go:579-609 */
        fromM = comp_get_diff(M, s, oe, k - 1, &v1);
        fromI = comp_get_diff(I, s, e, k - 1, &v2);
        if (fromM && (int)v1 > m) { fromM = 0; v1 = 0; }
        if (fromI && (int)v2 > m) { fromI = 0; v2 = 0; }
        Isk = max2(v1, v2) + 1;
        if (fromM || fromI) {
            if (fromM && fromI) tI = v1 >= v2 ? T_INS_OPEN : T_INS_EXT;
            else tI = fromM ? T_INS_OPEN : T_INS_EXT;
            updI = 1;
            comp_set(I, P, s, k, Isk, tI);
        } else Isk = 0;
        /* deletion, wfa.go:614-645 */
        fromM = comp_get_diff(M, s, oe, k + 1, &v1);
        fromD = comp_get_diff(D, s, e, k + 1, &v2);
        if (fromM && (int)v1 - k > n) { fromM = 0; v1 = 0; }
        if (fromD && (int)v2 - k > n) { fromD = 0; v2 = 0; }
        Dsk = max2(v1, v2);
        if (fromM || fromD) {
            if (fromM && fromD) tD = v1 >= v2 ? T_DEL_OPEN : T_DEL_EXT;
            else tD = fromM ? T_DEL_OPEN : T_DEL_EXT;
            updD = 1;
            comp_set(D, P, s, k, Dsk, tD);
        } else Dsk = 0;
        /* mismatch, wfa.go:650-698 */
        fromM = comp_get_diff(M, s, x, k, &v1);
        if (fromM && ((int)v1 > m || (int)v1 - k > n)) { fromM = 0; v1 = 0; }
        Msk = max2(max2(Isk, Dsk), v1 + 1);
        if (updI || updD || fromM) {
            if (updI && updD && fromM) {
                if (Msk == v1 + 1) tM = T_MISMATCH; else if (Msk == Isk) tM = tI; else tM = tD;
            } else if (updI) {
                if (updD) tM = Msk == Isk ? tI : tD;
                else if (fromM) tM = Msk == v1 + 1 ? T_MISMATCH : tI;
                else tM = tI;
            } else if (updD) {
                if (fromM) tM = Msk == v1 + 1 ? T_MISMATCH : tD;
                else tM = tD;
            } else tM = T_MISMATCH;
            comp_set(M, P, s, k, Msk, tM);
        }
    }
}

/* wfa.go:270-375 (backtraceStartPosistion) */
static void backtrace_start(oracle_aligner *a, int n, int m, uint32_t s, uint32_t *pminS, int *plastK)
{
    comp_t *M = &a->M;
    uint32_t minS = s; int Ak = m - n, lastK = Ak;
    for (uint32_t _s = s;; _s--) {
        if (comp_has_score(M, _s)) {
            int lo, hi; comp_krange(M, _s, 0, &lo, &hi);
            for (int pass = 0; pass < 2; pass++) {
                int hit = 0, k = pass == 0 ? Ak : Ak + 1;
                for (;;) {
                    if (pass == 0 ? k < lo : k > hi) break;
                    uint32_t off;
                    if (!comp_get_diff(M, _s, 0, k, &off)) { k += pass == 0 ? -1 : 1; continue; }
                    int h = (int)off, v = h - k;
                    if (v <= 0 || v > n || h > m) break;                        /* wfa.go:314,349 */
                    if ((v == n && h >= n) || (h == m && v >= m)) { hit = 1; break; } /* :319,354 */
                    k += pass == 0 ? -1 : 1;
                }
                if (hit && _s <= minS) { lastK = k; minS = _s; }
            }
        }
        if (_s == 0) break;
    }
    *pminS = minS; *plastK = lastK;
}

/* wfa.go:703-983 (backTrace) + wfa_cigar.go:136-214 (process) */
static void back_trace(oracle_aligner *a, int n, int m, uint32_t s, int Ak, oracle_result *res)
{
    const int semi = !a->cfg.global_alignment;
    comp_t *M = &a->M, *I = &a->I, *D = &a->D, *M0 = NULL;
    const uint32_t x = a->cfg.mismatch, o = a->cfg.gap_open, e = a->cfg.gap_ext;
    a->nops = 0;
    res->score = s;
    /* wfa_cigar.go:77-89 reset() leaves TBegin..QEnd stale; we define them 0. */
    res->tbegin = res->tend = res->qbegin = res->qend = 0;

    int k = Ak, h, v, h0, first_match = 1, prev_from_M = 1, n_matches;
    int q_begin = 0, t_begin = 0;
    uint32_t offset, type, v1, v2, Isk = 0, Dsk = 0, offset0 = 0;
    int from_MI, from_MD, from_itself = 0, fromI, fromD, fromM;
    uint32_t sX, sO, sE;

    comp_get_raw(M, s, k, &offset);
    type = offset & T_MASK;
    h = (int)(offset >> T_BITS);
    v = h - k;
    if (h < m) ops_add(a, OPS_OF_TYPE[T_INS_OPEN], (uint32_t)m - (uint32_t)h);
    else if (v < n) ops_add(a, 'H', (uint32_t)n - (uint32_t)v);

    while (v > 0 && h > 0) {
        sX = s - x; sO = s - o - e; sE = s - e;        /* uint32 wrap, wfa.go:760-762 */
        from_MI = from_MD = 0;
        switch (type) {
        case T_INS_EXT:
            fromM = comp_get(M, sO, k - 1, &v1, NULL);
            fromI = comp_get(I, sE, k - 1, &v2, NULL);
            if (fromM || fromI) { from_MI = 1; offset0 = max2(v1, v2) + 1; } else offset0 = 0;
            M0 = I;
            break;
        case T_DEL_EXT:
            fromM = comp_get(M, sO, k + 1, &v1, NULL);
            fromD = comp_get(D, sE, k + 1, &v2, NULL);
            if (fromM || fromD) { from_MD = 1; offset0 = max2(v1, v2); } else offset0 = 0;
            M0 = D;
            break;
        default:
            fromM = comp_get(M, sO, k - 1, &v1, NULL);
            fromI = comp_get(I, sE, k - 1, &v2, NULL);
            if (fromM || fromI) { from_MI = 1; Isk = max2(v1, v2) + 1; } else Isk = 0;
            fromM = comp_get(M, sO, k + 1, &v1, NULL);
            fromD = comp_get(D, sE, k + 1, &v2, NULL);
            if (fromM || fromD) { from_MD = 1; Dsk = max2(v1, v2); } else Dsk = 0;
            fromM = comp_get(M, sX, k, &v1, NULL);
            if (from_MI || from_MD || fromM) { offset0 = max2(max2(Isk, Dsk), v1 + 1); from_itself = 0; }
            else from_itself = 1;
            M0 = M;
        }
        if (from_itself) break;
        if (offset0 == 0) break;
        h0 = (int)offset0;

        if (prev_from_M) {                               /* wfa.go:833-869 */
            n_matches = h - h0;
            if (n_matches > 0) {
                if (first_match) { first_match = 0; res->tend = h; res->qend = v; }
                ops_add(a, OPS_OF_TYPE[T_MATCH], (uint32_t)n_matches);
            }
            offset = offset0; h = (int)offset; v = h - k;
            if (type == T_MATCH) { t_begin = h; q_begin = v; }
            else if (n_matches > 0) { t_begin = h + 1; q_begin = v + 1; }
            if (h <= 0 || v <= 0) break;
        }
        ops_add(a, OPS_OF_TYPE[type], 1);               /* wfa.go:872-873 */
        if (semi && (h == 1 || v == 1)) break;          /* :876-879 */

        prev_from_M = 1;
        int leave = 0;
        switch (type) {                                  /* :886-909 */
        case T_MISMATCH: s = sX; h--; break;
        case T_INS_OPEN: s = sO; k--; h--; break;
        case T_INS_EXT:  s = sE; k--; h--; prev_from_M = 0; break;
        case T_DEL_OPEN: s = sO; k++; break;
        case T_DEL_EXT:  s = sE; k++; prev_from_M = 0; break;
        default: leave = 1;
        }
        if (leave) break;
        v = h - k;
        if (!comp_get_raw(M0, s, k, &offset)) break;     /* :915-919 */
        type = offset & T_MASK;
    }

    if (h > 0 && v > 0) {                                /* wfa.go:930-968 */
        n_matches = imin(h, v) - 1;
        if (n_matches > 0) {
            if (first_match) { first_match = 0; res->tend = h; res->qend = v; }
            ops_add(a, OPS_OF_TYPE[T_MATCH], (uint32_t)n_matches);
            h -= n_matches; v -= n_matches;
            if (type == T_MATCH) { t_begin = h; q_begin = v; }
            else { t_begin = h + 1; q_begin = v + 1; }
        } else if (type == T_MATCH) {
            t_begin = h; q_begin = v;
            if (first_match) { first_match = 0; res->tend = h; res->qend = v; }
        }
        ops_add(a, OPS_OF_TYPE[type], 1);
    }
    if (v > 1) ops_add(a, 'H', (uint32_t)(v - 1));       /* :970-972 */
    if (h > 1) ops_add(a, OPS_OF_TYPE[T_INS_OPEN], (uint32_t)(h - 1));
    res->tbegin = t_begin; res->qbegin = q_begin;
    ops_process(a, res);
}

/* ---------------------------------------------------------------- API */

oracle_aligner *oracle_new(const oracle_config *cfg)
{
    oracle_aligner *a = (oracle_aligner *)calloc(1, sizeof(*a));
    a->cfg = *cfg;
    return a;
}

static void comp_free(comp_t *c)
{
    for (uint32_t i = 0; i < c->len; i++) if (c->wfs[i]) { free(c->wfs[i]->a); free(c->wfs[i]); }
    free(c->wfs);
}

void oracle_free(oracle_aligner *a)
{
    if (!a) return;
    comp_free(&a->M); comp_free(&a->I); comp_free(&a->D);
    for (size_t i = 0; i < a->pool.n; i++) { free(a->pool.v[i]->a); free(a->pool.v[i]); }
    free(a->pool.v); free(a->ops); free(a->dist); free(a);
}

/* wfa.go:201-268 (AlignPointers) */
int oracle_align(oracle_aligner *a, const uint8_t *q, uint32_t qn, const uint8_t *t, uint32_t tm,
                 oracle_result *res, const uint64_t **ops, oracle_counters *ctr)
{
    memset(res, 0, sizeof(*res));
    memset(&a->ctr, 0, sizeof(a->ctr));
    if (ops) *ops = NULL;
    if (qn == 0 || tm == 0) { res->status = ORACLE_EMPTY; return ORACLE_EMPTY; }
    if (qn > ORACLE_MAX_SEQ_LEN || tm > ORACLE_MAX_SEQ_LEN) { res->status = ORACLE_TOO_LONG; return ORACLE_TOO_LONG; }
    int n = (int)qn, m = (int)tm;

    init_components(a, q, n, t, m);

    int Ak = m - n; uint32_t Aoffset = (uint32_t)m, s = 0;
    const int do_reduce = a->cfg.adaptive;
    const int min_wf_len = (int)a->cfg.min_wf_len;
    for (;;) {
        if (comp_has_score(&a->M, s)) {
            int lo, hi; uint32_t offset;
            extend(a, q, n, t, m, s, &lo, &hi);
            a->ctr.scores++;
            a->ctr.cells += (uint64_t)(hi - lo + 1);
            if ((uint64_t)(hi - lo + 1) > a->ctr.max_width) a->ctr.max_width = (uint64_t)(hi - lo + 1);
            comp_get_diff(&a->M, s, 0, Ak, &offset);
            if (offset >= Aoffset) break;
            if (do_reduce && hi - lo + 1 >= min_wf_len) reduce(a, n, m, s);
        }
        s++;
        next(a, n, m, s);
    }
    a->last_score = s;
    uint32_t minS = s; int lastK = Ak;
    if (!a->cfg.global_alignment) backtrace_start(a, n, m, s, &minS, &lastK);
    back_trace(a, n, m, minS, lastK, res);
    a->ctr.ops = a->nops;
    if (ops) *ops = a->ops;
    if (ctr) *ctr = a->ctr;
    return ORACLE_OK;
}

static const comp_t *pick(const oracle_aligner *a, int comp) { return comp == 0 ? &a->M : comp == 1 ? &a->I : &a->D; }
int oracle_get_raw(const oracle_aligner *a, int comp, uint32_t s, int k, uint32_t *raw) { return comp_get_raw(pick(a, comp), s, k, raw); }
int oracle_krange(const oracle_aligner *a, int comp, uint32_t s, int *lo, int *hi)
{
    const comp_t *c = pick(a, comp);
    if (!comp_has_score(c, s)) return 0;
    *lo = c->wfs[s]->lo; *hi = c->wfs[s]->hi;
    return 1;
}
uint32_t oracle_max_score(const oracle_aligner *a) { return a->last_score; }

/* ---------------------------------------------------------------- batch */
typedef struct {
    const oracle_config *cfg; uint64_t n_pairs; const uint8_t *seq;
    const uint64_t *q_off, *t_off; const uint32_t *q_len, *t_len;
    oracle_result *results; uint64_t **ops_tmp;   /* per pair malloc'd copy or NULL */
    int want_ops; uint64_t *next; oracle_counters ctr;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    oracle_aligner *a = oracle_new(j->cfg);
    for (;;) {
        uint64_t i0 = __atomic_fetch_add(j->next, 64, __ATOMIC_RELAXED);
        if (i0 >= j->n_pairs) break;
        uint64_t i1 = i0 + 64 < j->n_pairs ? i0 + 64 : j->n_pairs;
        for (uint64_t i = i0; i < i1; i++) {
            const uint64_t *ops; oracle_counters c;
            int st = oracle_align(a, j->seq + j->q_off[i], j->q_len[i], j->seq + j->t_off[i], j->t_len[i], &j->results[i], &ops, &c);
            if (st != ORACLE_OK) continue;
            j->ctr.cells += c.cells; j->ctr.visits += c.visits; j->ctr.words += c.words; j->ctr.ops += c.ops; j->ctr.scores += c.scores;
            if (c.max_width > j->ctr.max_width) j->ctr.max_width = c.max_width;
            if (j->want_ops) {
                j->ops_tmp[i] = (uint64_t *)malloc((size_t)j->results[i].n_ops * 8);
                memcpy(j->ops_tmp[i], ops, (size_t)j->results[i].n_ops * 8);
            }
        }
    }
    oracle_free(a);
    return NULL;
}

int oracle_align_batch(const oracle_config *cfg, uint64_t n_pairs, const uint8_t *seq_bytes,
                       const uint64_t *q_off, const uint32_t *q_len, const uint64_t *t_off, const uint32_t *t_len,
                       oracle_result *results, uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off,
                       uint64_t *ops_needed, int nthreads, oracle_counters *ctr_sum)
{
    if (nthreads < 1) nthreads = 1;
    uint64_t next = 0;
    uint64_t **tmp = ops ? (uint64_t **)calloc(n_pairs ? n_pairs : 1, sizeof(uint64_t *)) : NULL;
    job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) {
        jobs[i] = (job_t){cfg, n_pairs, seq_bytes, q_off, t_off, q_len, t_len, results, tmp, ops != NULL, &next, {0}};
        pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    oracle_counters sum; memset(&sum, 0, sizeof(sum));
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        sum.cells += jobs[i].ctr.cells; sum.visits += jobs[i].ctr.visits; sum.words += jobs[i].ctr.words;
        sum.ops += jobs[i].ctr.ops; sum.scores += jobs[i].ctr.scores;
        if (jobs[i].ctr.max_width > sum.max_width) sum.max_width = jobs[i].ctr.max_width;
    }
    if (ctr_sum) *ctr_sum = sum;
    int rc = 0;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_pairs; i++) {
        if (ops_off) ops_off[i] = total;
        total += results[i].n_ops;
    }
    if (ops_needed) *ops_needed = total;
    if (ops) {
        if (total > ops_capacity) rc = -1;
        else for (uint64_t i = 0, at = 0; i < n_pairs; i++) {
            if (tmp[i]) memcpy(ops + at, tmp[i], (size_t)results[i].n_ops * 8);
            at += results[i].n_ops;
        }
        for (uint64_t i = 0; i < n_pairs; i++) free(tmp[i]);
        free(tmp);
    }
    free(jobs); free(th);
    return rc;
}
