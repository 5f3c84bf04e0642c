/*
 * wfa.hpp -- C++ host-side mirror of the reference's Go API on top of the C ABI
 * (include/wfacuda.h).  The reference is compiled Go and no Go toolchain exists
 * in this image, so the host layer above the C ABI is C++: same names, argument
 * meaning and error behaviour as /root/reference/wfa.go and wfa_cigar.go, so
 * that code written against the Go package ports line by line:
 *
 *     auto *algn = wfa::New(&wfa::DefaultPenalties, &wfa::DefaultOptions);
 *     algn->AdaptiveReduction(&wfa::DefaultAdaptiveOption);
 *     wfa::AlignmentResult *r; wfa::Error err = algn->Align(q, t, &r);
 *     r->CIGAR(false); r->AlignmentText(q, t, false, &Q, &A, &T);
 *     wfa::RecycleAlignmentResult(r); wfa::RecycleAligner(algn);
 *
 * plus the batched entry point AlignBatch.  Header-only; link with -lwfacuda.
 * There is no CPU path here: every Align goes through libwfacuda.so.
 */
#ifndef WFA_HOST_WFA_HPP
#define WFA_HOST_WFA_HPP

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wfacuda.h"

namespace wfa {

struct Penalties { uint32_t Mismatch, GapOpen, GapExt; };                    /* wfa.go:32-36 */
struct AdaptiveReductionOption { uint32_t MinWFLen, MaxDistDiff, CutoffStep; }; /* wfa.go:46-50 */
struct Options { bool GlobalAlignment; };                                    /* wfa.go:64-66 */

static Penalties DefaultPenalties = {4, 6, 2};                               /* wfa.go:39-43 */
static AdaptiveReductionOption DefaultAdaptiveOption = {10, 50, 1};          /* wfa.go:56-60 */
static Options DefaultOptions = {true};                                      /* wfa.go:69-71 */

constexpr int MaxSeqLen = (1 << 29) - 1;                                     /* wfa.go:190 */

/* Go `error`: nullptr == nil; sentinels compare by identity like the reference's. */
typedef const char *Error;
static const char ErrEmptySeqText[] = "wfa: invalid empty sequence";                              /* wfa.go:187 */
static const char ErrSeqTooLongText[] = "wfa: sequences longer than 536870911 are not supported"; /* wfa.go:193 */
static const char ErrResourcesText[] = "wfacuda: pair needs more device memory than available";
static const char ErrCutoffText[] = "cutoff step should not be 0";                                /* wfa.go:136 */
static const Error ErrEmptySeq = ErrEmptySeqText;
static const Error ErrSeqTooLong = ErrSeqTooLongText;
static const Error ErrResources = ErrResourcesText;

constexpr uint64_t OpM = 'M', OpD = 'D', OpI = 'I', OpX = 'X', OpH = 'H';    /* wfa_cigar.go:60-64 */
constexpr uint64_t MaskLower32 = 4294967295ull;

inline void Op(uint64_t op, char *o, uint32_t *n) { *o = (char)(op >> 32); *n = (uint32_t)(op & MaskLower32); } /* :56-58 */

/* The Ops of a result: like a Go slice, a view of words owned elsewhere -- the Aligner's page-locked
 * output buffer for batch results (valid until the next Align / AlignBatch call on that Aligner
 * or RecycleAligner), or the result's own copy after Detach(). */
struct OpsView {
    const uint64_t *p = nullptr; size_t n = 0;
    size_t size() const { return n; }
    const uint64_t &operator[](size_t i) const { return p[i]; }
    const uint64_t *begin() const { return p; }
    const uint64_t *end() const { return p + n; }
};

/* wfa_cigar.go:29-46, as left by process() (:136-214) -- computed on the GPU */
struct AlignmentResult {
    OpsView Ops;
    std::vector<uint64_t> own_;                       /* storage of a detached result */
    bool pooled_ = false;                             /* lives in its Aligner's result pool (wfa_cigar.go:62-96) */
    /* copy the ops out of the Aligner's buffer: the result then outlives the next call */
    void Detach() { if (Ops.p != own_.data() || own_.empty()) { own_.assign(Ops.begin(), Ops.end()); Ops.p = own_.data(); Ops.n = own_.size(); } }
    uint32_t Score = 0;
    int TBegin = 0, TEnd = 0, QBegin = 0, QEnd = 0;
    uint32_t AlignLen = 0, Matches = 0, Gaps = 0, GapRegions = 0;

    /* trimOps, wfa_cigar.go:217-233 */
    void aligned_range(size_t *b, size_t *e) const
    {
        long start = -1, end = -1;
        for (size_t i = 0; i < Ops.size(); i++) if (Ops[i] >> 32 == OpM) { start = (long)i; break; }
        for (size_t i = Ops.size(); i-- > 0;) if (Ops[i] >> 32 == OpM) { end = (long)i; break; }
        *b = start < 0 ? 0 : (size_t)start; *e = (size_t)(end + 1);
        if (start < 0) *e = 0;
    }
    /* wfa_cigar.go:236-255 */
    std::string CIGAR(bool onlyAignedRegion) const
    {
        size_t b = 0, e = Ops.size();
        if (onlyAignedRegion) aligned_range(&b, &e);
        std::string s;
        for (size_t i = b; i < e; i++) { s += std::to_string(Ops[i] & MaskLower32); s += (char)(Ops[i] >> 32); }
        return s;
    }
    /* wfa_cigar.go:259-333 */
    void AlignmentText(const std::string &q0, const std::string &t0, bool onlyAignedRegion,
                       std::string *Q, std::string *A, std::string *T) const
    {
        size_t b = 0, e = Ops.size();
        std::string q = q0, t = t0;
        if (onlyAignedRegion) {
            q = q0.substr((size_t)QBegin - 1, (size_t)(QEnd - QBegin + 1));
            t = t0.substr((size_t)TBegin - 1, (size_t)(TEnd - TBegin + 1));
            aligned_range(&b, &e);
        }
        Q->clear(); A->clear(); T->clear();
        size_t v = 0, h = 0;
        for (size_t i = b; i < e; i++) {
            const uint64_t n = Ops[i] & MaskLower32;
            switch (Ops[i] >> 32) {
            case OpM: case OpX:
                Q->append(q, v, n); A->append(n, (Ops[i] >> 32) == OpM ? '|' : ' '); T->append(t, h, n); v += n; h += n; break;
            case OpI:
                Q->append(n, '-'); A->append(n, ' '); T->append(t, h, n); h += n; break;
            case OpD: case OpH:
                Q->append(q, v, n); A->append(n, ' '); T->append(n, '-'); v += n; break;
            }
        }
    }
};

class Aligner {
public:
    /* wfa.go:79-87: one Aligner per thread, not safe for concurrent Align calls */
    Aligner(const Penalties *p, const Options *opt, int device) : p_(*p), opt_(*opt)
    {
        wfacuda_config c = config();
        ctx_ = wfacuda_create(device, &c);
        if (!ctx_) err_ = wfacuda_last_error(nullptr);
    }
    ~Aligner() { release_buffers(); if (ctx_) wfacuda_destroy(ctx_); }
    Aligner(const Aligner &) = delete;
    Aligner &operator=(const Aligner &) = delete;
    bool ok() const { return ctx_ != nullptr; }
    const std::string &error() const { return err_; }

    /* wfa.go:134-140 */
    Error AdaptiveReduction(const AdaptiveReductionOption *ad)
    {
        if (ad->MinWFLen == 0) return ErrCutoffText;
        ad_ = *ad; has_ad_ = true;
        wfacuda_config c = config();
        if (wfacuda_set_config(ctx_, &c) != 0) { err_ = wfacuda_last_error(ctx_); return err_.c_str(); }
        return nullptr;
    }

    /* wfa.go:196-198.  The result is detached (owns its ops): free it with RecycleAlignmentResult. */
    Error Align(const std::string &q, const std::string &t, AlignmentResult **out)
    {
        std::vector<AlignmentResult *> rs; std::vector<Error> es;
        Error e = AlignBatch({q}, {t}, &rs, &es);
        *out = nullptr;
        if (e) return e;
        if (rs[0]) { AlignmentResult *r = new AlignmentResult(*rs[0]); r->pooled_ = false; r->Ops = rs[0]->Ops; r->own_.clear(); r->Detach(); *out = r; }
        return es[0];
    }

    /* New: many pairs per call -- the reference's call shape: [][]byte in, []*AlignmentResult out
     * (wfa.go:196-201, result pool wfa_cigar.go:62-96).  results[i] == nullptr where errors[i] != nil.
     * The sequences are flattened into one page-locked pool that the Aligner keeps across calls
     * (the DMA engines read it directly), by several host threads; the result objects come from
     * the Aligner's pool and their Ops are views of its page-locked output buffer: valid until the
     * next call on this Aligner (Detach() copies), RecycleAlignmentResult is a no-op for them. */
    Error AlignBatch(const std::vector<std::string> &qs, const std::vector<std::string> &ts,
                     std::vector<AlignmentResult *> *results, std::vector<Error> *errors)
    {
        return align_batch_on(&ctx_, 1, qs, ts, results, errors);
    }

    /* The same over several devices: one Aligner (ctx) per device, sharded inside the library
     * (wfacuda_align_batch_multi: length-binned LPT, one host thread per device, no collective).
     * Buffers and result pool are this Aligner's. */
    Error AlignBatchMulti(const std::vector<Aligner *> &others, const std::vector<std::string> &qs, const std::vector<std::string> &ts,
                          std::vector<AlignmentResult *> *results, std::vector<Error> *errors)
    {
        std::vector<wfacuda_ctx *> ctxs{ctx_};
        for (Aligner *o : others) ctxs.push_back(o->ctx_);
        return align_batch_on(ctxs.data(), (int)ctxs.size(), qs, ts, results, errors);
    }

    /* CIGAR(onlyAignedRegion) and the AlignmentText lines of every pair of a batch, formatted on the
     * GPU (wfacuda_batch_render; wfa_cigar.go:236-333).  cigars[i] / Q,A,T[i] are empty where errors[i] != nil. */
    Error AlignBatchRendered(const std::vector<std::string> &qs, const std::vector<std::string> &ts, bool onlyAignedRegion,
                             std::vector<std::string> *cigars, std::vector<std::string> *Q, std::vector<std::string> *A,
                             std::vector<std::string> *T, std::vector<Error> *errors)
    {
        const size_t n = qs.size();
        std::vector<uint8_t> pool; std::vector<uint64_t> qo(n), to(n); std::vector<uint32_t> ql(n), tl(n);
        for (size_t i = 0; i < n; i++) {
            qo[i] = pool.size(); ql[i] = (uint32_t)qs[i].size(); pool.insert(pool.end(), qs[i].begin(), qs[i].end());
            to[i] = pool.size(); tl[i] = (uint32_t)ts[i].size(); pool.insert(pool.end(), ts[i].begin(), ts[i].end());
        }
        pool.resize(pool.size() + 16);
        wfacuda_batch *b = wfacuda_batch_upload(ctx_, n, pool.data(), qo.data(), ql.data(), to.data(), tl.data());
        if (!b) { err_ = wfacuda_last_error(ctx_); return err_.c_str(); }
        std::vector<wfacuda_result> res(n); std::vector<uint64_t> off(n);
        int rc = wfacuda_batch_run(ctx_, b);
        if (rc == 0) rc = wfacuda_batch_download(ctx_, b, res.data(), nullptr, 0, off.data());
        std::vector<uint8_t> cig(1), txt(1); std::vector<uint64_t> co(n), xo(n); std::vector<uint32_t> cl(n), xl(n);
        for (int pass = 0; rc == 0 && pass < 2; pass++) {       /* first call sizes the buffers */
            rc = wfacuda_batch_render(ctx_, b, onlyAignedRegion ? 1 : 0, cig.data(), cig.size(), co.data(), cl.data(),
                                      txt.data(), txt.size(), xo.data(), xl.data());
            if (rc != WFACUDA_E_OPS_CAPACITY) break;
            uint64_t a = 0, c = 0;
            wfacuda_last_render_total(ctx_, &a, &c);
            cig.resize(a + 1); txt.resize(c + 1);
            rc = 0;
            if (pass == 1) rc = WFACUDA_E_OPS_CAPACITY;
        }
        wfacuda_batch_free(ctx_, b);
        if (rc != 0) { err_ = wfacuda_last_error(ctx_); return err_.c_str(); }
        cigars->assign(n, ""); Q->assign(n, ""); A->assign(n, ""); T->assign(n, ""); errors->assign(n, nullptr);
        for (size_t i = 0; i < n; i++) {
            if (res[i].status == WFACUDA_ERR_EMPTY_SEQ) { (*errors)[i] = ErrEmptySeq; continue; }
            if (res[i].status == WFACUDA_ERR_SEQ_TOO_LONG) { (*errors)[i] = ErrSeqTooLong; continue; }
            if (res[i].status != WFACUDA_OK) { (*errors)[i] = ErrResources; continue; }
            (*cigars)[i].assign((const char *)cig.data() + co[i], cl[i]);
            (*Q)[i].assign((const char *)txt.data() + xo[i], xl[i]);
            (*A)[i].assign((const char *)txt.data() + xo[i] + xl[i], xl[i]);
            (*T)[i].assign((const char *)txt.data() + xo[i] + 2ull * xl[i], xl[i]);
        }
        return nullptr;
    }

    /* Aligner.M / I / D of one pair (wfa.go:80-86) from the GPU's wavefront store: rows of
     * {score, lo, hi, first_cell} and three raw words offset<<3|code per diagonal (wfacuda_align_components). */
    struct Components {
        std::vector<wfacuda_wavefront> rows; std::vector<uint32_t> cells;
        /* Component.GetRaw (wfa_component.go:148-155): comp 0 = M, 1 = I, 2 = D; 0 = absent */
        uint32_t GetRaw(int comp, uint32_t s, int k) const
        {
            for (const wfacuda_wavefront &w : rows)
                if (w.score == s) return (k < w.lo || k > w.hi) ? 0u : cells[w.first_cell + 3ull * (uint64_t)(k - w.lo) + (uint64_t)comp];
            return 0u;
        }
    };
    Error AlignComponents(const std::string &q, const std::string &t, AlignmentResult **out, Components *comps)
    {
        wfacuda_result res{}; std::vector<uint64_t> ops(q.size() + t.size() + 16);
        uint32_t nr = 0; uint64_t nc = 0;
        comps->rows.assign(1, wfacuda_wavefront{}); comps->cells.assign(1, 0u);
        int rc = 0;
        for (int pass = 0; pass < 2; pass++) {
            rc = wfacuda_align_components(ctx_, (const uint8_t *)q.data(), (uint32_t)q.size(), (const uint8_t *)t.data(), (uint32_t)t.size(),
                                          &res, ops.data(), ops.size(), comps->rows.data(), (uint32_t)comps->rows.size(), &nr,
                                          comps->cells.data(), comps->cells.size(), &nc);
            if (rc != WFACUDA_E_OPS_CAPACITY) break;
            comps->rows.assign(nr + 1, wfacuda_wavefront{}); comps->cells.assign(nc + 1, 0u);
        }
        *out = nullptr;
        if (rc != 0) { err_ = wfacuda_last_error(ctx_); return err_.c_str(); }
        if (res.status == WFACUDA_ERR_EMPTY_SEQ) return ErrEmptySeq;
        if (res.status == WFACUDA_ERR_SEQ_TOO_LONG) return ErrSeqTooLong;
        if (res.status != WFACUDA_OK) return ErrResources;
        comps->rows.resize(nr); comps->cells.resize(nc);
        AlignmentResult *r = new AlignmentResult();
        r->own_.assign(ops.begin(), ops.begin() + res.n_ops); r->Ops.p = r->own_.data(); r->Ops.n = r->own_.size();
        r->Score = res.score; r->TBegin = res.tbegin; r->TEnd = res.tend; r->QBegin = res.qbegin; r->QEnd = res.qend;
        r->AlignLen = res.align_len; r->Matches = res.matches; r->Gaps = res.gaps; r->GapRegions = res.gap_regions;
        *out = r;
        return nullptr;
    }

private:
    /* page-locked, grow-only buffers shared by consecutive calls */
    template <class T> struct Pinned {
        T *p = nullptr; size_t cap = 0;
        bool reserve(size_t n) { if (n <= cap) return true; if (p) wfacuda_host_free(p); cap = n + n / 4 + 64; p = (T *)wfacuda_host_alloc(cap * sizeof(T)); if (!p) cap = 0; return p != nullptr; }
        void release() { if (p) wfacuda_host_free(p); p = nullptr; cap = 0; }
    };
    Pinned<uint64_t> qo_, to_, off_, ops_; Pinned<uint32_t> ql_, tl_; Pinned<wfacuda_result> res_;
    std::vector<AlignmentResult> objs_;
    void release_buffers() { qo_.release(); to_.release(); off_.release(); ops_.release(); ql_.release(); tl_.release(); res_.release(); }

    /* f(chunk) for chunk in [c0, c1) on up to 16 host threads */
    template <class F> static void parallel_chunks(size_t c0, size_t c1, size_t threads, F f)
    {
        std::vector<std::thread> th;
        const size_t T = std::max<size_t>(1, std::min(threads, c1 - c0));
        for (size_t k = 1; k < T; k++) th.emplace_back([=] { for (size_t c = c0 + k; c < c1; c += T) f(c); });
        for (size_t c = c0; c < c1; c += T) f(c);
        for (auto &t : th) t.join();
    }

public:
    double last_flatten_ms = 0, last_call_ms = 0, last_objects_ms = 0;      /* where the last AlignBatch spent its time */
private:
    static double now_ms_() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    Error align_batch_on(wfacuda_ctx *const *ctxs, int n_ctx, const std::vector<std::string> &qs, const std::vector<std::string> &ts,
                         std::vector<AlignmentResult *> *results, std::vector<Error> *errors)
    {
        const size_t n = qs.size();
        if (ts.size() != n) { err_ = "AlignBatch: qs and ts differ in length"; return err_.c_str(); }
        /* The per-pair byte strings are scattered over the heap.  They are NOT copied together here:
         * the C ABI takes offsets into one address range, so the range is the heap span of the
         * strings and the offsets are their addresses relative to the lowest one; the library's
         * pipeline workers gather every chunk's sequences into their page-locked pools themselves
         * (wfacuda.h: scattered pools), overlapped with the copies and kernels of the other chunks. */
        const double t_begin = now_ms_();
        const size_t T = n < 65536 ? 1 : std::max<size_t>(1, std::min<size_t>(16, std::thread::hardware_concurrency())), C = T * 2;
        if (!qo_.reserve(n) || !to_.reserve(n) || !ql_.reserve(n) || !tl_.reserve(n) || !res_.reserve(n) || !off_.reserve(n)) { err_ = wfacuda_last_error(nullptr); return err_.c_str(); }
        std::vector<size_t> cut(C + 1), bytes(C, 0);
        std::vector<uintptr_t> lowest(C, ~(uintptr_t)0);
        for (size_t c = 0; c <= C; c++) cut[c] = n * c / C;
        parallel_chunks(0, C, T, [&](size_t c) {
            uintptr_t lo = ~(uintptr_t)0; size_t s_ = 0;
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                lo = std::min(lo, std::min((uintptr_t)qs[i].data(), (uintptr_t)ts[i].data())); s_ += qs[i].size() + ts[i].size();
            }
            lowest[c] = lo; bytes[c] = s_;
        });
        uintptr_t base = ~(uintptr_t)0; size_t total = 0;
        for (size_t c = 0; c < C; c++) { base = std::min(base, lowest[c]); total += bytes[c]; }
        if (n == 0) base = 0;
        if (!ops_.reserve(std::max<size_t>(ops_.cap, total / 4 + 16 * n + 64))) { err_ = wfacuda_last_error(nullptr); return err_.c_str(); }
        parallel_chunks(0, C, T, [&](size_t c) {
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                qo_.p[i] = (uintptr_t)qs[i].data() - base; ql_.p[i] = (uint32_t)qs[i].size();
                to_.p[i] = (uintptr_t)ts[i].data() - base; tl_.p[i] = (uint32_t)ts[i].size();
            }
        });
        last_flatten_ms = now_ms_() - t_begin;
        const double c0 = now_ms_();
        auto call = [&]() {
            return n_ctx == 1 ? wfacuda_align_batch(ctxs[0], n, (const uint8_t *)base, qo_.p, ql_.p, to_.p, tl_.p, res_.p, ops_.p, ops_.cap, off_.p)
                              : wfacuda_align_batch_multi(ctxs, n_ctx, n, (const uint8_t *)base, qo_.p, ql_.p, to_.p, tl_.p, res_.p, ops_.p, ops_.cap, off_.p);
        };
        int rc = call();
        if (rc == WFACUDA_E_OPS_CAPACITY) {
            if (!ops_.reserve(wfacuda_last_ops_total(ctxs[0]))) { err_ = wfacuda_last_error(nullptr); return err_.c_str(); }
            rc = call();
        }
        if (rc != 0) { err_ = wfacuda_last_error(ctxs[0]); return err_.c_str(); }
        last_call_ms = now_ms_() - c0;
        /* result objects from the pool; Ops alias the output buffer (no per-pair allocation) */
        const double o0 = now_ms_();
        if (objs_.size() < n) objs_.resize(n);
        results->resize(n); errors->resize(n);
        parallel_chunks(0, C, T, [&](size_t c) {
            for (size_t i = cut[c]; i < cut[c + 1]; i++) {
                const wfacuda_result &w = res_.p[i];
                (*results)[i] = nullptr; (*errors)[i] = nullptr;
                switch (w.status) {
                case WFACUDA_OK: {
                    AlignmentResult *r = &objs_[i];
                    r->pooled_ = true; r->own_.clear();
                    r->Ops.p = ops_.p + off_.p[i]; r->Ops.n = w.n_ops;
                    r->Score = w.score; r->TBegin = w.tbegin; r->TEnd = w.tend; r->QBegin = w.qbegin; r->QEnd = w.qend;
                    r->AlignLen = w.align_len; r->Matches = w.matches; r->Gaps = w.gaps; r->GapRegions = w.gap_regions;
                    (*results)[i] = r; break;
                }
                case WFACUDA_ERR_EMPTY_SEQ: (*errors)[i] = ErrEmptySeq; break;
                case WFACUDA_ERR_SEQ_TOO_LONG: (*errors)[i] = ErrSeqTooLong; break;
                default: (*errors)[i] = ErrResources;
                }
            }
        });
        last_objects_ms = now_ms_() - o0;
        return nullptr;
    }

    wfacuda_config config() const
    {
        wfacuda_config c{};
        c.mismatch = p_.Mismatch; c.gap_open = p_.GapOpen; c.gap_ext = p_.GapExt;
        c.global_alignment = opt_.GlobalAlignment ? 1 : 0;
        c.adaptive = has_ad_ ? 1 : 0;
        if (has_ad_) { c.min_wf_len = ad_.MinWFLen; c.max_dist_diff = ad_.MaxDistDiff; c.cutoff_step = ad_.CutoffStep; }
        return c;
    }
    Penalties p_; Options opt_; AdaptiveReductionOption ad_{}; bool has_ad_ = false;
    wfacuda_ctx *ctx_ = nullptr; std::string err_;
};

/* wfa.go:120-131 (ad is reset, unlike the reference's pooled Aligner) / :102-116 / wfa_cigar.go:92-96 */
inline Aligner *New(const Penalties *p, const Options *opt, int device = 0) { return new Aligner(p, opt, device); }
inline void RecycleAligner(Aligner *a) { delete a; }
inline void RecycleAlignmentResult(AlignmentResult *r) { if (r && !r->pooled_) delete r; }      /* pooled results go back with their Aligner */

} // namespace wfa
#endif
