/*
 * wfacuda.cu -- host driver + C ABI of libwfacuda.so (see include/wfacuda.h).
 *
 * Flow of one batch (wfacuda_align_batch = upload + run + download):
 *   upload    validate lengths (wfa.go:202-209), stage the byte pool and the
 *             pair descriptors into HBM through pinned double buffers, bin
 *             pairs by cost (counting sort, longest first)
 *   run       pack_kernel (bytes -> 2-bit, flags non-ACGT pairs), then the
 *             persistent align kernels: WARP class (one warp per pair) and CTA
 *             class (one block per pair); pairs that ran out of ring width,
 *             arena or ops pool are re-queued with more of it -- still on the
 *             GPU, there is no CPU path
 *   download  D2H of the result records, each pair's position in the ops pool and the pool's used prefix
 */
#include "../../include/wfacuda.h"
#include "wfa_kernels.cuh"
#include "wfa_lane.cuh"
#include "wfa_slim.cuh"
#include "wfa_wide.cuh"
#include "wfa_render.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace wfak;


static_assert(sizeof(Result) == sizeof(wfacuda_result), "result layout");
static_assert(sizeof(RowHdr) == 24, "row header layout");

namespace {

thread_local std::string g_tls_error;
double g_dbg_t0 = 0.0;          /* WFACUDA_DEBUG: start of the current wfacuda_align_batch call */

struct DevBuf { void *p = nullptr; size_t cap = 0; };

} // namespace

/* What a pipeline worker sends per pair instead of a 40-byte PairDesc: offsets relative to the
 * chunk (32 bits are enough below 4 GB of sequence / 2^32 packed words per chunk); t_word follows
 * from q_word and n.  expand_descs_kernel turns them into PairDescs on the device -- the e2e path
 * is bound by host-to-device bytes, and this is 20 MB less per million pairs. */
struct WireDesc { uint32_t q_byte, t_byte, q_word, n, m; };

namespace {
__global__ void expand_descs_kernel(const WireDesc *__restrict__ w, uint32_t n_pairs, PairDesc *__restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
        const WireDesc d = w[i];
        PairDesc o;
        o.q_byte = d.q_byte; o.t_byte = d.t_byte; o.q_word = d.q_word;
        o.t_word = (uint64_t)d.q_word + ((((uint64_t)d.n + 15) >> 4) + 3 & ~3ull);
        o.n = d.n; o.m = d.m;
        if (d.n == 0) { o.q_byte = o.t_byte = o.q_word = o.t_word = 0; o.m = 0; }
        out[i] = o;
    }
}
} // namespace

namespace {
/* INT32 issue microbenchmark ("INT32 peak must be measured on the box"): eight
 * independent chains per thread, mode 0 = add and xor (xor issues on the ALU pipe; ptxas turns part
 * of the adds into IMAD.IADD on the FMA pipe by itself), mode 1 = add alternating with mad.lo
 * (ALU pipe + FMA pipe).  Inline PTX, every op mixing in a neighbouring chain, so that nothing is
 * folded away: 32 ops per thread and iteration in the SASS. */
template <int mode>
__global__ void __launch_bounds__(256) int32_peak_kernel(uint32_t *out, int iters)
{
    uint32_t a[8];
    const uint32_t b = threadIdx.x * 2654435761u + 12345u, c = blockIdx.x | 1u;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = b + (uint32_t)j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            /* every op mixes in a neighbouring chain, so that ptxas cannot fold repeated steps */
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (mode == 0) {
                    if (u & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(a[(j + 1) & 7]));
                    else       asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[j]) : "r"(a[(j + 3) & 7]));
                } else {
                    if (u & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(a[(j + 1) & 7]));
                    else       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(a[(j + 3) & 7]), "r"(c));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
} // namespace

/* SLIM worker (wfa_slim.cuh): one instantiation per (passes per row, size class, wf-adaptive) */
namespace {
constexpr int kSlimLadder[] = {4, 8};           /* MAXP: rows of up to 128 / 256 diagonals */
constexpr int kSlimLadderN = 2;
typedef void (*SlimKernel)(const KParams);
template <int SZ, bool ADAPT> SlimKernel slim_kernel_sel(int maxp)
{
    switch (maxp) {
    case 4: return slim_kernel<4, SZ, ADAPT>;
    case 8: return slim_kernel<8, SZ, ADAPT>;
    }
    return nullptr;
}
SlimKernel slim_kernel_ptr(int maxp, int sz, bool adapt)
{
    switch (sz) {
    case 0: return adapt ? slim_kernel_sel<0, true>(maxp) : slim_kernel_sel<0, false>(maxp);
    case 1: return adapt ? slim_kernel_sel<1, true>(maxp) : slim_kernel_sel<1, false>(maxp);
    default: return adapt ? slim_kernel_sel<2, true>(maxp) : slim_kernel_sel<2, false>(maxp);
    }
}
int slim_size_class(uint32_t max_len) { return max_len <= SLIM_MAX_M10 ? 0 : max_len <= SLIM_MAX_SHORT ? 1 : 2; }
} // namespace

struct wfacuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;     /* the stream work is issued on (normally stream_main) */
    cudaStream_t stream_main = nullptr;
    /* High-priority twin: once a batch's big kernel is done, its follow-ups (hand-over of the few
     * pairs the LANE worker could not hold, retries, D2H) go here so that they do not queue behind
     * the thousands of pending blocks of other pipeline workers' kernels. */
    cudaStream_t stream_hi = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    wfacuda_config cfg{};
    uint32_t g = 1; int xg = 0, oeg = 0, eg = 0, dM = 0, dE = 0;
    int sm_count = 0; size_t smem_optin = 0; size_t total_mem = 0;
    /* grow-only device buffers shared by consecutive batches */
    DevBuf arena, retry, work, ctr, ops_pool;
    std::vector<std::pair<void *, size_t>> free_dev;     /* returned batch buffers */
    void *pinned[2] = {nullptr, nullptr}; size_t pinned_cap = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    double arena_scale = 1.0;      /* learned: observed / estimated arena need */
    double arena_scale_slim = 1.0;  /* the same for the REG worker's slots */
    double arena_scale_wide = 1.0;  /* the same for the WIDE worker's slots */
    DevBuf wide_rec;                /* WIDE class: FwdOut of every item of a launch, read by its finish kernel */
    int slim_p_learned = 0;         /* learned: cells per lane (row capacity / 32) the REG worker needed */
    int slim_occ[3][2][9] = {};    /* [size class][adaptive][MAXP]: resident blocks per SM, 0 = unknown */
    /* LANE class: sampled histogram of the score index at which the previous batch's pairs
     * finished, and the stage boundaries taken from it */
    uint64_t lane_hist[64] = {}; uint64_t lane_hist_n = 0;
    int lane_bounds[3] = {0, 0, 0}; int lane_n_bounds = 0;
    int lane_occ[3] = {0, 0, 0}, lane_occ_sw = 0;   /* LANE kernel: resident blocks per SM for lane_occ_sw words per sequence, rings of 64 / 56 / 48 columns */
    int ring_cap_learned = 0;      /* learned: ring width that the WARP class needed */
    int occ_cache[2][2][8] = {};   /* [cta][bits==8][log2(ring_cap/64)+1]: blocks per SM, 0 = unknown */
    uint64_t budget_cache = 0;     /* arena budget; refreshed when the arena has to grow */
    wfacuda_stats stats{};
    uint64_t last_ops_total = 0;
    const wfacuda_batch *pool_owner = nullptr;   /* batch whose ops the pool currently holds */
    int last_rc = 0;
    std::string err;
    std::vector<wfacuda_ctx *> subs;   /* pipeline workers of wfacuda_align_batch (same device) */
    /* Pipeline workers take turns on the H2D copy engine: copies issued from several streams at
     * once are served round-robin, so every chunk would arrive late and no kernel could start
     * before most of the batch is across PCIe; one chunk at a time keeps arrival FIFO. */
    struct Turns {                     /* counting semaphore, optionally served in ticket (= chunk) order */
        std::mutex mu; std::condition_variable cv; int free_slots = 1; uint64_t next_ticket = 0;
        void acquire() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return free_slots > 0; }); free_slots--; }
        /* uploads go out in chunk order, so that the small chunks at the end of a batch really are
         * the last ones across PCIe (a worker that is free early would otherwise overtake) */
        void acquire_ordered(uint64_t ticket)
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return free_slots > 0 && next_ticket >= ticket; });
            free_slots--; if (next_ticket == ticket) next_ticket = ticket + 1;
            lk.unlock(); cv.notify_all();
        }
        void pass(uint64_t ticket) { { std::lock_guard<std::mutex> lk(mu); if (next_ticket <= ticket) next_ticket = ticket + 1; } cv.notify_all(); }
        void release() { { std::lock_guard<std::mutex> lk(mu); free_slots++; } cv.notify_all(); }
    } h2d_turns;                       /* owned by the parent ctx; one copy at a time measured best (WFACUDA_H2D_TURNS) */
    Turns *h2d_turn = nullptr;         /* set in a worker ctx: the parent's semaphore */
    int64_t h2d_ticket = -1;           /* worker: index of the chunk it is about to upload (-1: unordered) */
    /* The pipeline workers' sequence uploads all go through ONE stream of the parent ctx: copies
     * queued on one stream run back to back in FIFO order with no host round trip between them
     * (the semaphore above left ~50 us of PCIe idle per chunk: sync wake-up, hand-over, launch);
     * the worker's own stream waits for its copy through an event. */
    cudaStream_t h2d_fifo = nullptr;   /* parent: the shared upload stream */
    std::mutex h2d_mu;                 /* parent: serialises the enqueue + event record */
    wfacuda_ctx *h2d_parent = nullptr; /* worker: whose h2d_fifo to use */
    cudaEvent_t ev_h2d = nullptr;      /* worker: completion of its chunk's sequence upload */
    WireDesc *pin_descs = nullptr; size_t pin_descs_cap = 0;   /* worker: page-locked wire descriptors, DMA'd without a staging copy */
    /* Host waits.  cudaStreamSynchronize spins on a core; a pipeline worker that shares few cores with
     * many other workers (several ranks on one box) waits on an event created with
     * cudaEventBlockingSync instead and sleeps. */
    bool blocking_sync = false; cudaEvent_t ev_block = nullptr;
    unsigned core_share = 0;           /* host cores this ctx's pipeline may count on (0: all of them) */
    /* the counters as last fetched from the device, valid while nothing was launched since (saves
     * the second read-back + wait at the end of a run whose last class already fetched them) */
    Counters hc_cache{}; bool hc_cache_valid = false;
    uint8_t *pin_pool = nullptr; size_t pin_pool_cap = 0;      /* page-locked copy of a batch's sequences when they are scattered over a much larger pool */
    DevBuf wire_dev;                                           /* their landing place on the device */
    DevBuf render_meta, render_cigar, render_text;             /* wfacuda_batch_render: offsets / lengths / cursors, strings */
    uint64_t render_cigar_total = 0, render_text_total = 0;
    /* wfacuda_align_components: one pair on one worker, whose slot is read back afterwards */
    bool dump_mode = false, dump_cta = false; uint64_t dump_slot_bytes = 0, dump_rows = 0;
    uint64_t lane_handed = 0;          /* pairs of the last run that the LANE class handed to the WARP kernel on the device */
    const wfacuda_batch *pin_descs_owner = nullptr;
};

struct wfacuda_batch {
    uint64_t n_pairs = 0;
    std::vector<uint8_t> host_status;       /* EMPTY / TOO_LONG decided on the host */
    std::vector<uint32_t> order_warp, order_cta, order_lane;
    std::vector<uint32_t> order_slim_small, order_slim_rest;   /* order_warp split by SLIM cell word (reads of about 1 kbp), see wfacuda_batch_run */
    int identity_cls = -1;                  /* class (0 warp, 1 cta, 2 lane) whose order is 0..n-1: no work list needed */
    uint32_t lane_maxlen = 1;               /* longest sequence of the LANE class */
    PairDesc *descs = nullptr;              /* descs_own's storage; nullptr when the batch was uploaded with wire descriptors */
    std::vector<PairDesc> descs_own;
    const WireDesc *wire = nullptr;         /* pipeline workers: the ctx's page-locked wire descriptors (expanded on the device) */
    uint32_t n_of(uint64_t i) const { return wire ? wire[i].n : descs[i].n; }
    uint32_t m_of(uint64_t i) const { return wire ? wire[i].m : descs[i].m; }
    uint64_t raw_bytes = 0, packed_words = 0, seq_bases = 0, max_nm = 0;
    uint32_t max_len = 0;                   /* longest single sequence */
    void *d_raw = nullptr, *d_packed = nullptr, *d_descs = nullptr, *d_flags = nullptr;
    void *d_results = nullptr, *d_where = nullptr;
    size_t sz_raw = 0, sz_packed = 0, sz_descs = 0, sz_flags = 0, sz_results = 0, sz_where = 0;
    uint64_t ops_total = 0, n_invalid = 0;
    bool ran = false;
};

namespace {

int fail(wfacuda_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (ctx) { ctx->err = buf; ctx->last_rc = code; }
    g_tls_error = buf;
    return code;
}

#define CU(ctx, call)                                                                            \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? WFACUDA_E_NOMEM : WFACUDA_E_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

cudaError_t wait_stream(wfacuda_ctx *ctx, cudaStream_t st)
{
    if (!ctx->blocking_sync) return cudaStreamSynchronize(st);
    if (!ctx->ev_block) { cudaError_t e = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    cudaError_t e = cudaEventRecord(ctx->ev_block, st);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->ev_block);
}

double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6; }


/* Device fills run as kernels, not cudaMemsetAsync: a memset may be executed by a copy engine,
 * where it queues behind every upload the pipeline workers have in flight. */
__global__ void fill_kernel(uint4 *p16, size_t n16, unsigned char *p1, size_t n1, unsigned int byte)
{
    const unsigned int w = byte * 0x01010101u;
    const uint4 v = make_uint4(w, w, w, w);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p16[i] = v;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n1; i += (size_t)gridDim.x * blockDim.x) p1[i] = (unsigned char)byte;
}
thread_local uint32_t g_fill_launches = 0;     /* fill kernels launched by this thread (counted into stats.kernel_launches per run) */
cudaError_t dev_fill(void *p, int byte, size_t bytes, cudaStream_t stream)
{
    if (!bytes) return cudaSuccess;
    unsigned char *c = (unsigned char *)p;
    size_t head = (16 - ((uintptr_t)c & 15)) & 15; if (head > bytes) head = bytes;
    /* bytes before the first 16-byte boundary and after the last one go through the byte loop */
    const size_t n16 = (bytes - head) / 16, tail = bytes - head - n16 * 16;
    if (head) { fill_kernel<<<1, 32, 0, stream>>>(nullptr, 0, c, head, (unsigned int)(byte & 255)); g_fill_launches++; }
    if (n16 || tail) {
        const int blocks = (int)std::min<size_t>(std::max<size_t>((n16 + 255) / 256, 1), 1184);
        fill_kernel<<<blocks, 256, 0, stream>>>((uint4 *)(c + head), n16, c + head + n16 * 16, tail, (unsigned int)(byte & 255));
        g_fill_launches++;
    }
    return cudaGetLastError();
}

uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { uint32_t t = a % b; a = b; b = t; } return a; }

int apply_config(wfacuda_ctx *ctx, const wfacuda_config *cfg)
{
    if (!cfg) return fail(ctx, WFACUDA_E_INVALID, "config is NULL");
    if (cfg->mismatch == 0 || cfg->gap_ext == 0)
        return fail(ctx, WFACUDA_E_INVALID, "mismatch and gap_ext penalties must be > 0 (the reference's recurrences read the score being written otherwise)");
    if (cfg->adaptive && cfg->min_wf_len == 0)      /* wfa.go:135-137 */
        return fail(ctx, WFACUDA_E_INVALID, "cutoff step should not be 0");
    const uint64_t oe = (uint64_t)cfg->gap_open + cfg->gap_ext;
    if (cfg->mismatch > (1u << 20) || oe > (1u << 20))
        return fail(ctx, WFACUDA_E_INVALID, "penalties above 2^20 are not supported");
    const uint32_t g = gcd_u32(gcd_u32(cfg->mismatch, (uint32_t)oe), cfg->gap_ext);
    const int xg = (int)(cfg->mismatch / g), oeg = (int)(oe / g), eg = (int)(cfg->gap_ext / g);
    const int dM = std::max(xg, oeg) + 1;
    if (dM > 1024) return fail(ctx, WFACUDA_E_INVALID, "max(mismatch, gap_open+gap_ext)/gcd must be <= 1023");
    ctx->cfg = *cfg; ctx->g = g; ctx->xg = xg; ctx->oeg = oeg; ctx->eg = eg; ctx->dM = dM; ctx->dE = eg + 1;
    return 0;
}

int ensure(wfacuda_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    bytes = (bytes + 255) & ~(size_t)255;
    const double t0 = now_ms();
    CU(ctx, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    if (getenv("WFACUDA_DEBUG_ALLOC")) fprintf(stderr, "[wfacuda] ensure: cudaMalloc(%.1f MB) took %.2f ms\n", bytes / 1e6, now_ms() - t0);
    return 0;
}

/* batch buffers are recycled through the ctx so that steady-state batches do no cudaMalloc */
int dev_take(wfacuda_ctx *ctx, void **p, size_t *sz, size_t bytes)
{
    bytes = (std::max<size_t>(bytes, 256) + 255) & ~(size_t)255;
    int best = -1;
    for (size_t i = 0; i < ctx->free_dev.size(); i++)
        if (ctx->free_dev[i].second >= bytes && ctx->free_dev[i].second <= 2 * bytes + (1 << 20) &&
            (best < 0 || ctx->free_dev[i].second < ctx->free_dev[best].second)) best = (int)i;
    if (best >= 0) {
        *p = ctx->free_dev[best].first; *sz = ctx->free_dev[best].second;
        ctx->free_dev.erase(ctx->free_dev.begin() + best);
        return 0;
    }
    const double t0 = now_ms();
    cudaError_t e = cudaMalloc(p, bytes);
    if (getenv("WFACUDA_DEBUG_ALLOC")) fprintf(stderr, "[wfacuda] dev_take: cudaMalloc(%.1f MB) took %.2f ms (cache holds %zu)\n", bytes / 1e6, now_ms() - t0, ctx->free_dev.size());
    if (e != cudaSuccess) {
        /* drop the cache and retry once */
        for (auto &f : ctx->free_dev) cudaFree(f.first);
        ctx->free_dev.clear();
        cudaGetLastError();
        CU(ctx, cudaMalloc(p, bytes));
    }
    *sz = bytes;
    return 0;
}
void dev_give(wfacuda_ctx *ctx, void **p, size_t *sz)
{
    if (*p) ctx->free_dev.emplace_back(*p, *sz);
    *p = nullptr; *sz = 0;
}

/* true when the host range starts in page-locked memory known to CUDA (wfacuda_host_alloc,
 * cudaHostAlloc, cudaHostRegister): the DMA engine can then read / write it directly */
bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

/* H2D of a host region: straight from the caller's memory when it is page-locked, else
 * through two pinned staging buffers (host copy overlapped with the DMA) */
int staged_h2d(wfacuda_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes >= 65536 && is_pinned(src)) {
        CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.h2d_bytes += bytes;
        return 0;
    }
    size_t done = 0; int which = 0;
    while (done < bytes) {
        const size_t chunk = std::min(ctx->pinned_cap, bytes - done);
        CU(ctx, cudaEventSynchronize(ctx->pin_ev[which]));
        memcpy(ctx->pinned[which], (const char *)src + done, chunk);
        CU(ctx, cudaMemcpyAsync((char *)dst + done, ctx->pinned[which], chunk, cudaMemcpyHostToDevice, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->pin_ev[which], ctx->stream));
        done += chunk; which ^= 1;
    }
    ctx->stats.h2d_bytes += bytes;
    return 0;
}

/* D2H into a host (pageable) region through the two pinned buffers (double buffered) */
int staged_d2h(wfacuda_ctx *ctx, void *dst, const void *src, size_t bytes, bool defer_sync = false)
{
    if (bytes >= 65536 && is_pinned(dst)) {
        CU(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (!defer_sync) CU(ctx, wait_stream(ctx, ctx->stream));
        ctx->stats.d2h_bytes += bytes;
        return 0;
    }
    size_t issued = 0, done = 0, len[2] = {0, 0}, at[2] = {0, 0};
    auto issue = [&](int w) -> cudaError_t {
        const size_t chunk = std::min(ctx->pinned_cap, bytes - issued);
        cudaError_t e = cudaMemcpyAsync(ctx->pinned[w], (const char *)src + issued, chunk, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->pin_ev[w], ctx->stream);
        len[w] = chunk; at[w] = issued; issued += chunk;
        return e;
    };
    if (bytes) CU(ctx, issue(0));
    if (issued < bytes) CU(ctx, issue(1));
    for (int w = 0; done < bytes; w ^= 1) {
        CU(ctx, cudaEventSynchronize(ctx->pin_ev[w]));
        memcpy((char *)dst + at[w], ctx->pinned[w], len[w]);
        done += len[w];
        if (issued < bytes) CU(ctx, issue(w));
    }
    ctx->stats.d2h_bytes += bytes;
    return 0;
}

/* D2H of a small device region on the ctx's own stream, through the pinned staging buffer.
 * (A plain cudaMemcpy runs on the legacy default stream; with several pipeline workers it was
 * seen to wait for the other workers' queued kernels.) */
int fetch_small(wfacuda_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    size_t done = 0;
    while (done < bytes) {
        const size_t chunk = std::min(ctx->pinned_cap, bytes - done);
        CU(ctx, cudaMemcpyAsync(ctx->pinned[0], (const char *)src + done, chunk, cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, wait_stream(ctx, ctx->stream));
        memcpy((char *)dst + done, ctx->pinned[0], chunk);
        done += chunk;
    }
    return 0;
}

/* ---- planning ------------------------------------------------------------ */

struct Need { uint64_t arena; int width; };

/* Rough per-pair need before anything is known about the error rate.  Pairs
 * that overflow are re-queued with 4x, and ctx->arena_scale learns the ratio
 * observed/estimated from every batch, so only the first batch pays. */
Need estimate(const wfacuda_ctx *ctx, uint32_t n, uint32_t m)
{
    const wfacuda_config &c = ctx->cfg;
    const double L = std::min(n, m), diag = (double)n + m - 1;
    const double oe = (double)c.gap_open + c.gap_ext;
    /* score guess: ~10 % edits at the mean edit cost, plus the end gap */
    const double per_edit = (c.mismatch + 2.0 * oe) / 3.0;
    const double score = 0.10 * L * per_edit + oe + std::abs((double)m - n) * c.gap_ext + 4.0 * c.mismatch;
    const double rows = score / ctx->g + 2;
    double width_final;
    if (!c.global_alignment) width_final = diag;
    else {
        width_final = std::min(diag, 2.0 * score / c.gap_ext + 3);
        if (c.adaptive) width_final = std::min(width_final, 1.5 * c.max_dist_diff + c.min_wf_len + 16.0);
    }
    const double avg_width = (!c.global_alignment || c.adaptive) ? width_final : 0.55 * width_final + 3;
    double bytes = rows * (avg_width * 12.0 + sizeof(RowHdr)) + (n + m) * 0.5 + 4096;
    bytes *= 1.3 * ctx->arena_scale;
    Need nd; nd.arena = (uint64_t)bytes; nd.width = (int)std::min(diag, width_final);
    return nd;
}

/* The same for a SLIM slot: one 4- or 8-byte word per cell, 16-byte row headers */
Need estimate_slim(const wfacuda_ctx *ctx, uint32_t n, uint32_t m, bool wide)
{
    const wfacuda_config &c = ctx->cfg;
    const double L = std::min(n, m), diag = (double)n + m - 1;
    const double oe = (double)c.gap_open + c.gap_ext;
    const double per_edit = (c.mismatch + 2.0 * oe) / 3.0;
    const double score = 0.10 * L * per_edit + oe + std::abs((double)m - n) * c.gap_ext + 4.0 * c.mismatch;
    const double rows = score / ctx->g + 2;
    double width_final = std::min(diag, 2.0 * score / c.gap_ext + 3);
    if (c.adaptive) width_final = std::min(width_final, 1.5 * c.max_dist_diff + c.min_wf_len + 16.0);
    const double avg_width = c.adaptive ? width_final : 0.55 * width_final + 3;
    double bytes = rows * (avg_width * (wide ? 8.0 : 4.0) + sizeof(SlimHdr)) + 4096;
    bytes *= 1.3 * ctx->arena_scale_slim;
    Need nd; nd.arena = (uint64_t)bytes; nd.width = (int)std::min(diag, width_final);
    return nd;
}

struct LaunchPlan {
    bool cta; int threads; int blocks; int ring_cap; int seq_cap; size_t smem; uint64_t slot_bytes; uint64_t workers; bool slot_at_max;
    int group;               /* WARP class: pairs per warp group; slot_bytes is per pair, a warp owns group * slot_bytes */
};

uint64_t arena_budget(wfacuda_ctx *ctx, bool refresh = true)
{
    if (!refresh && ctx->budget_cache) return ctx->budget_cache;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 1ull << 30; }
    const uint64_t avail = (uint64_t)free_b + ctx->arena.cap;
    uint64_t budget = (uint64_t)(avail * 0.85);
    if (ctx->cfg.arena_budget_bytes) budget = std::min<uint64_t>(budget, ctx->cfg.arena_budget_bytes);
    ctx->budget_cache = std::max<uint64_t>(budget, 1u << 20);
    return ctx->budget_cache;
}

/* choose worker count / slot size for one class launch */
int plan_launch(wfacuda_ctx *ctx, const wfacuda_batch *b, const std::vector<uint32_t> &order, bool cta, int bits,
                double boost, int min_ring_cap, LaunchPlan *lp, int slim_p = 0, int slim_sz = 0)
{
    uint64_t need_max = 0; int width_max = 1; uint32_t seq_words_max = 0;
    /* the list is sorted longest first; a prefix sample bounds the estimate cheaply */
    const size_t sample = std::min<size_t>(order.size(), 4096);
    for (size_t i = 0; i < sample; i++) {
        const uint32_t dn = b->n_of(order[i]), dm = b->m_of(order[i]);
        const Need nd = slim_p ? estimate_slim(ctx, dn, dm, slim_sz != 0) : estimate(ctx, dn, dm);
        need_max = std::max(need_max, nd.arena); width_max = std::max(width_max, nd.width);
        seq_words_max = std::max(seq_words_max, ((dn + 15) >> 4) + ((dm + 15) >> 4) + 2);
    }
    /* WARP worker, 2-bit: pairs of up to ~2 kbp keep their packed sequences in shared memory
     * (the kernel checks every pair against this capacity and reads longer ones from global) */
    lp->seq_cap = (!cta && !slim_p && bits == 2 && seq_words_max <= 264 && !getenv("WFACUDA_NO_SEQ_SMEM")) ? (int)((seq_words_max + 3) & ~3u) : 0;
    need_max = (uint64_t)((double)need_max * boost);
    lp->cta = cta; lp->slot_at_max = false;
    /* the device is only asked for its free memory when the arena may have to grow */
    uint64_t slot = (std::max<uint64_t>(need_max, 16384) + 255) & ~255ull;
    const uint64_t budget = arena_budget(ctx, ctx->arena.cap == 0 || boost > 1.0);
    int wpb, blocks_per_sm = 1;
    if (slim_p) {
        wpb = 4; lp->threads = 128; lp->ring_cap = 32 * slim_p; lp->smem = slim_smem_pad(slim_p) + slim_smem_bytes(slim_p) * 4;
        int &oc = ctx->slim_occ[slim_sz][ctx->cfg.adaptive ? 1 : 0][slim_p];
        if (!oc && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, slim_kernel_ptr(slim_p, slim_sz, ctx->cfg.adaptive != 0), lp->threads, lp->smem) != cudaSuccess) { cudaGetLastError(); oc = 1; }
        blocks_per_sm = std::max(1, oc);
    } else if (cta) {
        wpb = 1; lp->threads = WFA_CTA_THREADS; lp->ring_cap = 0;
        lp->smem = worker_smem_bytes<true>(ctx->dM, ctx->dE, 0);
        int &oc = ctx->occ_cache[1][bits == 8][0];
        if (!oc && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, bits == 2 ? align_kernel<2, true> : align_kernel<8, true>, lp->threads, lp->smem) != cudaSuccess) { cudaGetLastError(); oc = 1; }
        blocks_per_sm = oc;
        blocks_per_sm = std::max(1, std::min(blocks_per_sm, 4));
    } else {
        wpb = 4;
        int cap = 64;
        while (cap < (int)(width_max * 0.33) + 8 && cap < 512) cap *= 2;
        cap = std::max(cap, std::max(min_ring_cap, ctx->ring_cap_learned));
        size_t per_warp = worker_smem_bytes<false>(ctx->dM, ctx->dE, cap, lp->seq_cap);
        while (per_warp * wpb > ctx->smem_optin && cap > 32) { cap /= 2; per_warp = worker_smem_bytes<false>(ctx->dM, ctx->dE, cap, lp->seq_cap); }
        if (per_warp * wpb > ctx->smem_optin) return fail(ctx, WFACUDA_E_INVALID, "penalties need a deeper shared-memory ring than fits");
        lp->ring_cap = cap; lp->threads = wpb * 32; lp->smem = per_warp * wpb;
        int ci = 1; for (int c = 64; c < cap && ci < 7; c *= 2) ci++;
        int oc_seq = 0;
        int &oc = lp->seq_cap ? oc_seq : ctx->occ_cache[0][bits == 8][cap < 64 ? 0 : ci];
        if (!oc && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, bits == 2 ? align_kernel<2, false> : align_kernel<8, false>, lp->threads, lp->smem) != cudaSuccess) { cudaGetLastError(); oc = 1; }
        blocks_per_sm = oc;
        blocks_per_sm = std::max(1, blocks_per_sm);
    }
    uint64_t workers = (uint64_t)ctx->sm_count * blocks_per_sm * wpb;
    workers = std::min<uint64_t>(workers, ((order.size() + wpb - 1) / wpb) * wpb);
    if (slot * workers > budget) workers = std::max<uint64_t>(wpb, (budget / slot) / wpb * wpb);
    if (slot * workers > budget) { slot = (budget / workers) & ~255ull; lp->slot_at_max = true; }
    const uint64_t kWarpSlotMax = 15ull << 30;      /* the WARP worker indexes its slot with 32-bit word offsets */
    if (!cta && slot > kWarpSlotMax) { slot = kWarpSlotMax; lp->slot_at_max = true; }
    lp->group = 1;
    if (!cta) {
        /* as many pairs per group as the budget allows (lane-parallel backtrace), at most 32,
         * and no more than keeps every warp busy */
        uint64_t g = std::min<uint64_t>(32, budget / std::max<uint64_t>(1, slot * workers));
        g = std::min<uint64_t>(g, std::max<uint64_t>(1, order.size() / std::max<uint64_t>(1, workers)));
        if (const char *e = getenv("WFACUDA_GROUP")) g = std::min<uint64_t>(g, (uint64_t)std::max(1, atoi(e)));
        lp->group = (int)std::max<uint64_t>(1, g);
    }
    if (ctx->dump_mode) { workers = wpb; lp->group = 1; }            /* one block; the WARP kernel lets only its warp 0 work */
    lp->blocks = (int)(workers / wpb); lp->workers = workers; lp->slot_bytes = slot;
    return 0;
}

/* Runs one class of pairs to completion, re-queuing pairs that ran out of
 * ring width (WARP kernel -> wider ring or the CTA kernel through *to_cta),
 * arena (4x slot) or ops pool (pool doubled). */
int run_class(wfacuda_ctx *ctx, wfacuda_batch *b, const std::vector<uint32_t> &order0, bool identity, bool cta, int bits, KParams base,
              std::vector<uint32_t> *to_cta, std::vector<uint32_t> *to_8bit, bool slim = false, int slim_sz_force = -1)
{
    /* SLIM worker: `to_cta` takes the pairs whose rows outgrow its widest instantiation (the WARP
     * worker places them); passes per row start from what earlier batches needed */
    const int slim_sz = slim_sz_force >= 0 ? slim_sz_force : slim_size_class(b->max_len);
    int slim_li = 0;
    if (slim) {
        int p0 = ctx->slim_p_learned;
        if (!p0) {
            /* wf-adaptive: rows hover around MaxDistDiff + a few dozen diagonals; without heuristic
             * they grow by two per score, so the score guess decides */
            if (ctx->cfg.adaptive) p0 = (int)std::min<uint64_t>(8, ((uint64_t)ctx->cfg.max_dist_diff + 60 + 31) / 32);
            else p0 = std::min(8, (estimate_slim(ctx, b->max_len, b->max_len, slim_sz != 0).width + 31) / 32);
        }
        if (const char *e = getenv("WFACUDA_SLIM_P")) p0 = atoi(e);
        while (slim_li + 1 < kSlimLadderN && kSlimLadder[slim_li] < p0) slim_li++;
    }
    double &scale = slim ? ctx->arena_scale_slim : ctx->arena_scale;
    double boost = 1.0; int min_cap = 0;
    std::vector<uint32_t> requeued;
    for (int attempt = 0; ; attempt++) {
        const std::vector<uint32_t> &order = attempt == 0 ? order0 : requeued;
        if (order.empty()) break;
        const bool ident = identity && attempt == 0;      /* pair index == queue position: no work list */
        if (attempt > 24) return fail(ctx, WFACUDA_E_NOMEM, "pairs still out of resources after %d retries", attempt);
        LaunchPlan lp;
        const int slim_p = slim ? kSlimLadder[slim_li] : 0;
        int rc = plan_launch(ctx, b, order, cta, bits, boost, min_cap, &lp, slim_p, slim_sz);
        if (rc) return rc;
        if ((rc = ensure(ctx, ctx->arena, lp.slot_bytes * lp.group * lp.workers))) return rc;
        if (!ident && (rc = ensure(ctx, ctx->work, order.size() * 4))) return rc;
        if ((rc = ensure(ctx, ctx->retry, order.size() * 8 + 16))) return rc;
        if (!ident) { int rc2 = staged_h2d(ctx, ctx->work.p, order.data(), order.size() * 4); if (rc2) return rc2; }
        /* reset queue + retry counters, keep the work counters and the ops cursor */
        Counters *dc = (Counters *)ctx->ctr.p;
        CU(ctx, dev_fill(&dc->retry_n, 0, (size_t)((char *)&dc->launch_end - (char *)&dc->retry_n), ctx->stream));   /* the per-launch block */
        KParams P = base;
        P.work = ident ? nullptr : (const uint32_t *)ctx->work.p; P.n_work = (uint32_t)order.size();
        P.arena = (uint8_t *)ctx->arena.p; P.slot_bytes = lp.slot_bytes * lp.group; P.group = lp.group;
        P.retry = (uint64_t *)ctx->retry.p; P.ctr = dc; P.retry_ctr = &dc->retry_n;
        P.ring_cap = lp.ring_cap; P.seq_cap = lp.seq_cap; P.ops_pool = (uint64_t *)ctx->ops_pool.p; P.ops_cap = ctx->ops_pool.cap / 8;
        if (ctx->dump_mode) { P.single_worker = 1; P.semi_literal = 1; ctx->dump_cta = cta; ctx->dump_slot_bytes = lp.slot_bytes * lp.group; }
        const double tk0 = now_ms();
        if (slim) slim_kernel_ptr(slim_p, slim_sz, ctx->cfg.adaptive != 0)<<<lp.blocks, lp.threads, lp.smem, ctx->stream>>>(P);
        else if (cta) { if (bits == 2) align_kernel<2, true><<<lp.blocks, lp.threads, lp.smem, ctx->stream>>>(P);
                   else           align_kernel<8, true><<<lp.blocks, lp.threads, lp.smem, ctx->stream>>>(P); }
        else     { if (bits == 2) align_kernel<2, false><<<lp.blocks, lp.threads, lp.smem, ctx->stream>>>(P);
                   else           align_kernel<8, false><<<lp.blocks, lp.threads, lp.smem, ctx->stream>>>(P); }
        CU(ctx, cudaGetLastError());
        ctx->stats.kernel_launches++; ctx->stats.align_launches++; ctx->hc_cache_valid = false;
        ctx->stats.arena_bytes = std::max<uint64_t>(ctx->stats.arena_bytes, lp.slot_bytes * lp.group * lp.workers);
        Counters hc;
        { int rc2 = fetch_small(ctx, &hc, dc, sizeof hc); if (rc2) return rc2; }
        ctx->hc_cache = hc; ctx->hc_cache_valid = true;
        if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   %s kernel: host launch at %.2f, sync returned %.2f ms since call\n", cta ? "cta" : "warp", tk0 - g_dbg_t0, now_ms() - g_dbg_t0);
        if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   launch %s bits=%d attempt %d: %zu pairs, %d blocks x %d thr, ring_cap %d, slim passes %d, group %d, smem %zu, slot %.1f KB (scale %.3f), used max %.1f KB, retry %llu\n", slim ? "slim" : cta ? "cta" : "warp", bits, attempt, order.size(), lp.blocks, lp.threads, lp.ring_cap, slim_p, lp.group, lp.smem, lp.slot_bytes / 1024.0, scale, hc.arena_used_max / 1024.0, (unsigned long long)hc.retry_n);
        if (hc.retry_n == 0) {
            /* learn: aim the next batch's slots at 1.5x the largest slot use seen */
            if (boost == 1.0 && !lp.slot_at_max && lp.slot_bytes > 16384 && hc.arena_used_max) {
                const double r = 1.5 * (double)hc.arena_used_max / (double)lp.slot_bytes;
                scale = std::min(64.0, std::max(1.0 / 64, scale * std::min(1.0, std::max(r, 0.25))));
            }
            break;
        }
        std::vector<uint64_t> rl(hc.retry_n);
        { int rc2 = fetch_small(ctx, rl.data(), ctx->retry.p, hc.retry_n * 8); if (rc2) return rc2; }
        std::vector<uint32_t> again, wide;
        bool ops_full = false, arena_full = false;
        for (uint64_t r : rl) {
            const uint32_t st = (uint32_t)(r >> 32), pair = (uint32_t)r;
            if (st == ST_RING) wide.push_back(pair);
            else if (st == ST_NEED8) { if (to_8bit) to_8bit->push_back(pair); }
            else { again.push_back(pair); if (st == ST_OPS) ops_full = true; else arena_full = true; }
        }
        ctx->stats.retries += (uint32_t)(again.size() + wide.size());
        if (ops_full) {
            /* grow the completion-order pool; what successful pairs wrote stays valid */
            const uint64_t old_cap = ctx->ops_pool.cap / 8;
            const uint64_t new_cap = std::max<uint64_t>(2 * old_cap, 2 * hc.ops_cursor + (1u << 20));
            DevBuf nb;
            if ((rc = ensure(ctx, nb, new_cap * 8))) return rc;
            CU(ctx, cudaMemcpy(nb.p, ctx->ops_pool.p, std::min<uint64_t>(old_cap, hc.ops_cursor) * 8, cudaMemcpyDeviceToDevice));
            cudaFree(ctx->ops_pool.p);
            ctx->ops_pool = nb;
        }
        bool wide_requeued = false;
        if (!wide.empty() && slim) {
            /* a row outgrew the ring: the wider instantiation (remembered when it was more than a
             * few pairs), and past the widest one the WARP worker */
            if (slim_li + 1 < kSlimLadderN && !getenv("WFACUDA_SLIM_P")) {
                slim_li++; wide_requeued = true;
                if (wide.size() * 50 > order.size()) ctx->slim_p_learned = std::max(ctx->slim_p_learned, kSlimLadder[slim_li]);
                again.insert(again.end(), wide.begin(), wide.end());
            } else if (to_cta) to_cta->insert(to_cta->end(), wide.begin(), wide.end());
        } else if (!wide.empty()) {
            /* too wide for the ring: retry them with a ring twice as wide while shared memory
             * allows (remembered for later batches when it was more than a few), else the CTA kernel */
            const bool can_double = !cta && lp.ring_cap < 512 && worker_smem_bytes<false>(ctx->dM, ctx->dE, lp.ring_cap * 2) * 4 <= ctx->smem_optin;
            if (can_double) {
                min_cap = lp.ring_cap * 2; wide_requeued = true;
                if (wide.size() * 50 > order.size()) ctx->ring_cap_learned = min_cap;
                again.insert(again.end(), wide.begin(), wide.end());
            } else if (to_cta) to_cta->insert(to_cta->end(), wide.begin(), wide.end());
        }
        if (arena_full) {
            if (!cta && lp.slot_at_max && lp.slot_bytes >= (15ull << 30) && to_cta) {
                /* beyond a warp slot's 32-bit reach: the CTA worker takes them */
                for (uint64_t r : rl) if ((uint32_t)(r >> 32) == ST_ARENA) to_cta->push_back((uint32_t)r);
                std::vector<uint32_t> keep;
                for (uint64_t r : rl) if ((uint32_t)(r >> 32) == ST_OPS) keep.push_back((uint32_t)r);
                if (wide_requeued) keep.insert(keep.end(), wide.begin(), wide.end());
                again.swap(keep);
            } else if (lp.slot_at_max && lp.workers <= (uint64_t)(cta ? 1 : 4)) {
                /* one worker already owns the whole budget: these pairs cannot be aligned on this device */
                std::vector<uint32_t> keep;
                for (uint64_t r : rl) {
                    const uint32_t st = (uint32_t)(r >> 32), pair = (uint32_t)r;
                    if (st != ST_ARENA) continue;
                    Result res; memset(&res, 0, sizeof res); res.status = ST_RESOURCES;
                    CU(ctx, cudaMemcpy((Result *)b->d_results + pair, &res, sizeof res, cudaMemcpyHostToDevice));
                }
                for (uint64_t r : rl) if ((uint32_t)(r >> 32) == ST_OPS) keep.push_back((uint32_t)r);
                if (wide_requeued) keep.insert(keep.end(), wide.begin(), wide.end());
                again.swap(keep);
            } else boost *= 4.0;
            scale = std::min(64.0, scale * 2.0);
        }
        requeued.swap(again);
    }
    return 0;
}


/* ---- WIDE class: one thread-block cluster per pair, live rows in (distributed) shared memory (wfa_wide.cuh) ---- */

bool wide_class_enabled(const wfacuda_ctx *ctx)
{
    const wfacuda_config &c = ctx->cfg;
    if (c.adaptive || ctx->dump_mode) return false;
    if (c.flags & (WFACUDA_FLAG_FORCE_8BIT | WFACUDA_FLAG_SEMIGLOBAL_LITERAL | WFACUDA_FLAG_NO_WIDE)) return false;
    if (getenv("WFACUDA_NO_WIDE")) return false;
    return ctx->xg == SLIM_XG && ctx->oeg == SLIM_OEG && ctx->eg == SLIM_EG;
}

/* Pairs of `order0` through the WIDE worker: forward launch (clusters) + finish launch (backtraces) per
 * sub-batch of as many pairs as the arena holds slots.  Pairs it cannot take (too wide for eight CTAs'
 * shared memory, target beyond 16-bit offsets) go to *to_cta, pairs with a non-ACGT byte to *to_8bit. */
int run_wide_class(wfacuda_ctx *ctx, wfacuda_batch *b, const std::vector<uint32_t> &order0, KParams base,
                   std::vector<uint32_t> *to_cta, std::vector<uint32_t> *to_8bit)
{
    if (order0.empty()) return 0;
    const bool semi = !ctx->cfg.global_alignment;
    const void *kfn = semi ? (const void *)wide_kernel<true> : (const void *)wide_kernel<false>;
    double boost = 1.0;
    std::vector<uint32_t> requeued;
    for (int attempt = 0; ; attempt++) {
        const std::vector<uint32_t> &order = attempt == 0 ? order0 : requeued;
        if (order.empty()) break;
        if (attempt > 24) return fail(ctx, WFACUDA_E_NOMEM, "pairs still out of resources after %d retries", attempt);
        /* geometry of the launch: widest pair decides cluster size and segment, the longest sequences the windows */
        uint64_t wmax = 1, need_max = 0; uint32_t seq_ent = 4;
        std::vector<uint32_t> fit, too_wide;
        const uint32_t head = (uint32_t)WIDE_HEAD_BYTES + 64;
        for (uint32_t pr : order) {
            const uint32_t dn = b->n_of(pr), dm = b->m_of(pr);
            const uint64_t w = (uint64_t)dn + dm - 1;
            const uint32_t ent = ((dn + 15) >> 4) + ((dm + 15) >> 4) + 2;
            /* fits eight CTAs?  (9 rows of 16-bit offsets + the windows, per CTA) */
            const uint64_t seg8 = ((w + 7) / 8 + 63) & ~63ull;
            if (dm > WIDE_MAX_M || wide_smem_bytes((uint32_t)seg8, ent) + head > ctx->smem_optin) { too_wide.push_back(pr); continue; }
            fit.push_back(pr);
            wmax = std::max(wmax, w); seq_ent = std::max(seq_ent, ent);
            need_max = std::max(need_max, estimate(ctx, dn, dm).arena);
        }
        if (to_cta) to_cta->insert(to_cta->end(), too_wide.begin(), too_wide.end());
        else if (!too_wide.empty()) return fail(ctx, WFACUDA_E_INVALID, "internal: WIDE class without a fallback");
        if (fit.empty()) break;
        int C = 1, threads = 0; uint32_t seg = 0; uint64_t smem64 = 0;
        if (wfacuda_wide_plan(wmax, seq_ent, ctx->smem_optin, &C, &seg, &threads, &smem64) != 0) return fail(ctx, WFACUDA_E_INVALID, "internal: WIDE geometry");
        if (const char *e = getenv("WFACUDA_WIDE_CLUSTER")) {         /* testing: small pairs over several CTAs */
            C = std::max(1, std::min(WIDE_MAX_CLUSTER, atoi(e)));
            seg = (uint32_t)(((wmax + C - 1) / C + 63) & ~63ull);
            threads = (int)std::min<uint32_t>(1024, std::max<uint32_t>(128, ((seg / 2 + 31) / 32) * 32));
        }
        const size_t smem = wide_smem_bytes(seg, seq_ent);
        if (const char *e = getenv("WFACUDA_WIDE_THREADS")) threads = std::max(32, std::min(1024, atoi(e) / 32 * 32));
        /* slots: 8-byte cells instead of the estimate's 12, one slot per pair of a sub-batch */
        uint64_t slot = (uint64_t)((double)need_max * (8.0 / 12.0) * ctx->arena_scale_wide / std::max(ctx->arena_scale, 1e-9) * boost);
        slot = (std::max<uint64_t>(slot, 65536) + 255) & ~255ull;
        const uint64_t budget = arena_budget(ctx, ctx->arena.cap == 0 || boost > 1.0);
        bool slot_at_max = false;
        if (slot > budget) { slot = budget & ~255ull; slot_at_max = true; }
        const uint64_t kWideSlotMax = 30ull << 30;                   /* 32-bit cell indices */
        if (slot > kWideSlotMax) { slot = kWideSlotMax; slot_at_max = true; }
        const size_t cap = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(fit.size(), budget / slot));
        int rc;
        if ((rc = ensure(ctx, ctx->arena, slot * cap))) return rc;
        if ((rc = ensure(ctx, ctx->work, fit.size() * 4))) return rc;
        if ((rc = ensure(ctx, ctx->retry, fit.size() * 8 + 16))) return rc;
        if ((rc = ensure(ctx, ctx->wide_rec, cap * sizeof(FwdOut)))) return rc;
        if ((rc = staged_h2d(ctx, ctx->work.p, fit.data(), fit.size() * 4))) return rc;
        CU(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t lc{};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.blockDim = dim3((unsigned)threads); lc.dynamicSmemBytes = smem; lc.stream = ctx->stream; lc.attrs = at; lc.numAttrs = 1;
        lc.gridDim = dim3((unsigned)C);
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &lc) != cudaSuccess || max_clusters < 1) {
            /* no such cluster can be resident on this device (partitioned GPU, other tenants' shared memory ...): the CTA worker takes the pairs */
            cudaGetLastError();
            if (!to_cta) return fail(ctx, WFACUDA_E_CUDA, "WIDE kernel: a cluster of %d CTAs x %d threads with %zu bytes of shared memory cannot be resident", C, threads, smem);
            to_cta->insert(to_cta->end(), fit.begin(), fit.end());
            break;
        }
        std::vector<uint32_t> again;
        bool ops_full = false, arena_full = false;
        uint64_t used_max = 0;
        Counters *dc = (Counters *)ctx->ctr.p;
        for (size_t i0 = 0; i0 < fit.size(); i0 += cap) {
            const size_t cnt = std::min(cap, fit.size() - i0);
            CU(ctx, dev_fill(&dc->retry_n, 0, (size_t)((char *)&dc->launch_end - (char *)&dc->retry_n), ctx->stream));
            KParams P = base;
            P.work = (const uint32_t *)ctx->work.p + i0; P.n_work = (uint32_t)cnt;
            P.arena = (uint8_t *)ctx->arena.p; P.slot_bytes = slot; P.group = 1;
            P.retry = (uint64_t *)ctx->retry.p; P.ctr = dc; P.retry_ctr = &dc->retry_n;
            P.ops_pool = (uint64_t *)ctx->ops_pool.p; P.ops_cap = ctx->ops_pool.cap / 8;
            P.wide_seg = (int32_t)seg; P.wide_seq_cap = seq_ent; P.wide_rec = (FwdOut *)ctx->wide_rec.p;
            lc.gridDim = dim3((unsigned)(std::min<size_t>(cnt, (size_t)max_clusters) * C));
            void *args[1] = {&P};
            CU(ctx, cudaLaunchKernelExC(&lc, kfn, args));
            wide_finish_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(P);
            CU(ctx, cudaGetLastError());
            ctx->stats.kernel_launches += 2; ctx->stats.align_launches += 2; ctx->hc_cache_valid = false;
            ctx->stats.arena_bytes = std::max<uint64_t>(ctx->stats.arena_bytes, slot * cap);
            Counters hc;
            { int rc2 = fetch_small(ctx, &hc, dc, sizeof hc); if (rc2) return rc2; }
            ctx->hc_cache = hc; ctx->hc_cache_valid = true;
            used_max = std::max<uint64_t>(used_max, hc.arena_used_max);
            if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   launch wide attempt %d: %zu pairs (of %zu), %d clusters of %d x %d thr, seg %u, smem %zu, slot %.1f MB (scale %.3f), used max %.1f MB, retry %llu\n",
                                                 attempt, cnt, fit.size(), (int)std::min<size_t>(cnt, (size_t)max_clusters), C, threads, seg, smem, slot / 1048576.0, ctx->arena_scale_wide, hc.arena_used_max / 1048576.0, (unsigned long long)hc.retry_n);
            if (hc.retry_n) {
                std::vector<uint64_t> rl(hc.retry_n);
                { int rc2 = fetch_small(ctx, rl.data(), ctx->retry.p, hc.retry_n * 8); if (rc2) return rc2; }
                for (uint64_t r : rl) {
                    const uint32_t st = (uint32_t)(r >> 32), pair = (uint32_t)r;
                    if (st == ST_RING) { if (to_cta) to_cta->push_back(pair); }
                    else if (st == ST_NEED8) { if (to_8bit) to_8bit->push_back(pair); }
                    else { again.push_back(pair); if (st == ST_OPS) ops_full = true; else arena_full = true; }
                }
                if (ops_full) {
                    /* grow the completion-order pool before the next sub-batch; what successful pairs wrote stays valid */
                    const uint64_t old_cap = ctx->ops_pool.cap / 8;
                    const uint64_t new_cap = std::max<uint64_t>(2 * old_cap, 2 * hc.ops_cursor + (1u << 20));
                    DevBuf nb;
                    if ((rc = ensure(ctx, nb, new_cap * 8))) return rc;
                    CU(ctx, cudaMemcpy(nb.p, ctx->ops_pool.p, std::min<uint64_t>(old_cap, hc.ops_cursor) * 8, cudaMemcpyDeviceToDevice));
                    cudaFree(ctx->ops_pool.p);
                    ctx->ops_pool = nb;
                    ops_full = false;
                }
            }
        }
        ctx->stats.retries += (uint32_t)again.size();
        if (again.empty()) {
            if (boost == 1.0 && !slot_at_max && slot > 65536 && used_max) {
                const double r = 1.5 * (double)used_max / (double)slot;
                ctx->arena_scale_wide = std::min(64.0, std::max(1.0 / 64, ctx->arena_scale_wide * std::min(1.0, std::max(r, 0.25))));
            }
            break;
        }
        if (arena_full) {
            if (slot_at_max && cap <= 1) {
                /* one pair already owns the whole budget: cannot be aligned on this device */
                for (uint32_t pair : again) {
                    Result res; memset(&res, 0, sizeof res); res.status = ST_RESOURCES;
                    CU(ctx, cudaMemcpy((Result *)b->d_results + pair, &res, sizeof res, cudaMemcpyHostToDevice));
                }
                again.clear();
            } else boost *= 4.0;
            ctx->arena_scale_wide = std::min(64.0, ctx->arena_scale_wide * 2.0);
        }
        requeued.swap(again);
    }
    return 0;
}


/* ---- LANE class: 32 short pairs per warp in lockstep (wfa_lane.cuh) ---------------------- */

constexpr int kLaneW = 64;                       /* ring columns: diagonals -32..31 */

bool slim_class_enabled(const wfacuda_ctx *ctx)
{
    const wfacuda_config &c = ctx->cfg;
    if (!c.global_alignment || ctx->dump_mode) return false;
    if (c.flags & (WFACUDA_FLAG_FORCE_CTA | WFACUDA_FLAG_FORCE_8BIT | WFACUDA_FLAG_NO_SLIM)) return false;
    if (getenv("WFACUDA_NO_SLIM")) return false;
    return ctx->xg == SLIM_XG && ctx->oeg == SLIM_OEG && ctx->eg == SLIM_EG;
}

bool lane_class_enabled(const wfacuda_ctx *ctx)
{
    const wfacuda_config &c = ctx->cfg;
    if (!c.global_alignment || c.adaptive) return false;
    if (c.flags & (WFACUDA_FLAG_FORCE_CTA | WFACUDA_FLAG_FORCE_8BIT | WFACUDA_FLAG_NO_LANE)) return false;
    if (getenv("WFACUDA_NO_LANE") || ctx->dump_mode) return false;
    /* at least two resident blocks per SM */
    return lane_smem_bytes(ctx->dM, ctx->dE, kLaneW, LANE_SEQ_WORDS) * WFA_LANE_WARPS * 2 <= ctx->smem_optin;
}

int grow_ops_pool(wfacuda_ctx *ctx, uint64_t cursor)
{
    /* grow the completion-order pool; what successful pairs wrote stays valid */
    const uint64_t old_cap = ctx->ops_pool.cap / 8;
    const uint64_t new_cap = std::max<uint64_t>(2 * old_cap, 2 * cursor + (1u << 20));
    DevBuf nb;
    int rc;
    if ((rc = ensure(ctx, nb, new_cap * 8))) return rc;
    CU(ctx, cudaMemcpy(nb.p, ctx->ops_pool.p, std::min<uint64_t>(old_cap, cursor) * 8, cudaMemcpyDeviceToDevice));
    cudaFree(ctx->ops_pool.p);
    ctx->ops_pool = nb;
    return 0;
}

/* Row geometry of a LANE launch (LaneGeom): the loop range of `next` (wfa.go:557-563) for every
 * score index, as the kernel's rings see it -- the union over both ways a pair can start
 * (M[0][0] or M[x][0], wfa.go:155-158), clamped with the longest sequence of the class; the
 * per-pair clamp to [-(n-1), m-1] stays in the kernel.  Rows are cut into stages at
 * `bounds` (score indices, ascending), each stage with its own group slots. */
void lane_geometry(const wfacuda_ctx *ctx, uint32_t maxlen, const std::vector<int> &bounds, uint32_t scratch_ops, LaneGeom *G)
{
    memset(G, 0, sizeof *G);
    const int KC = kLaneW / 2, L = (int)std::max<uint32_t>(maxlen, 1);
    /* which cells `next` can write at all: I[s][k] needs M[s-o-e][k-1] or I[s-e][k-1], D[s][k] needs
     * M[s-o-e][k+1] or D[s-e][k+1], M[s][k] needs M[s-x][k], I[s][k] or D[s][k] (wfa.go:579-698),
     * plus the init cell k = 0 at s = 0 and s = x.  Intervals; the row range is kept monotone
     * (hull with the previous rows) so that a ring row always covers the row it replaces. */
    struct Iv { int lo, hi; bool ok() const { return lo <= hi; } };
    const Iv none{1, 0};
    auto hull = [](Iv a, Iv b) { if (!a.ok()) return b; if (!b.ok()) return a; return Iv{std::min(a.lo, b.lo), std::max(a.hi, b.hi)}; };
    auto shift = [](Iv a, int d) { return a.ok() ? Iv{a.lo + d, a.hi + d} : a; };
    Iv Mv[64], Iw[64], Dv[64], run = none;
    int n_rows = 0;
    for (int si = 0; si < 64; si++) {
        auto at = [&](Iv *v, int src) { return src >= 0 ? v[src] : none; };
        Iw[si] = shift(hull(at(Mv, si - ctx->oeg), at(Iw, si - ctx->eg)), +1);
        Dv[si] = shift(hull(at(Mv, si - ctx->oeg), at(Dv, si - ctx->eg)), -1);
        Mv[si] = hull(hull(at(Mv, si - ctx->xg), Iw[si]), Dv[si]);
        if (si == 0 || si == ctx->xg) Mv[si] = hull(Mv[si], Iv{0, 0});
        for (Iv *v : {&Mv[si], &Iw[si], &Dv[si]}) if (v->ok()) { v->lo = std::max(v->lo, -(L - 1)); v->hi = std::min(v->hi, L - 1); if (!v->ok()) *v = none; }
        Iv row = Mv[si];
        if (row.ok()) { row = hull(row, run); run = row; }
        if (row.ok() && (row.lo < -KC + 1 || row.hi > KC - 2)) break;     /* beyond the byte rings: WARP worker */
        G->lo[si] = (int8_t)(row.ok() ? row.lo : 1); G->hi[si] = (int8_t)(row.ok() ? row.hi : 0);
        n_rows = si + 1;
    }
    G->n_rows = n_rows;
    int ns = 0;
    for (int bd : bounds) if (ns < LANE_MAX_STAGES - 1 && bd >= 0 && bd < n_rows - 1 && (ns == 0 || bd > G->stage_end[ns - 1])) G->stage_end[ns++] = bd;
    G->stage_end[ns++] = n_rows - 1;
    G->n_stages = ns;
    G->scratch_words = scratch_ops * 64u;                                   /* 8-byte ops, 32 lanes */
    const int ring_rows = (ctx->dM - 1) + 2 * (ctx->dE - 1);
    G->state_words = (uint32_t)(ring_rows * (kLaneW / 4) + 4);               /* 16-byte units: ring quarters, then the counters */
    int si = 0;
    for (int j = 0; j < ns; j++) {
        uint32_t run = G->scratch_words / 32u;
        for (; si <= G->stage_end[j]; si++)
            if (G->lo[si] <= G->hi[si]) { G->off[si] = (uint16_t)run; run += (uint32_t)(G->hi[si] - G->lo[si] + 1); }
        G->slot_words[j] = run * 32u;
    }
}

/* Runs the LANE class; pairs whose wavefront outgrows the byte ring go to *to_warp, pairs
 * with a non-ACGT byte to *to_8bit, pairs out of op scratch / ops pool / stage room are
 * re-queued here (single stage, larger scratch). */
int run_lane_class(wfacuda_ctx *ctx, wfacuda_batch *b, const std::vector<uint32_t> &order0, bool identity, KParams base,
                   std::vector<uint32_t> *to_warp, std::vector<uint32_t> *to_8bit)
{
    uint32_t scratch_ops = 32;
    const int threads = 32 * WFA_LANE_WARPS;
    std::vector<uint32_t> requeued;
    for (int attempt = 0; ; attempt++) {
        const std::vector<uint32_t> &order = attempt == 0 ? order0 : requeued;
        if (order.empty()) break;
        const bool ident = identity && attempt == 0;
        if (attempt > 8) return fail(ctx, WFACUDA_E_NOMEM, "pairs still out of resources after %d retries", attempt);
        /* words per sequence in shared memory: the longest of the class + 1 for the funnel shift */
        const int sw = (int)((b->lane_maxlen + 15) / 16) + 1;
        /* ring columns per stage: a stage runs with the narrowest ring its rows fit -- 48 columns
         * (diagonals -23..22) are 12 KB instead of 15 KB of shared memory per warp, 9 instead of
         * 7 blocks per SM; 56 columns (-27..26) give 8 */
        constexpr int NW = 3;
        const int ring_w[NW] = {kLaneW, 56, 48};
        size_t smem_w[NW];
        for (int i = 0; i < NW; i++) smem_w[i] = lane_smem_bytes(ctx->dM, ctx->dE, ring_w[i], sw) * WFA_LANE_WARPS;
        if (ctx->lane_occ_sw != sw) {
            ctx->lane_occ_sw = sw;
            for (int i = 0; i < NW; i++)
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->lane_occ[i], lane_kernel, threads, smem_w[i]) != cudaSuccess) { cudaGetLastError(); ctx->lane_occ[i] = 1; }
        }
        /* stages: where the previous batch's pairs finished (score-index quantiles); the first
         * batch of a ctx, retries and tiny launches run in one stage */
        std::vector<int> bounds;
        const uint64_t groups = (order.size() + 31) / 32;
        const bool staged = attempt == 0 && ctx->lane_hist_n >= 256 && groups >= 64 && !getenv("WFACUDA_NO_STAGES");
        if (staged) bounds.assign(ctx->lane_bounds, ctx->lane_bounds + ctx->lane_n_bounds);
        KParams P = base;
        lane_geometry(ctx, b->lane_maxlen, bounds, scratch_ops, &P.lg);
        const LaneGeom &G = P.lg;
        /* groups per stage: all of them enter stage 0; later stages are sized from the histogram
         * (a pair that finds its next stage full is re-queued) */
        double frac[LANE_MAX_STAGES] = {1.0, 1.0, 1.0, 1.0};
        for (int j = 1; j < G.n_stages; j++) {
            uint64_t later = 0;
            for (int si = G.stage_end[j - 1] + 1; si < 64; si++) later += ctx->lane_hist[si];
            frac[j] = std::min(1.0, 1.25 * (double)later / (double)std::max<uint64_t>(ctx->lane_hist_n, 1) + 0.02);
        }
        /* device memory per group of the round: its slots, the pairs' records, saved states and queue entries */
        double per_group = 32.0 * (LANE_REC_WORDS * 4 + (G.n_stages > 1 ? G.state_words * 4 + 4 * (G.n_stages - 1) : 0));
        for (int j = 0; j < G.n_stages; j++) per_group += frac[j] * G.slot_words[j] * 4.0;
        const uint64_t budget = arena_budget(ctx, ctx->arena.cap == 0);
        const uint64_t round_groups = std::min<uint64_t>(groups, std::max<uint64_t>(1, (uint64_t)((double)budget / per_group)));
        if ((double)budget < per_group) {
            /* not even one group fits: the WARP class places these pairs */
            for (uint32_t pi : order) to_warp->push_back(pi);
            break;
        }
        uint64_t cap[LANE_MAX_STAGES], arena_off[LANE_MAX_STAGES], total = 0;
        for (int j = 0; j < G.n_stages; j++) {
            cap[j] = j == 0 ? round_groups : std::min<uint64_t>(round_groups, (uint64_t)(frac[j] * (double)round_groups) + 8);
            arena_off[j] = total; total += (cap[j] * G.slot_words[j] * 4 + 255) & ~255ull;
        }
        const uint64_t round_pairs = round_groups * 32;
        const uint64_t rec_off = total; total += (round_pairs * LANE_REC_WORDS * 4 + 255) & ~255ull;
        uint64_t state_off = total, list_off[LANE_MAX_STAGES] = {0, 0, 0, 0};
        if (G.n_stages > 1) {
            total += (round_pairs * G.state_words * 4 + 255) & ~255ull;
            for (int j = 1; j < G.n_stages; j++) { list_off[j] = total; total += (cap[j] * 32 * 4 + 255) & ~255ull; }
        }
        /* Hand-over on the device: the pairs the byte rings cannot hold (ST_RING in the retry list
         * the finish kernel writes) go to a WARP launch queued right behind the finish kernel, which
         * reads that list and its length from device memory -- no host round trip in between.  Its
         * slots are generous (these are the high-score pairs); what still fails there (wider ring,
         * more arena, ops pool) lands in a second list and takes the host path below. */
        const uint32_t Lmax = b->lane_maxlen;
        const uint64_t ho_slot = (std::max<uint64_t>(estimate(ctx, Lmax, Lmax).arena * 4, 65536) + 255) & ~255ull;
        const int ho_seq_cap = (int)((((Lmax + 15) >> 4) * 2 + 2 + 3) & ~3u);
        const size_t ho_smem = worker_smem_bytes<false>(ctx->dM, ctx->dE, 64, ho_seq_cap) * 4;
        const int ho_blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx->sm_count, (order.size() + 511) / 512));
        const bool handover = !getenv("WFACUDA_HOST_HANDOVER") && ho_smem <= ctx->smem_optin && ho_slot * (uint64_t)ho_blocks * 4 <= budget;
        const uint64_t retry2_off = order.size() + 2;                          /* in 8-byte entries */
        int rc;
        if ((rc = ensure(ctx, ctx->arena, std::max<uint64_t>(total, handover ? ho_slot * (uint64_t)ho_blocks * 4 : 0)))) return rc;
        if (!ident && (rc = ensure(ctx, ctx->work, order.size() * 4))) return rc;
        if ((rc = ensure(ctx, ctx->retry, (retry2_off + order.size() + 2) * 8))) return rc;
        if (!ident) { int rc2 = staged_h2d(ctx, ctx->work.p, order.data(), order.size() * 4); if (rc2) return rc2; }
        Counters *dc = (Counters *)ctx->ctr.p;
        /* one fill for everything a launch of the class counts (retry lists, queues, histogram) */
        CU(ctx, dev_fill(&dc->retry_n, 0, (size_t)((char *)&dc->launch_end - (char *)&dc->retry_n), ctx->stream));
        if (getenv("WFACUDA_DEBUG")) CU(ctx, dev_fill(&dc->t_first, 0xff, 8, ctx->stream));
        uint8_t *abase = (uint8_t *)ctx->arena.p;
        for (int j = 0; j < G.n_stages; j++) {
            P.la.arena[j] = abase + arena_off[j]; P.la.cap[j] = (uint32_t)cap[j];
            P.la.list[j] = j ? (uint32_t *)(abase + list_off[j]) : nullptr;
        }
        P.la.rec = (uint32_t *)(abase + rec_off); P.la.state = (uint32_t *)(abase + state_off);
        P.arena = abase; P.slot_bytes = (uint64_t)G.slot_words[0] * 4; P.group = sw;      /* LANE kernel: group = words per sequence */
        P.retry = (uint64_t *)ctx->retry.p; P.ctr = dc; P.retry_ctr = &dc->retry_n;
        P.ring_cap = kLaneW; P.ops_pool = (uint64_t *)ctx->ops_pool.p; P.ops_cap = ctx->ops_pool.cap / 8;
        int stage_w[LANE_MAX_STAGES];
        {
            int si = 0;
            for (int j = 0; j < G.n_stages; j++) {
                int klo = 0, khi = 0;
                for (; si <= G.stage_end[j]; si++) if (G.lo[si] <= G.hi[si]) { klo = std::min(klo, (int)G.lo[si]); khi = std::max(khi, (int)G.hi[si]); }
                stage_w[j] = 0;
                if (!getenv("WFACUDA_LANE_W64"))
                    for (int i = NW - 1; i > 0; i--) if (klo >= -(ring_w[i] / 2 - 1) && khi <= ring_w[i] / 2 - 2) { stage_w[j] = i; break; }
            }
        }
        const double tl0 = now_ms();
        int blocks = 0;
        for (uint64_t g0 = 0; g0 < groups; g0 += round_groups) {
            const uint64_t g1 = std::min(groups, g0 + round_groups);
            const uint64_t p0 = g0 * 32, p1 = std::min<uint64_t>(order.size(), g1 * 32);
            if (g0) CU(ctx, dev_fill(dc->lane_count, 0, sizeof dc->lane_count, ctx->stream));      /* later rounds: the stage queues start over */
            P.work = ident ? nullptr : (const uint32_t *)ctx->work.p + p0; P.pair_base = (uint32_t)p0; P.n_work = (uint32_t)(p1 - p0);
            for (int j = 0; j < G.n_stages; j++) {
                const uint64_t gj = j == 0 ? g1 - g0 : std::min<uint64_t>(cap[j], g1 - g0);
                const uint64_t workers = (uint64_t)ctx->sm_count * std::max(1, ctx->lane_occ[stage_w[j]]) * WFA_LANE_WARPS;
                const uint64_t w = std::min<uint64_t>(workers, ((gj + WFA_LANE_WARPS - 1) / WFA_LANE_WARPS) * WFA_LANE_WARPS);
                blocks = (int)(w / WFA_LANE_WARPS);
                P.ring_cap = ring_w[stage_w[j]];
                lane_kernel<<<blocks, threads, smem_w[stage_w[j]], ctx->stream>>>(P, j);
                CU(ctx, cudaGetLastError());
            }
            const int fblocks = (int)std::min<uint64_t>((g1 - g0 + LANE_FINISH_WARPS - 1) / LANE_FINISH_WARPS, (uint64_t)ctx->sm_count * 16);
            lane_finish_kernel<<<fblocks, 32 * LANE_FINISH_WARPS, 0, ctx->stream>>>(P);
            CU(ctx, cudaGetLastError());
            ctx->stats.kernel_launches += G.n_stages + 1; ctx->stats.align_launches++; ctx->hc_cache_valid = false;
        }
        if (handover) {
            KParams H = base;
            H.handover = (const uint64_t *)ctx->retry.p; H.retry = (uint64_t *)ctx->retry.p + retry2_off; H.retry_ctr = &dc->retry2_n;
            H.ctr = dc; H.work = nullptr; H.n_work = 0;
            H.arena = (uint8_t *)ctx->arena.p; H.slot_bytes = ho_slot; H.group = 1; H.ring_cap = 64; H.seq_cap = ho_seq_cap;
            H.ops_pool = (uint64_t *)ctx->ops_pool.p; H.ops_cap = ctx->ops_pool.cap / 8;
            align_kernel<2, false><<<ho_blocks, 128, ho_smem, ctx->stream>>>(H);
            CU(ctx, cudaGetLastError());
            ctx->stats.kernel_launches++; ctx->stats.align_launches++; ctx->hc_cache_valid = false;
        }
        const double tl1 = now_ms();
        ctx->stats.arena_bytes = std::max<uint64_t>(ctx->stats.arena_bytes, total);
        Counters hc;
        { int rc2 = fetch_small(ctx, &hc, dc, sizeof hc); if (rc2) return rc2; }
        ctx->hc_cache = hc; ctx->hc_cache_valid = true;
        if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   lane kernels: host launch at %.2f (took %.2f), sync returned %.2f ms since call; device span %.3f ms\n", tl0 - g_dbg_t0, tl1 - tl0, now_ms() - g_dbg_t0, (hc.t_last - hc.t_first) / 1e6);
        if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   launch lane attempt %d: %zu pairs, %d stages (ends %d %d %d %d; entered %llu %llu %llu), rows %d, slot0 %.1f KB, %.1f MB device, retry %llu\n", attempt, order.size(), G.n_stages, G.stage_end[0], G.stage_end[1], G.stage_end[2], G.stage_end[3], (unsigned long long)hc.lane_count[1], (unsigned long long)hc.lane_count[2], (unsigned long long)hc.lane_count[3], G.n_rows, G.slot_words[0] / 256.0, total / 1e6, (unsigned long long)hc.retry_n);
        /* learn the next batch's stage boundaries: the score indices by which 40 % and 85 % of
         * the sampled pairs had finished (first attempt of a batch only) */
        if (attempt == 0) {
            uint64_t tot = 0;
            for (int i = 0; i < 64; i++) tot += hc.lane_hist[i];
            if (tot >= 256) {
                for (int i = 0; i < 64; i++) ctx->lane_hist[i] = hc.lane_hist[i];
                ctx->lane_hist_n = tot;
                double qs[3] = {0.40, 0.85, 1.0}; int nq = 2;           /* measured best on config 2 (profiles/r1_lane_stages.md) */
                if (const char *e = getenv("WFACUDA_LANE_QUANTILES")) {          /* e.g. "0.3,0.8": experiments */
                    nq = 0;
                    for (const char *p = e; *p && nq < 3; ) { char *end; const double v = strtod(p, &end); if (end == p) break; qs[nq++] = v; p = *end ? end + 1 : end; }
                }
                ctx->lane_n_bounds = 0;
                uint64_t cum = 0; int qi = 0;
                for (int i = 0; i < 64 && qi < nq; i++) {
                    cum += hc.lane_hist[i];
                    while (qi < nq && (double)cum >= qs[qi] * (double)tot) {
                        if (ctx->lane_n_bounds == 0 || ctx->lane_bounds[ctx->lane_n_bounds - 1] < i) ctx->lane_bounds[ctx->lane_n_bounds++] = i;
                        qi++;
                    }
                }
            }
        }
        if (handover && hc.retry2_n == 0 && hc.handover_other == 0) {
            /* every entry of the list was a ring overflow and the WARP launch aligned them all */
            ctx->lane_handed += hc.retry_n;
            break;
        }
        std::vector<uint64_t> rl(hc.retry_n), rl2(handover ? hc.retry2_n : 0);
        if (hc.retry_n) { int rc2 = fetch_small(ctx, rl.data(), ctx->retry.p, hc.retry_n * 8); if (rc2) return rc2; }
        if (!rl2.empty()) { int rc2 = fetch_small(ctx, rl2.data(), (const uint64_t *)ctx->retry.p + retry2_off, rl2.size() * 8); if (rc2) return rc2; }
        std::vector<uint32_t> again;
        bool ops_full = false;
        uint64_t rings = 0;
        for (uint64_t r : rl) {
            const uint32_t st = (uint32_t)(r >> 32), pair = (uint32_t)r;
            if (st == ST_RING) { if (handover) rings++; else to_warp->push_back(pair); }
            else if (st == ST_NEED8) to_8bit->push_back(pair);
            else { again.push_back(pair); if (st == ST_OPS) ops_full = true; }
        }
        for (uint64_t r : rl2) {                                               /* what the hand-over launch could not place */
            const uint32_t st = (uint32_t)(r >> 32), pair = (uint32_t)r;
            if (st == ST_NEED8) to_8bit->push_back(pair); else to_warp->push_back(pair);
            if (st == ST_OPS) ops_full = true;
        }
        if (handover) ctx->lane_handed += rings - std::min<uint64_t>(rings, rl2.size());
        if (hc.retry_n == 0) break;
        ctx->stats.retries += (uint32_t)again.size();
        if (ops_full && (rc = grow_ops_pool(ctx, hc.ops_cursor))) return rc;
        if (scratch_ops >= 512) {
            /* more ops than a LANE pair can have: let the WARP class place what is left */
            for (uint32_t pi : again) to_warp->push_back(pi);
            again.clear();
        }
        scratch_ops *= 4;
        requeued.swap(again);
    }
    return 0;
}

} // namespace

/* ========================================================================== API */

extern "C" {

const char *wfacuda_last_error(const wfacuda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_tls_error.c_str(); }

int wfacuda_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return 0; fail(nullptr, 0, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); return WFACUDA_E_CUDA; }
    int ok = 0;
    for (int i = 0; i < n; i++) { cudaDeviceProp p; if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++; }
    return ok;
}

wfacuda_ctx *wfacuda_create(int device, const wfacuda_config *cfg)
{
    wfacuda_ctx *ctx = new wfacuda_ctx();
    auto bail = [&](void) -> wfacuda_ctx * { g_tls_error = ctx->err; wfacuda_destroy(ctx); return nullptr; };
    if (apply_config(ctx, cfg) != 0) return bail();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        fail(ctx, WFACUDA_E_CUDA, "no CUDA device available (%s); libwfacuda has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
        return bail();
    }
    if (device < 0 || device >= n) { fail(ctx, WFACUDA_E_INVALID, "device %d out of range (have %d)", device, n); return bail(); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "cudaGetDeviceProperties failed"); return bail(); }
    if (prop.major != 10) { fail(ctx, WFACUDA_E_CUDA, "device %d is sm_%d%d; libwfacuda is built for sm_100a only", device, prop.major, prop.minor); return bail(); }
    ctx->device = device; ctx->sm_count = prop.multiProcessorCount; ctx->smem_optin = prop.sharedMemPerBlockOptin; ctx->total_mem = prop.totalGlobalMem;
    if (cudaSetDevice(device) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "cudaSetDevice failed"); return bail(); }
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&ctx->stream_main, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
            cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "stream creation failed"); return bail(); }
        ctx->stream = ctx->stream_main;
    }
    for (auto &ev : ctx->ev) if (cudaEventCreate(&ev) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "event creation failed"); return bail(); }
    ctx->pinned_cap = 16u << 20;
    for (int i = 0; i < 2; i++) {
        if (cudaHostAlloc(&ctx->pinned[i], ctx->pinned_cap, cudaHostAllocDefault) != cudaSuccess) { fail(ctx, WFACUDA_E_NOMEM, "pinned staging allocation failed"); return bail(); }
        if (cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "event creation failed"); return bail(); }
    }
    ctx->smem_optin -= 1024;       /* room for the kernels' static shared memory */
    if (cudaFuncSetAttribute(align_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess ||
        cudaFuncSetAttribute(align_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess ||
        cudaFuncSetAttribute(align_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess ||
        cudaFuncSetAttribute(align_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess ||
        cudaFuncSetAttribute(lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin) != cudaSuccess) {
        fail(ctx, WFACUDA_E_CUDA, "cudaFuncSetAttribute(max dynamic shared memory) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return bail();
    }
    /* One shared-memory carveout for every kernel: an SM has to drain before it can switch
     * carveout, so a small follow-up kernel with the default (small) preference would wait
     * behind the whole backlog of other pipeline workers' big kernels (measured: 6 ms). */
    {
        const int mx = (int)cudaSharedmemCarveoutMaxShared;
        cudaFuncSetAttribute(align_kernel<2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(align_kernel<2, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(align_kernel<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(align_kernel<8, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(lane_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(pack_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaFuncSetAttribute(pack_short_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        for (int li = 0; li < kSlimLadderN; li++) for (int z = 0; z < 3; z++) for (int a = 0; a < 2; a++)
            cudaFuncSetAttribute(slim_kernel_ptr(kSlimLadder[li], z, a != 0), cudaFuncAttributePreferredSharedMemoryCarveout, mx);
        cudaGetLastError();
    }
    return ctx;
}

void wfacuda_destroy(wfacuda_ctx *ctx)
{
    if (!ctx) return;
    for (wfacuda_ctx *c : ctx->subs) wfacuda_destroy(c);
    ctx->subs.clear();
    cudaSetDevice(ctx->device);
    if (ctx->stream_main) cudaStreamSynchronize(ctx->stream_main);
    if (ctx->stream_hi) cudaStreamSynchronize(ctx->stream_hi);
    for (DevBuf *b : {&ctx->arena, &ctx->retry, &ctx->work, &ctx->ctr, &ctx->ops_pool, &ctx->wire_dev, &ctx->render_meta, &ctx->render_cigar, &ctx->render_text}) if (b->p) cudaFree(b->p);
    for (auto &f : ctx->free_dev) cudaFree(f.first);
    for (int i = 0; i < 2; i++) { if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]); if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]); }
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
    if (ctx->pin_descs) cudaFreeHost(ctx->pin_descs);
    if (ctx->pin_pool) cudaFreeHost(ctx->pin_pool);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    if (ctx->h2d_fifo) { cudaStreamSynchronize(ctx->h2d_fifo); cudaStreamDestroy(ctx->h2d_fifo); }
    if (ctx->stream_main) cudaStreamDestroy(ctx->stream_main);
    if (ctx->stream_hi) cudaStreamDestroy(ctx->stream_hi);
    delete ctx;
}

void *wfacuda_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, std::max<size_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
        fail(nullptr, WFACUDA_E_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

void wfacuda_host_free(void *p) { if (p) cudaFreeHost(p); }

int wfacuda_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return fail(nullptr, WFACUDA_E_INVALID, "nothing to register");
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess)
        return fail(nullptr, WFACUDA_E_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

int wfacuda_host_unregister(void *p)
{
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return WFACUDA_E_INVALID; }
    return 0;
}

int wfacuda_set_config(wfacuda_ctx *ctx, const wfacuda_config *cfg)
{
    if (!ctx) return fail(nullptr, WFACUDA_E_INVALID, "ctx is NULL");
    wfacuda_config old = ctx->cfg;
    int rc = apply_config(ctx, cfg);
    if (rc == 0 && memcmp(&old, cfg, sizeof old) != 0) {
        ctx->arena_scale = 1.0; ctx->arena_scale_slim = 1.0; ctx->arena_scale_wide = 1.0; ctx->slim_p_learned = 0; ctx->lane_hist_n = 0; ctx->lane_n_bounds = 0; ctx->lane_occ[0] = ctx->lane_occ[1] = ctx->lane_occ[2] = 0; ctx->lane_occ_sw = 0; ctx->ring_cap_learned = 0; memset(ctx->occ_cache, 0, sizeof ctx->occ_cache);
        for (wfacuda_ctx *c : ctx->subs) wfacuda_destroy(c);      /* re-created with the new config on demand */
        ctx->subs.clear();
    }
    return rc;
}

int wfacuda_get_stats(const wfacuda_ctx *ctx, wfacuda_stats *out)
{
    if (!ctx || !out) return WFACUDA_E_INVALID;
    *out = ctx->stats;
    return 0;
}

uint64_t wfacuda_last_ops_total(const wfacuda_ctx *ctx) { return ctx ? ctx->last_ops_total : 0; }
uint64_t wfacuda_batch_ops_total(const wfacuda_batch *b) { return b ? b->ops_total : 0; }

void wfacuda_batch_free(wfacuda_ctx *ctx, wfacuda_batch *b)
{
    if (!b) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        dev_give(ctx, &b->d_raw, &b->sz_raw); dev_give(ctx, &b->d_packed, &b->sz_packed);
        dev_give(ctx, &b->d_descs, &b->sz_descs); dev_give(ctx, &b->d_flags, &b->sz_flags);
        dev_give(ctx, &b->d_results, &b->sz_results); dev_give(ctx, &b->d_where, &b->sz_where);
        if (ctx->pool_owner == b) ctx->pool_owner = nullptr;
        if (ctx->pin_descs_owner == b) ctx->pin_descs_owner = nullptr;
    }
    delete b;
}

wfacuda_batch *wfacuda_batch_upload(wfacuda_ctx *ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                                    const uint64_t *q_off, const uint32_t *q_len,
                                    const uint64_t *t_off, const uint32_t *t_len)
{
    if (!ctx) { fail(nullptr, WFACUDA_E_INVALID, "ctx is NULL"); return nullptr; }
    if (n_pairs > 0xfffffff0ull) { fail(ctx, WFACUDA_E_INVALID, "at most 2^32-16 pairs per batch"); return nullptr; }
    if (n_pairs && (!seq_bytes || !q_off || !q_len || !t_off || !t_len)) { fail(ctx, WFACUDA_E_INVALID, "NULL input array"); return nullptr; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { fail(ctx, WFACUDA_E_CUDA, "cudaSetDevice failed"); return nullptr; }
    wfacuda_batch *b = new wfacuda_batch();
    b->n_pairs = n_pairs;
    b->host_status.assign(n_pairs, ST_PENDING);
    const bool dbg = getenv("WFACUDA_DEBUG") != nullptr;
    /* pipeline worker with page-locked caller memory: sequence bytes and descriptors go through
     * the parent's FIFO upload stream */
    const bool fifo_possible = ctx->h2d_parent && ctx->h2d_parent->h2d_fifo && n_pairs && ctx->pin_descs_owner == nullptr && !getenv("WFACUDA_NO_FIFO");
    bool fifo = false;        /* decided in body(): the bytes to upload must sit in page-locked memory (the caller's, or the gathered pool) */
    bool fifo_turn_held = false;
    struct TurnGuard { wfacuda_ctx *c; bool &held; ~TurnGuard() { if (held && c->h2d_turn) c->h2d_turn->release(); } } turn_guard{ctx, fifo_turn_held};
    auto body = [&]() -> int {
        const double t0 = now_ms();
        ctx->stats = wfacuda_stats{};
        /* validation (wfa.go:202-209) + extent of the byte pool actually referenced */
        uint64_t lo = UINT64_MAX, hi = 0, words = 0, sum_len = 0;
        for (uint64_t i = 0; i < n_pairs; i++) {
            const uint32_t dn = q_len[i], dm = t_len[i];
            if (dn == 0 || dm == 0) { b->host_status[i] = ST_EMPTY; b->n_invalid++; continue; }
            if (dn > WFACUDA_MAX_SEQ_LEN || dm > WFACUDA_MAX_SEQ_LEN) { b->host_status[i] = ST_TOO_LONG; b->n_invalid++; continue; }
            lo = std::min(lo, std::min(q_off[i], t_off[i]));
            hi = std::max(hi, std::max(q_off[i] + dn, t_off[i] + dm));
            words += (((uint64_t)((dn + 15) >> 4) + 3) & ~3ull) + (((uint64_t)((dm + 15) >> 4) + 3) & ~3ull);
            sum_len += (uint64_t)dn + dm;
        }
        if (lo == UINT64_MAX) { lo = 0; hi = 0; }
        uint64_t base = lo & ~(uint64_t)15;                  /* keep the caller's alignment mod 16 */
        /* The batch's sequences are copied as ONE range [base, hi) of the caller's pool -- right for
         * the usual pool of interleaved pairs.  When they are scattered over a much larger pool (all
         * queries then all targets cut into chunks, windows into one shared reference, the runs of a
         * multi-GPU shard) that range would be most of the pool for every chunk: gather them into a
         * page-locked pool of the ctx instead and address that. */
        const bool gather = hi - base > sum_len + sum_len / 2 + (1u << 18);          /* (an interleaved pool with padded pairs stays below 1.5x) */
        const uint8_t *src = seq_bytes;
        if (gather) {
            if (ctx->pin_pool_cap < sum_len + 64) {
                if (ctx->pin_pool) cudaFreeHost(ctx->pin_pool);
                ctx->pin_pool = nullptr; ctx->pin_pool_cap = 0;
                const size_t want = sum_len + sum_len / 8 + 4096;
                if (cudaHostAlloc((void **)&ctx->pin_pool, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(ctx, WFACUDA_E_NOMEM, "page-locked sequence pool allocation failed"); }
                ctx->pin_pool_cap = want;
            }
            src = ctx->pin_pool; base = 0; hi = sum_len;
        }
        /* a pipeline worker stages a dense range of pageable caller memory the same way: one copy into
         * its page-locked pool, which then goes through the FIFO like page-locked caller memory does */
        const bool caller_pinned = is_pinned(seq_bytes);
        if (!gather && fifo_possible && !caller_pinned && hi > base) {
            const size_t want = hi - base + 64;
            if (ctx->pin_pool_cap < want) {
                if (ctx->pin_pool) cudaFreeHost(ctx->pin_pool);
                ctx->pin_pool = nullptr; ctx->pin_pool_cap = 0;
                if (cudaHostAlloc((void **)&ctx->pin_pool, want + want / 8, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return fail(ctx, WFACUDA_E_NOMEM, "page-locked sequence pool allocation failed"); }
                ctx->pin_pool_cap = want + want / 8;
            }
            memcpy(ctx->pin_pool, seq_bytes + base, hi - base);
            src = ctx->pin_pool - base;                      /* src + base is the staged copy */
        }
        fifo = fifo_possible && (gather || src != seq_bytes || caller_pinned);
        if (fifo) {
            if (ctx->pin_descs_cap < n_pairs) {
                if (ctx->pin_descs) cudaFreeHost(ctx->pin_descs);
                ctx->pin_descs = nullptr; ctx->pin_descs_cap = 0;
                if (cudaHostAlloc((void **)&ctx->pin_descs, (n_pairs + n_pairs / 8) * sizeof(WireDesc), cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError(); return fail(ctx, WFACUDA_E_NOMEM, "page-locked descriptor buffer allocation failed");
                }
                ctx->pin_descs_cap = n_pairs + n_pairs / 8;
            }
            ctx->pin_descs_owner = b;
        }
        uint64_t gcur = 0;
        /* pipeline worker: 20-byte wire descriptors when the chunk's offsets fit 32 bits */
        const bool wire = fifo && hi - base < 0xfffffff0ull && words < 0xfffffff0ull;
        if (!wire) { b->descs_own.resize(n_pairs); b->descs = b->descs_own.data(); }
        else b->wire = ctx->pin_descs;
        words = 0;
        const bool any_invalid = b->n_invalid != 0;
        for (uint64_t i = 0; i < n_pairs; i++) {
            const bool ok = !any_invalid || b->host_status[i] == ST_PENDING;
            const uint32_t dn = ok ? q_len[i] : 0, dm = ok ? t_len[i] : 0;
            uint64_t qb = ok ? q_off[i] - base : 0, tb = ok ? t_off[i] - base : 0;
            const uint64_t qw = ok ? words : 0;
            if (gather && ok) {
                qb = gcur; memcpy(ctx->pin_pool + gcur, seq_bytes + q_off[i], dn); gcur += dn;
                tb = gcur; memcpy(ctx->pin_pool + gcur, seq_bytes + t_off[i], dm); gcur += dm;
            }
            if (ok) {
                words += ((uint64_t)((dn + 15) >> 4) + 3) & ~3ull;
                b->seq_bases += (uint64_t)dn + dm;
                b->max_nm = std::max<uint64_t>(b->max_nm, (uint64_t)dn + dm);
                b->max_len = std::max(b->max_len, std::max(dn, dm));
            }
            const uint64_t tw = ok ? words : 0;
            if (ok) words += ((uint64_t)((dm + 15) >> 4) + 3) & ~3ull;
            if (wire) { WireDesc &w = ctx->pin_descs[i]; w.q_byte = (uint32_t)qb; w.t_byte = (uint32_t)tb; w.q_word = (uint32_t)qw; w.n = dn; w.m = dm; }
            else { PairDesc &d = b->descs[i]; d.q_byte = qb; d.t_byte = tb; d.q_word = qw; d.t_word = tw; d.n = dn; d.m = dm; }
        }
        b->raw_bytes = hi - base; b->packed_words = words;
        const double t1 = now_ms();
        int rc;
        if ((rc = dev_take(ctx, &b->d_raw, &b->sz_raw, b->raw_bytes + 64))) return rc;
        if ((rc = dev_take(ctx, &b->d_packed, &b->sz_packed, (words + 16) * 4))) return rc;
        if ((rc = dev_take(ctx, &b->d_descs, &b->sz_descs, n_pairs * sizeof(PairDesc)))) return rc;
        if ((rc = dev_take(ctx, &b->d_flags, &b->sz_flags, n_pairs + 16))) return rc;
        if ((rc = dev_take(ctx, &b->d_results, &b->sz_results, n_pairs * sizeof(Result)))) return rc;
        if ((rc = dev_take(ctx, &b->d_where, &b->sz_where, n_pairs * 8))) return rc;
        if (fifo) {
            /* descriptors first, then the sequence bytes, queued together: the copy engine serves
             * copies in the order they were issued, whatever stream they are on */
            wfacuda_ctx *par = ctx->h2d_parent;
            if (wire && (rc = ensure(ctx, ctx->wire_dev, n_pairs * sizeof(WireDesc)))) return rc;
            /* bounded queue depth, chunk order: released when this chunk's copies are done */
            if (ctx->h2d_ticket >= 0) ctx->h2d_turn->acquire_ordered((uint64_t)ctx->h2d_ticket); else ctx->h2d_turn->acquire();
            fifo_turn_held = true;                           /* released when this chunk's copies are done, or by the guard on any early return */
            std::lock_guard<std::mutex> lk(par->h2d_mu);
            if (wire) CU(ctx, cudaMemcpyAsync(ctx->wire_dev.p, ctx->pin_descs, n_pairs * sizeof(WireDesc), cudaMemcpyHostToDevice, par->h2d_fifo));
            else CU(ctx, cudaMemcpyAsync(b->d_descs, b->descs, n_pairs * sizeof(PairDesc), cudaMemcpyHostToDevice, par->h2d_fifo));
            if (b->raw_bytes) CU(ctx, cudaMemcpyAsync(b->d_raw, src + base, b->raw_bytes, cudaMemcpyHostToDevice, par->h2d_fifo));
            CU(ctx, cudaEventRecord(ctx->ev_h2d, par->h2d_fifo));
            ctx->stats.h2d_bytes += b->raw_bytes + n_pairs * (wire ? sizeof(WireDesc) : sizeof(PairDesc));
        } else if (b->raw_bytes) {
            if (ctx->h2d_turn && b->raw_bytes >= 65536 && is_pinned(src + base)) {
                struct Turn { wfacuda_ctx::Turns *t; Turn(wfacuda_ctx::Turns *t_) : t(t_) { t->acquire(); } ~Turn() { t->release(); } } turn(ctx->h2d_turn);
                CU(ctx, cudaMemcpyAsync(b->d_raw, src + base, b->raw_bytes, cudaMemcpyHostToDevice, ctx->stream));
                CU(ctx, wait_stream(ctx, ctx->stream));
                ctx->stats.h2d_bytes += b->raw_bytes;
            } else if ((rc = staged_h2d(ctx, b->d_raw, src + base, b->raw_bytes))) return rc;
        }
        if (b->raw_bytes) CU(ctx, dev_fill((char *)b->d_raw + b->raw_bytes, 0, 64, ctx->stream));
        const double t2 = now_ms();
        if (n_pairs && !fifo) if ((rc = staged_h2d(ctx, b->d_descs, b->descs, n_pairs * sizeof(PairDesc)))) return rc;
        const double t3 = now_ms();
        /* cost bins: longest first (counting sort on half-octave buckets of n+m).  A pair goes to
         * the CTA class when its wavefront cannot fit the widest shared-memory ring: width is
         * n+m-1 for semi-global, bounded by the score guess otherwise (same guess as estimate()) */
        const bool force_cta = ctx->cfg.flags & WFACUDA_FLAG_FORCE_CTA;
        const int warp_cap_max = 512;
        const wfacuda_config &c = ctx->cfg;
        const double oe = (double)c.gap_open + c.gap_ext, per_edit = (c.mismatch + 2.0 * oe) / 3.0;
        const bool narrow_always = c.global_alignment && c.adaptive && 1.5 * c.max_dist_diff + c.min_wf_len + 16.0 <= warp_cap_max;
        const bool lane_ok = lane_class_enabled(ctx);
        std::vector<uint8_t> key(n_pairs);                       /* bucket | class << 6; 255 = not aligned */
        uint32_t counts[3][66] = {{0}};
        int first_key = -1; bool uniform = true;
        for (uint64_t i = 0; i < n_pairs; i++) {
            if (b->host_status[i] != ST_PENDING) { key[i] = 255; uniform = false; continue; }
            const struct { uint32_t n, m; } d = {b->n_of(i), b->m_of(i)};
            const uint64_t nm = (uint64_t)d.n + d.m;
            const int lg = 63 - __builtin_clzll(nm | 1);
            const int bk = 63 - (2 * lg + (int)((nm >> (lg > 0 ? lg - 1 : 0)) & 1));
            int cls = force_cta ? 1 : 0;
            if (lane_ok && d.n <= (uint32_t)LANE_MAX_LEN && d.m <= (uint32_t)LANE_MAX_LEN) { cls = 2; b->lane_maxlen = std::max(b->lane_maxlen, std::max(d.n, d.m)); }
            else if (!cls && !narrow_always && nm - 1 > (uint64_t)warp_cap_max) {
                if (!c.global_alignment) cls = 1;
                else {
                    const double L = std::min(d.n, d.m);
                    const double score = 0.10 * L * per_edit + oe + std::abs((double)d.m - d.n) * c.gap_ext + 4.0 * c.mismatch;
                    double w = std::min((double)nm - 1, 2.0 * score / c.gap_ext + 3);
                    if (c.adaptive) w = std::min(w, 1.5 * c.max_dist_diff + c.min_wf_len + 16.0);
                    cls = w > warp_cap_max;
                }
            }
            key[i] = (uint8_t)(bk | cls << 6);
            if (first_key < 0) first_key = key[i]; else if (key[i] != first_key) uniform = false;
            counts[cls][bk + 1]++;
        }
        std::vector<uint32_t> *ords[3] = {&b->order_warp, &b->order_cta, &b->order_lane};
        if (uniform && n_pairs) {
            /* one bucket, one class (the usual batch of equal-length reads): identity order */
            std::vector<uint32_t> &ord = *ords[first_key >> 6];
            ord.resize(n_pairs);
            std::iota(ord.begin(), ord.end(), 0u);
            b->identity_cls = first_key >> 6;
        } else {
            for (int k2 = 0; k2 < 3; k2++) for (int k = 1; k < 66; k++) counts[k2][k] += counts[k2][k - 1];
            for (int k2 = 0; k2 < 3; k2++) ords[k2]->resize(counts[k2][65]);
            for (uint64_t i = 0; i < n_pairs; i++) {
                if (key[i] == 255) continue;
                const int cls = key[i] >> 6, bk = key[i] & 63;
                (*ords[cls])[counts[cls][bk]++] = (uint32_t)i;
            }
        }
        if (slim_class_enabled(ctx) && slim_size_class(b->max_len) == 1 && b->order_warp.size() >= 100000) {      /* (a second launch per pipeline chunk would cost more than it saves) */
            /* reads of about 1 kbp: the pairs whose target fits the SLIM worker's 10-bit fields will take its 32-bit
             * cell word (half the arena bytes, cheaper packing), only the few longer ones the 64-bit word */
            b->order_slim_small.reserve(b->order_warp.size());
            for (uint32_t pr : b->order_warp) (b->m_of(pr) <= SLIM_MAX_M10 && b->n_of(pr) <= SLIM_MAX_SHORT ? b->order_slim_small : b->order_slim_rest).push_back(pr);
            if (b->order_slim_small.size() < 3 * b->order_slim_rest.size()) { b->order_slim_small.clear(); b->order_slim_rest.clear(); }
        }
        const double t4 = now_ms();
        /* FIFO uploads: host-side wait for this chunk's copies (kernels queued behind a device-side
         * event wait were seen to hold up the other workers' streams -- streams share hardware
         * queues); the copies queued behind these are not affected by when this thread wakes up */
        if (fifo) { const cudaError_t e = cudaEventSynchronize(ctx->ev_h2d); ctx->h2d_turn->release(); fifo_turn_held = false; CU(ctx, e); }
        if (b->wire && n_pairs) {
            expand_descs_kernel<<<(int)std::min<uint64_t>((n_pairs + 255) / 256, 1184), 256, 0, ctx->stream>>>((const WireDesc *)ctx->wire_dev.p, (uint32_t)n_pairs, (PairDesc *)b->d_descs);
            CU(ctx, cudaGetLastError());
        }
        if (!fifo) CU(ctx, wait_stream(ctx, ctx->stream));      /* (FIFO copies were waited for above; the kernels of run() follow on this stream) */
        if (dbg) fprintf(stderr, "[wfacuda] upload: validate+descs %.2f ms, alloc+seq h2d %.2f, descs h2d %.2f, binning %.2f, sync %.2f\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, now_ms() - t4);
        return 0;
    };
    if (body() != 0) { wfacuda_batch_free(ctx, b); return nullptr; }
    return b;
}

} // extern "C"

extern "C" {

int wfacuda_batch_run(wfacuda_ctx *ctx, wfacuda_batch *b)
{
    if (!ctx || !b) return fail(ctx, WFACUDA_E_INVALID, "NULL ctx or batch");
    const bool dbg = getenv("WFACUDA_DEBUG") != nullptr;
    const double t_begin = now_ms();
    {   /* per-run counters start from zero; the upload's H2D bytes are kept */
        const uint64_t h2d = b->ran ? 0 : ctx->stats.h2d_bytes;
        ctx->stats = wfacuda_stats{};
        ctx->stats.h2d_bytes = h2d;
    }
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = b->n_pairs;
    const uint32_t fills0 = g_fill_launches;
    int rc;
    if ((rc = ensure(ctx, ctx->ctr, sizeof(Counters)))) return rc;
    ctx->hc_cache_valid = false;
    CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(ctx, dev_fill(ctx->ctr.p, 0, sizeof(Counters), ctx->stream));
    CU(ctx, dev_fill(b->d_flags, 0, b->sz_flags, ctx->stream));
    if (n) CU(ctx, dev_fill(b->d_where, 0, n * 8, ctx->stream));
    /* results start as PENDING, or as the host-side verdict (EMPTY / TOO_LONG) */
    if (b->n_invalid) {
        std::vector<Result> init(n);
        memset(init.data(), 0, n * sizeof(Result));
        for (uint64_t i = 0; i < n; i++) init[i].status = b->host_status[i];
        if ((rc = staged_h2d(ctx, b->d_results, init.data(), n * sizeof(Result)))) return rc;
        CU(ctx, wait_stream(ctx, ctx->stream));     /* init goes out of scope */
    } else if (n) CU(ctx, dev_fill(b->d_results, 0xff, n * sizeof(Result), ctx->stream));
    const uint64_t n_valid = b->order_warp.size() + b->order_cta.size() + b->order_lane.size();
    if (n_valid) {
        if (b->max_len <= 160 && !getenv("WFACUDA_PACK_GENERIC")) {
            const int pack_blocks = (int)std::min<uint64_t>((n + 23) / 24, (uint64_t)ctx->sm_count * 16);   /* three pairs per warp */
            pack_short_kernel<<<pack_blocks, 256, 0, ctx->stream>>>((const PairDesc *)b->d_descs, (uint32_t)n, (const uint32_t *)b->d_raw,
                                                                   (uint32_t *)b->d_packed, (uint8_t *)b->d_flags);
        } else {
            const int pack_blocks = (int)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 16);      /* one warp per pair */
            pack_kernel<<<pack_blocks, 256, 0, ctx->stream>>>((const PairDesc *)b->d_descs, (uint32_t)n, (const uint32_t *)b->d_raw,
                                                             (uint32_t *)b->d_packed, (uint8_t *)b->d_flags);
        }
        CU(ctx, cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    CU(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    /* ops pool: first guess 1/4 op per base, grown on demand */
    uint64_t ops_cap = std::max<uint64_t>(b->seq_bases / 5 + 16 * n_valid + 1024, ctx->ops_pool.cap / 8);
    if ((rc = ensure(ctx, ctx->ops_pool, ops_cap * 8))) return rc;
    ops_cap = ctx->ops_pool.cap / 8;

    KParams P{};
    P.pairs = (const PairDesc *)b->d_descs; P.packed = (const uint32_t *)b->d_packed; P.raw = (const uint32_t *)b->d_raw;
    P.pflags = (const uint8_t *)b->d_flags; P.results = (Result *)b->d_results; P.ops_where = (uint64_t *)b->d_where;
    P.ops_cap = ops_cap;
    P.x = ctx->cfg.mismatch; P.oe = ctx->cfg.gap_open + ctx->cfg.gap_ext; P.e = ctx->cfg.gap_ext; P.g = ctx->g;
    P.xg = ctx->xg; P.oeg = ctx->oeg; P.eg = ctx->eg; P.dM = ctx->dM; P.dE = ctx->dE;
    P.global_aln = ctx->cfg.global_alignment ? 1 : 0; P.adaptive = ctx->cfg.adaptive ? 1 : 0;
    P.semi_literal = (ctx->cfg.flags & WFACUDA_FLAG_SEMIGLOBAL_LITERAL) ? 1 : 0;
    P.min_wf_len = (int32_t)std::min<uint32_t>(ctx->cfg.min_wf_len, 0x7fffffffu);
    P.max_dist_diff = (int32_t)std::min<uint32_t>(ctx->cfg.max_dist_diff, 0x7fffffffu);

    const double t_prep = now_ms();
    /* WARP class first (2-bit, then the pairs it handed back as 8-bit), then the CTA class */
    const bool force8 = ctx->cfg.flags & WFACUDA_FLAG_FORCE_8BIT;
    /* LANE class first (short global pairs, 2-bit only), then the WARP class incl. what the LANE
     * class handed over (2-bit, then 8-bit), then the CTA class */
    std::vector<uint32_t> to_warp, to_cta, warp8, cta8;
    /* everything after the first (big) launch of a batch has been waited for by the host, so
     * moving to the high-priority stream needs no event; restored (and drained) before returning */
    struct StreamGuard { wfacuda_ctx *c; ~StreamGuard() { if (c->stream != c->stream_main) { cudaStreamSynchronize(c->stream); c->stream = c->stream_main; } } } guard{ctx};
    ctx->lane_handed = 0;
    if ((rc = run_lane_class(ctx, b, b->order_lane, b->identity_cls == 2, P, &to_warp, &warp8))) return rc;
    if (!b->order_lane.empty()) ctx->stream = ctx->stream_hi;
    ctx->stats.pairs_lane = (uint32_t)(b->order_lane.size() - to_warp.size() - warp8.size() - ctx->lane_handed);
    std::vector<uint32_t> warp_extra;                       /* only built when the LANE class handed pairs over */
    if (!to_warp.empty()) { warp_extra = b->order_warp; warp_extra.insert(warp_extra.end(), to_warp.begin(), to_warp.end()); }
    const std::vector<uint32_t> &warp_order = to_warp.empty() ? b->order_warp : warp_extra;
    /* SLIM worker first (global, narrow wavefronts, penalties of the default shape); what outgrows
     * its ring goes on to the WARP worker */
    std::vector<uint32_t> slim_left;
    bool use_slim = slim_class_enabled(ctx) && b->max_len <= SLIM_MAX_M21 && !force8 && !warp_order.empty();
    /* without heuristic a row is as wide as the score allows: not worth a try beyond the ring's reach */
    if (use_slim && !ctx->cfg.adaptive && !ctx->slim_p_learned && estimate_slim(ctx, b->max_len, b->max_len, false).width > 32 * kSlimLadder[kSlimLadderN - 1]) use_slim = false;
    if (use_slim && !b->order_slim_small.empty() && to_warp.empty()) {
        /* (split by cell word at upload time) */
        if ((rc = run_class(ctx, b, b->order_slim_small, false, false, 2, P, &slim_left, &warp8, true, 0))) return rc;
        if ((rc = run_class(ctx, b, b->order_slim_rest, false, false, 2, P, &slim_left, &warp8, true, 1))) return rc;
        ctx->stats.pairs_slim = (uint32_t)(warp_order.size() - slim_left.size() - warp8.size());
        if ((rc = run_class(ctx, b, slim_left, false, false, 2, P, &to_cta, &warp8))) return rc;
    } else if (use_slim) {
        if ((rc = run_class(ctx, b, warp_order, b->identity_cls == 0 && to_warp.empty(), false, 2, P, &slim_left, &warp8, true))) return rc;
        ctx->stats.pairs_slim = (uint32_t)(warp_order.size() - slim_left.size() - warp8.size());
        if ((rc = run_class(ctx, b, slim_left, false, false, 2, P, &to_cta, &warp8))) return rc;
    } else
    if ((rc = run_class(ctx, b, warp_order, b->identity_cls == 0 && to_warp.empty(), false, force8 ? 8 : 2, P, &to_cta, &warp8))) return rc;
    if ((rc = run_class(ctx, b, warp8, false, false, 8, P, &to_cta, nullptr))) return rc;
    const double t_warp = now_ms();
    std::vector<uint32_t> cta_extra;
    if (!to_cta.empty()) { cta_extra = b->order_cta; cta_extra.insert(cta_extra.end(), to_cta.begin(), to_cta.end()); }
    const std::vector<uint32_t> &cta_order = to_cta.empty() ? b->order_cta : cta_extra;
    ctx->stats.pairs_warp = (uint32_t)(warp_order.size() + warp8.size() - to_cta.size() + ctx->lane_handed) - ctx->stats.pairs_slim;
    /* WIDE worker first (no heuristic, penalties of the default shape: clusters with the live rows on chip);
     * what it cannot hold goes on to the CTA worker */
    std::vector<uint32_t> wide_left;
    const bool use_wide = wide_class_enabled(ctx) && !cta_order.empty();
    if (use_wide) {
        if ((rc = run_wide_class(ctx, b, cta_order, P, &wide_left, &cta8))) return rc;
        ctx->stats.pairs_wide = (uint32_t)(cta_order.size() - wide_left.size() - cta8.size());
    }
    const std::vector<uint32_t> &cta_rest = use_wide ? wide_left : cta_order;
    ctx->stats.pairs_cta = (uint32_t)cta_rest.size();
    if ((rc = run_class(ctx, b, cta_rest, !use_wide && b->identity_cls == 1 && to_cta.empty(), true, force8 ? 8 : 2, P, nullptr, &cta8))) return rc;
    if ((rc = run_class(ctx, b, cta8, false, true, 8, P, nullptr, nullptr))) return rc;
    ctx->stats.pairs_8bit = force8 ? (uint32_t)n_valid : (uint32_t)(warp8.size() + cta8.size());
    CU(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));

    /* ops stay where the kernels put them (completion order): pair i's words start at
     * d_where[i] of the pool, and the pool's used prefix is what download() copies */
    Counters hc{};
    if (ctx->hc_cache_valid) hc = ctx->hc_cache;
    else if ((rc = fetch_small(ctx, &hc, ctx->ctr.p, sizeof hc))) return rc;
    CU(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    CU(ctx, wait_stream(ctx, ctx->stream));
    cudaEventElapsedTime(&ctx->stats.ms_pack, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.ms_align, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stats.ms_total_device, ctx->ev[0], ctx->ev[3]);
    ctx->stats.kernel_launches += g_fill_launches - fills0;              /* the device fills are kernels of ours too */
    ctx->stats.pairs = n_valid; ctx->stats.cells = hc.cells; ctx->stats.cells_written = hc.cells_written;
    ctx->stats.score_steps = hc.steps; ctx->stats.ops = hc.ops; ctx->stats.seq_bases = b->seq_bases;
    b->ops_total = hc.ops_cursor;
    ctx->dump_rows = hc.dump_rows;
    ctx->pool_owner = b;
    ctx->last_ops_total = b->ops_total;
    b->ran = true;
    if (dbg) fprintf(stderr, "[wfacuda] run: prep %.2f ms, warp class %.2f ms, rest %.2f ms | device: pack %.2f align %.2f total %.2f\n",
                     t_prep - t_begin, t_warp - t_prep, now_ms() - t_warp, ctx->stats.ms_pack, ctx->stats.ms_align, ctx->stats.ms_total_device);
    return 0;
}

int wfacuda_batch_download(wfacuda_ctx *ctx, wfacuda_batch *b, wfacuda_result *results,
                           uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off)
{
    if (!ctx || !b) return fail(ctx, WFACUDA_E_INVALID, "NULL ctx or batch");
    if (!b->ran) return fail(ctx, WFACUDA_E_INVALID, "batch has not been run");
    if (ops && ctx->pool_owner != b) return fail(ctx, WFACUDA_E_INVALID, "another batch ran on this ctx since: its ops replaced this batch's (download right after run)");
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = b->n_pairs;
    const double t0 = now_ms();
    int rc;
    if (n && results) if ((rc = staged_d2h(ctx, results, b->d_results, n * sizeof(Result), true))) return rc;
    if (n && ops_off) if ((rc = staged_d2h(ctx, ops_off, b->d_where, n * 8, true))) return rc;
    if (ops) {
        if (b->ops_total > ops_capacity) {
            cudaStreamSynchronize(ctx->stream);
            return fail(ctx, WFACUDA_E_OPS_CAPACITY, "ops buffer holds %llu words, %llu needed", (unsigned long long)ops_capacity, (unsigned long long)b->ops_total);
        }
        if (b->ops_total) if ((rc = staged_d2h(ctx, ops, ctx->ops_pool.p, b->ops_total * 8, true))) return rc;
    }
    CU(ctx, wait_stream(ctx, ctx->stream));       /* copies into page-locked caller memory were left in flight */
    if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda] download: %.2f ms for %.1f MB\n", now_ms() - t0, ctx->stats.d2h_bytes / 1e6);
    return 0;
}

/* CIGAR strings and alignment text of a batch that has been run (wfa_cigar.go:217-333), rendered
 * by the kernels of wfa_render.cuh from what the run left in HBM. */
int wfacuda_batch_render(wfacuda_ctx *ctx, wfacuda_batch *b, int only_aligned_region,
                         uint8_t *cigar, uint64_t cigar_capacity, uint64_t *cigar_off, uint32_t *cigar_len,
                         uint8_t *text, uint64_t text_capacity, uint64_t *text_off, uint32_t *text_len)
{
    if (!ctx || !b) return fail(ctx, WFACUDA_E_INVALID, "NULL ctx or batch");
    if (!b->ran || ctx->pool_owner != b) return fail(ctx, WFACUDA_E_INVALID, "render needs the batch that ran last on this ctx (its ops live in the ctx's pool)");
    const uint64_t n = b->n_pairs;
    if (n && (!cigar_off || !cigar_len)) return fail(ctx, WFACUDA_E_INVALID, "cigar_off / cigar_len are NULL");
    if (text && n && (!text_off || !text_len)) return fail(ctx, WFACUDA_E_INVALID, "text_off / text_len are NULL");
    CU(ctx, cudaSetDevice(ctx->device));
    ctx->render_cigar_total = ctx->render_text_total = 0;
    if (n == 0) return 0;
    int rc;
    /* meta: cursors[2] | cigar_off[n] | text_off[n] | cigar_len[n] | text_len[n] */
    if ((rc = ensure(ctx, ctx->render_meta, 16 + n * 24))) return rc;
    uint8_t *meta = (uint8_t *)ctx->render_meta.p;
    RenderParams R{};
    R.pairs = (const PairDesc *)b->d_descs; R.results = (const Result *)b->d_results;
    R.ops_pool = (const uint64_t *)ctx->ops_pool.p; R.ops_where = (const uint64_t *)b->d_where;
    R.raw = (const uint8_t *)b->d_raw; R.n_pairs = (uint32_t)n; R.only_aligned = only_aligned_region ? 1 : 0;
    R.cursors = (unsigned long long *)meta;
    R.cigar_off = (uint64_t *)(meta + 16); R.text_off = R.cigar_off + n;
    R.cigar_len = (uint32_t *)(R.text_off + n); R.text_len = R.cigar_len + n;
    CU(ctx, dev_fill(meta, 0, 16, ctx->stream));
    const int blocks = (int)((n + 255) / 256);
    render_measure_kernel<<<blocks, 256, 0, ctx->stream>>>(R);
    CU(ctx, cudaGetLastError());
    unsigned long long tot[2];
    if ((rc = fetch_small(ctx, tot, meta, 16))) return rc;
    ctx->render_cigar_total = tot[0]; ctx->render_text_total = tot[1];
    ctx->stats.kernel_launches++;
    if (tot[0] > cigar_capacity || (tot[0] && !cigar)) return fail(ctx, WFACUDA_E_OPS_CAPACITY, "cigar buffer holds %llu bytes, %llu needed", (unsigned long long)cigar_capacity, tot[0]);
    if (text && tot[1] > text_capacity) return fail(ctx, WFACUDA_E_OPS_CAPACITY, "text buffer holds %llu bytes, %llu needed", (unsigned long long)text_capacity, tot[1]);
    if ((rc = ensure(ctx, ctx->render_cigar, tot[0] + 16))) return rc;
    R.cigar = (uint8_t *)ctx->render_cigar.p; R.cigar_cap = tot[0];
    render_cigar_kernel<<<blocks, 256, 0, ctx->stream>>>(R);
    CU(ctx, cudaGetLastError());
    ctx->stats.kernel_launches++;
    if (text) {
        if ((rc = ensure(ctx, ctx->render_text, tot[1] + 16))) return rc;
        R.text = (uint8_t *)ctx->render_text.p; R.text_cap = tot[1];
        const int tblocks = (int)std::min<uint64_t>((n + 7) / 8, (uint64_t)ctx->sm_count * 16);
        render_text_kernel<<<tblocks, 256, 0, ctx->stream>>>(R);
        CU(ctx, cudaGetLastError());
        ctx->stats.kernel_launches++;
    }
    if ((rc = staged_d2h(ctx, cigar_off, R.cigar_off, n * 8))) return rc;
    if ((rc = staged_d2h(ctx, cigar_len, R.cigar_len, n * 4))) return rc;
    if (tot[0] && (rc = staged_d2h(ctx, cigar, R.cigar, tot[0]))) return rc;
    if (text) {
        if ((rc = staged_d2h(ctx, text_off, R.text_off, n * 8))) return rc;
        if ((rc = staged_d2h(ctx, text_len, R.text_len, n * 4))) return rc;
        if (tot[1] && (rc = staged_d2h(ctx, text, R.text, tot[1]))) return rc;
    }
    CU(ctx, wait_stream(ctx, ctx->stream));
    return 0;
}

/* Aligner.M / I / D after Align (wfa.go:80-86, 143-268): the whole wavefront store of ONE pair,
 * read back from the arena slot of the worker that aligned it (what Plot / Print /
 * GetRaw of the reference need).  Every existing score's row comes with its post-reduce range
 * [lo, hi] = WaveFront.Lo / Hi of M; cells outside it are absent (equivalence 1). */
int wfacuda_align_components(wfacuda_ctx *ctx, const uint8_t *q, uint32_t q_len, const uint8_t *t, uint32_t t_len,
                             wfacuda_result *result, uint64_t *ops, uint64_t ops_capacity,
                             wfacuda_wavefront *rows, uint32_t rows_capacity, uint32_t *n_rows,
                             uint32_t *cells, uint64_t cells_capacity, uint64_t *n_cells)
{
    if (!ctx || !q || !t || !result || !n_rows || !n_cells) return fail(ctx, WFACUDA_E_INVALID, "NULL argument");
    *n_rows = 0; *n_cells = 0;
    std::vector<uint8_t> pool((size_t)q_len + t_len + 32, 0);
    memcpy(pool.data(), q, q_len); memcpy(pool.data() + q_len, t, t_len);
    const uint64_t q_off = 0, t_off = q_len;
    const double keep_scale = ctx->arena_scale; const int keep_cap = ctx->ring_cap_learned;
    ctx->dump_mode = true;
    struct Guard { wfacuda_ctx *c; double s; int r; ~Guard() { c->dump_mode = false; c->arena_scale = s; c->ring_cap_learned = r; } } guard{ctx, keep_scale, keep_cap};
    wfacuda_batch *b = wfacuda_batch_upload(ctx, 1, pool.data(), &q_off, &q_len, &t_off, &t_len);
    if (!b) return ctx->last_rc ? ctx->last_rc : WFACUDA_E_CUDA;
    struct Free { wfacuda_ctx *c; wfacuda_batch *b; ~Free() { wfacuda_batch_free(c, b); } } fr{ctx, b};
    int rc = wfacuda_batch_run(ctx, b);
    if (rc) return rc;
    uint64_t off1 = 0;
    if ((rc = wfacuda_batch_download(ctx, b, result, ops, ops_capacity, &off1))) return rc;
    if (result->status != ST_OK) return 0;                     /* ErrEmptySeq / ErrSeqTooLong / resources: nothing to read */
    /* row headers the forward pass wrote: up to the final score (global), up to the global corner
     * (semi-global: this mode runs the literal scan, the reference keeps those rows too) */
    const uint32_t used_hdr = (uint32_t)ctx->dump_rows;
    const uint64_t slot_words = ctx->dump_slot_bytes / 4;
    if (used_hdr == 0 || (uint64_t)used_hdr * sizeof(RowHdr) > ctx->dump_slot_bytes) return fail(ctx, WFACUDA_E_CUDA, "no row headers reported for the pair");
    std::vector<RowHdr> hdr(used_hdr);
    if ((rc = staged_d2h(ctx, hdr.data(), ctx->arena.p, (size_t)used_hdr * sizeof(RowHdr)))) return rc;
    CU(ctx, wait_stream(ctx, ctx->stream));
    uint64_t min_off = slot_words, want_cells = 0; uint32_t want_rows = 0;
    for (uint32_t i = 0; i < used_hdr; i++) {
        const RowHdr &h = hdr[i];
        if (h.lo > h.hi) continue;
        want_rows++; want_cells += 3ull * (uint64_t)(h.hi - h.lo + 1);
        min_off = std::min<uint64_t>(min_off, h.off);
    }
    /* initComponents (wfa.go:155-183) also seeds M[x] -- the start cells whose first bases differ.
     * The kernels overlay those cells when they reach score x; an alignment that ends before that
     * (semi-global, a start cell matches all the way at score 0) never gets there, while the
     * reference keeps the seeded, un-extended wavefront: it is rebuilt here from the same rule. */
    std::vector<uint32_t> seed; int seed_lo = 1, seed_hi = 0;
    if (!ctx->cfg.global_alignment && used_hdr <= (uint32_t)ctx->xg) {
        const int nq = (int)q_len, nt = (int)t_len;
        for (int k = -(nq - 1); k <= nt - 1; k++)
            if (q[k < 0 ? -k : 0] != t[k > 0 ? k : 0]) { if (seed_lo > seed_hi) seed_lo = k; seed_hi = k; }
        if (seed_lo <= seed_hi) {
            seed.assign((size_t)(seed_hi - seed_lo + 1), 0u);
            for (int k = seed_lo; k <= seed_hi; k++)
                if (q[k < 0 ? -k : 0] != t[k > 0 ? k : 0]) seed[(size_t)(k - seed_lo)] = (uint32_t)((k > 0 ? k : 0) + 1) << T_BITS | T_MISMATCH;
            want_rows++; want_cells += 3ull * seed.size();
        }
    }
    *n_rows = want_rows; *n_cells = want_cells;
    if (want_rows > rows_capacity || want_cells > cells_capacity || (want_rows && (!rows || !cells)))
        return fail(ctx, WFACUDA_E_OPS_CAPACITY, "wavefront store has %u rows / %llu cell words, buffers hold %u / %llu", want_rows, (unsigned long long)want_cells, rows_capacity, (unsigned long long)cells_capacity);
    if (!want_rows) return 0;
    std::vector<uint32_t> slot(slot_words > min_off ? slot_words - min_off : 1);
    if (slot_words > min_off) {
        if ((rc = staged_d2h(ctx, slot.data(), (const uint32_t *)ctx->arena.p + min_off, slot.size() * 4))) return rc;
        CU(ctx, wait_stream(ctx, ctx->stream));
    }
    uint64_t at = 0; uint32_t r = 0;
    for (uint32_t i = 0; i < used_hdr; i++) {
        const RowHdr &h = hdr[i];
        if (h.lo > h.hi) continue;
        rows[r].score = i * ctx->g; rows[r].lo = h.lo; rows[r].hi = h.hi; rows[r].reserved_ = 0; rows[r].first_cell = at;
        const uint32_t *base = slot.data() + (h.off - min_off);
        for (int k = h.lo; k <= h.hi; k++) {
            const uint64_t idx = (uint64_t)(k - h.alo);
            for (int c = 0; c < 3; c++) cells[at++] = ctx->dump_cta ? base[(uint64_t)c * (uint64_t)h.aw + idx] : base[3 * idx + c];
        }
        r++;
    }
    if (!seed.empty()) {
        rows[r].score = ctx->cfg.mismatch; rows[r].lo = seed_lo; rows[r].hi = seed_hi; rows[r].reserved_ = 0; rows[r].first_cell = at;
        for (uint32_t v : seed) { cells[at++] = v; cells[at++] = 0u; cells[at++] = 0u; }
    }
    return 0;
}

/* Measured INT32 instruction-issue peaks of the ctx's device, in thread-level Tops/s: add / xor
 * only (ALU pipe) and add + mad.lo alternating (ALU + FMA pipes) -- the denominators of the INT32
 * roofline in bench.py. */
int wfacuda_measure_issue_peak(wfacuda_ctx *ctx, double *alu_tops, double *mixed_tops)
{
    if (!ctx || !alu_tops || !mixed_tops) return fail(ctx, WFACUDA_E_INVALID, "NULL argument");
    CU(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    DevBuf out;
    int rc = ensure(ctx, out, (size_t)blocks * threads * 4);
    if (rc) return rc;
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {                    /* first repetition warms up */
            cudaEventRecord(ctx->ev[0], ctx->stream);
            if (mode == 0) int32_peak_kernel<0><<<blocks, threads, 0, ctx->stream>>>((uint32_t *)out.p, iters);
            else           int32_peak_kernel<1><<<blocks, threads, 0, ctx->stream>>>((uint32_t *)out.p, iters);
            cudaEventRecord(ctx->ev[1], ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaFree(out.p); return fail(ctx, WFACUDA_E_CUDA, "int32 peak kernel failed: %s", cudaGetErrorString(cudaGetLastError())); }
            float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
            if (rep > 0) best = std::min(best, ms);
        }
        res[mode] = (double)blocks * threads * (double)iters * 32.0 / (best * 1e-3) / 1e12;
    }
    cudaFree(out.p);
    *alu_tops = res[0]; *mixed_tops = res[1];
    return 0;
}

void wfacuda_last_render_total(const wfacuda_ctx *ctx, uint64_t *cigar_bytes, uint64_t *text_bytes)
{
    if (cigar_bytes) *cigar_bytes = ctx ? ctx->render_cigar_total : 0;
    if (text_bytes) *text_bytes = ctx ? ctx->render_text_total : 0;
}

/* Chunk boundaries of the pipeline: cuts[0] = 0 < ... < cuts.back() = n_pairs.  Every worker
 * validates its chunk before it can queue the upload, so the first chunks are small and double in
 * size (C/8, C/4, C/2: the copy engine gets its first bytes after an eighth of a chunk's host work
 * and is never idle afterwards), and the last `tail_levels` ones shrink again (C/2, C/4, ...): what
 * is left to do when the last upload completes is one small chunk's kernels + download.  Batches
 * below six chunks, and tail_levels < 0, are cut uniformly.  Inner boundaries are multiples of 32
 * pairs (whole LANE groups).  Pure host logic (exported as wfacuda_chunk_plan for tests). */
static std::vector<uint64_t> plan_chunks(uint64_t n_pairs, uint64_t C, int tail_levels)
{
    std::vector<uint64_t> cuts{0};
    C = std::max<uint64_t>(C, 32);
    std::vector<uint64_t> head, tail;
    if (tail_levels >= 0 && n_pairs >= 6 * C && C >= 32768) {
        head = {C / 8, C / 4, C / 2};
        for (int l = 1; l <= tail_levels; l++) tail.push_back(C >> l);
    }
    uint64_t used = 0, tail_total = 0;
    for (uint64_t h : head) used += h;
    for (uint64_t t : tail) { used += t; tail_total += t & ~31ull; }
    const uint64_t rem = n_pairs - used, nm = std::max<uint64_t>(1, (rem + C - 1) / C);
    for (uint64_t h : head) cuts.push_back(cuts.back() + (h & ~31ull));
    /* the odd pairs (n_pairs mod 32) go to the very last chunk, so that every inner boundary is a multiple of 32 */
    const uint64_t mid0 = cuts.back(), mid_total = tail.empty() ? n_pairs - mid0 : ((n_pairs - mid0 - tail_total) & ~31ull);
    for (uint64_t j = 1; j <= nm; j++) cuts.push_back(j == nm ? mid0 + mid_total : mid0 + ((mid_total * j / nm) & ~31ull));
    for (size_t j = 0; j < tail.size(); j++) cuts.push_back(j + 1 == tail.size() ? n_pairs : cuts.back() + (tail[j] & ~31ull));
    /* drop empty chunks (tiny batches) */
    std::vector<uint64_t> out{0};
    for (size_t j = 1; j < cuts.size(); j++) if (cuts[j] > out.back()) out.push_back(cuts[j]);
    if (out.back() != n_pairs) out.push_back(n_pairs);
    return out;
}

int wfacuda_wide_plan(uint64_t max_diagonals, uint32_t seq_entries, uint64_t smem_per_cta,
                      int *cluster_ctas, uint32_t *diagonals_per_cta, int *threads, uint64_t *smem_bytes)
{
    if (!cluster_ctas || !diagonals_per_cta || !threads || !smem_bytes || max_diagonals == 0) return WFACUDA_E_INVALID;
    const uint64_t head = (uint64_t)WIDE_HEAD_BYTES + 64;
    for (int C = 1; C <= WIDE_MAX_CLUSTER; C *= 2) {
        const uint64_t seg = ((max_diagonals + C - 1) / C + 63) & ~63ull;              /* whole warps of column pairs */
        if (seg > 0xffffffc0ull) continue;
        const uint64_t smem = wide_smem_bytes((uint32_t)seg, seq_entries);
        if (smem + head <= smem_per_cta) {
            *cluster_ctas = C; *diagonals_per_cta = (uint32_t)seg; *smem_bytes = smem;
            *threads = (int)std::min<uint64_t>(1024, std::max<uint64_t>(128, ((seg / 2 + 31) / 32) * 32));
            return 0;
        }
    }
    return WFACUDA_E_INVALID;
}

int wfacuda_chunk_plan(uint64_t n_pairs, uint64_t chunk_pairs, int tail_levels, uint64_t *cuts, uint32_t cuts_capacity, uint32_t *n_cuts)
{
    if (!n_cuts) return WFACUDA_E_INVALID;
    const std::vector<uint64_t> c = plan_chunks(n_pairs, chunk_pairs, tail_levels);
    *n_cuts = (uint32_t)c.size();
    if (c.size() > cuts_capacity || !cuts) return WFACUDA_E_OPS_CAPACITY;
    for (size_t i = 0; i < c.size(); i++) cuts[i] = c[i];
    return 0;
}

static int align_batch_single(wfacuda_ctx *ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                              const uint64_t *q_off, const uint32_t *q_len, const uint64_t *t_off, const uint32_t *t_len,
                              wfacuda_result *results, uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off)
{
    wfacuda_batch *b = wfacuda_batch_upload(ctx, n_pairs, seq_bytes, q_off, q_len, t_off, t_len);
    if (!b) return ctx->last_rc ? ctx->last_rc : WFACUDA_E_CUDA;
    int rc = wfacuda_batch_run(ctx, b);
    if (rc == 0) rc = wfacuda_batch_download(ctx, b, results, ops, ops_capacity, ops_off);
    wfacuda_batch_free(ctx, b);
    return rc;
}

static void add_stats(wfacuda_stats &a, const wfacuda_stats &s)
{
    a.pairs += s.pairs; a.cells += s.cells; a.cells_written += s.cells_written; a.score_steps += s.score_steps;
    a.ops += s.ops; a.seq_bases += s.seq_bases; a.arena_bytes = std::max(a.arena_bytes, s.arena_bytes);
    a.h2d_bytes += s.h2d_bytes; a.d2h_bytes += s.d2h_bytes; a.kernel_launches += s.kernel_launches;
    a.align_launches += s.align_launches; a.retries += s.retries; a.pairs_warp += s.pairs_warp;
    a.pairs_cta += s.pairs_cta; a.pairs_8bit += s.pairs_8bit; a.pairs_lane += s.pairs_lane; a.pairs_slim += s.pairs_slim; a.pairs_wide += s.pairs_wide; a.ms_pack += s.ms_pack; a.ms_align += s.ms_align;
    a.ms_total_device += s.ms_total_device;
}

/* Large batches are cut into chunks that flow through a few worker contexts on the same
 * device (own stream, pinned staging and arena each), so that host staging, H2D, kernels
 * and D2H of different chunks overlap.  Each chunk's ops land in one contiguous region of
 * the caller's buffer, claimed when the chunk's op count is known. */
int wfacuda_align_batch(wfacuda_ctx *ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                        const uint64_t *q_off, const uint32_t *q_len, const uint64_t *t_off, const uint32_t *t_len,
                        wfacuda_result *results, uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off)
{
    if (!ctx) return fail(nullptr, WFACUDA_E_INVALID, "ctx is NULL");
    if (n_pairs && !results) return fail(ctx, WFACUDA_E_INVALID, "results is NULL");
    if (n_pairs && (!seq_bytes || !q_off || !q_len || !t_off || !t_len)) return fail(ctx, WFACUDA_E_INVALID, "NULL input array");
    const uint64_t kMinChunk = 32768;
    uint64_t sample_bytes = 0;
    const uint64_t sample_n = std::max<uint64_t>(1, std::min<uint64_t>(n_pairs, 4096));
    for (uint64_t i = 0; i < sample_n && n_pairs; i++) sample_bytes += (uint64_t)q_len[i * (n_pairs / sample_n)] + t_len[i * (n_pairs / sample_n)];
    const double mean_bytes = std::max(1.0, (double)sample_bytes / (double)sample_n);
    /* A few hundred LONG pairs (config 5 as one GPU of eight sees it: 1 250 pairs of 100 kbp, 250 MB in, 230 MB of
     * ops out): one warp per pair, every pair's 40 k dependent score steps decide the kernel time, and the device
     * is far from full -- so four chunks whose kernels run side by side cost no more kernel time than one, while
     * their uploads and downloads overlap the other chunks' kernels. */
    const bool long_pairs = n_pairs >= 256 && mean_bytes >= 32768.0 && !getenv("WFACUDA_NO_LONG_PIPELINE");
    if ((n_pairs < 2 * kMinChunk && !long_pairs) || getenv("WFACUDA_NO_PIPELINE")) {
        int rc = align_batch_single(ctx, n_pairs, seq_bytes, q_off, q_len, t_off, t_len, results, ops, ops_capacity, ops_off);
        if (rc == 0 || rc == WFACUDA_E_OPS_CAPACITY) return rc;
        return rc;
    }
    /* chunk size: about 20 MB of sequence (0.45 ms of PCIe; ~one full wave of the LANE kernel for 150 bp reads), at least kMinChunk pairs */
    /* Pageable (or scattered: per-pair heap strings) caller memory is staged by the workers themselves -- a host copy
     * into their page-locked pools before the chunk can be uploaded.  All workers stage at once and share the host's
     * memory bandwidth, so with one chunk per worker every chunk is ready at about the same time and the uploads, which
     * go out in chunk order, only start when the staging is over (the reference's call shape: first full chunk uploaded
     * at 9.5 ms of 16.9).  Smaller chunks, several per worker, finish their staging in chunk order and the uploads
     * overlap the staging of the later ones: a dense pageable pool 48 -> 72-82 M alignments/s on config 2, per-pair
     * heap strings 18.5 -> 17.2 ms per million pairs (those are bound by the per-string copies themselves).  Letting
     * only a few workers stage at a time, in chunk order, was tried on top and changed nothing measurable. */
    const bool src_pinned = is_pinned(seq_bytes);
    double chunk_bytes = src_pinned ? 20e6 : 8e6;
    if (const char *e = getenv("WFACUDA_CHUNK_MB")) chunk_bytes = std::max(1.0, atof(e)) * 1e6;
    uint64_t chunk_pairs = std::max<uint64_t>(src_pinned ? kMinChunk : kMinChunk / 2, std::min<uint64_t>(262144, (uint64_t)(chunk_bytes / mean_bytes)));
    if (long_pairs) chunk_pairs = std::max<uint64_t>(32, (((n_pairs + 3) / 4 + 31) / 32) * 32);
    if (const char *e = getenv("WFACUDA_CHUNK_PAIRS")) chunk_pairs = std::max<uint64_t>(long_pairs ? 32 : 1024, strtoull(e, nullptr, 10));
    int tail_levels = 2;
    if (const char *e = getenv("WFACUDA_TAIL_LEVELS")) tail_levels = std::max(0, std::min(5, atoi(e)));
    const std::vector<uint64_t> cuts = plan_chunks(n_pairs, chunk_pairs, getenv("WFACUDA_UNIFORM_CHUNKS") ? -1 : tail_levels);
    const uint64_t n_chunks = cuts.size() - 1;
    const unsigned hw = ctx->core_share ? ctx->core_share : std::max(2u, std::thread::hardware_concurrency());
    /* one worker per chunk in flight: enough of them that uploads (PCIe-bound, taken in turns)
     * never wait for a worker that is still computing or downloading */
    /* (pageable memory is staged by the workers: host copies, one core each; with few cores per device
     * still enough workers to overlap copies and kernels -- they then sleep in their waits, see below) */
    unsigned kmax = std::min<unsigned>(16, std::max(6u, hw - 2));
    if (const char *e = getenv("WFACUDA_PIPE_WORKERS")) kmax = std::max(1, atoi(e));
    const int K = (int)std::min<uint64_t>(kmax, n_chunks);
    if ((int)ctx->subs.size() < K && ctx->arena.p) {
        /* the workers need the room: drop this ctx's own arena (re-grown on demand) */
        cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->arena.p); ctx->arena.p = nullptr; ctx->arena.cap = 0;
    }
    /* the workers created now split 70 % of what is free NOW in equal shares (their arenas grow on
     * demand, so free memory barely moves while they are created: dividing by the number still to
     * create, as an earlier version did, handed out 2.6x the free memory over 14 workers and ran
     * out of memory on config 3 at its full size); the rest is for their batch buffers and ops pools */
    size_t free_now = 0, total_now = 0;
    const int to_create = K - (int)ctx->subs.size();
    if (to_create > 0) {
        cudaSetDevice(ctx->device);
        if (cudaMemGetInfo(&free_now, &total_now) != cudaSuccess) { cudaGetLastError(); free_now = 8ull << 30; }
    }
    while ((int)ctx->subs.size() < K) {
        wfacuda_config c = ctx->cfg;
        uint64_t share = (uint64_t)(0.70 * (double)free_now) / (uint64_t)to_create;
        if (ctx->cfg.arena_budget_bytes) share = std::min<uint64_t>(share, ctx->cfg.arena_budget_bytes / (uint64_t)K);
        c.arena_budget_bytes = std::max<uint64_t>(share, 64u << 20);
        wfacuda_ctx *sub = wfacuda_create(ctx->device, &c);
        if (!sub) return fail(ctx, WFACUDA_E_CUDA, "pipeline worker: %s", g_tls_error.c_str());
        sub->h2d_turn = &ctx->h2d_turns;
        sub->h2d_parent = ctx;
        if (cudaEventCreateWithFlags(&sub->ev_h2d, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) { wfacuda_destroy(sub); return fail(ctx, WFACUDA_E_CUDA, "pipeline worker: event creation failed"); }
        ctx->subs.push_back(sub);
    }
    if (!ctx->h2d_fifo && cudaStreamCreateWithFlags(&ctx->h2d_fifo, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); ctx->h2d_fifo = nullptr; }
    {
        /* more workers than this process can expect cores for: sleep in the waits instead of spinning
         * (WFACUDA_BLOCKING_SYNC=0/1 decides otherwise; a launcher that runs several ranks per box sets it) */
        bool blocking = (unsigned)K + 1 > hw;
        if (const char *e = getenv("WFACUDA_BLOCKING_SYNC")) blocking = atoi(e) != 0;
        for (wfacuda_ctx *sub : ctx->subs) sub->blocking_sync = blocking;
    }
    /* uploads in flight: with the shared FIFO stream two keep the copy engine fed back to back
     * (the next copy is already queued when one completes) without letting chunks arrive late */
    ctx->h2d_turns.free_slots = ctx->h2d_fifo && !getenv("WFACUDA_NO_FIFO") ? 2 : 1;
    ctx->h2d_turns.next_ticket = 0;
    if (const char *e = getenv("WFACUDA_H2D_TURNS")) ctx->h2d_turns.free_slots = std::max(1, atoi(e));
    std::atomic<uint64_t> cursor{0};
    std::atomic<int> first_err{0};
    std::mutex mu;
    wfacuda_stats total{};
    std::string err_text;
    const double t_begin = now_ms();
    g_dbg_t0 = t_begin;
    std::vector<double> t_up(K, 0.0), t_run(K, 0.0), t_down(K, 0.0);
    auto work = [&](int k) {
        wfacuda_ctx *sub = ctx->subs[k];
        /* chunk c belongs to worker c mod K, call after call: every worker sees the same chunk
         * sizes again, so none of its device buffers (arena, staging) is ever re-allocated after
         * the first call -- a cudaMalloc / cudaFree in the middle of a batch stalls the whole device */
        for (uint64_t c = (uint64_t)k; c < n_chunks; c += (uint64_t)K) {
            if (first_err.load()) { ctx->h2d_turns.pass(n_chunks); break; }      /* nobody may wait for a chunk that will not come */
            const uint64_t a = cuts[c], cnt = cuts[c + 1] - cuts[c];
            const double w0 = now_ms();
            sub->h2d_ticket = getenv("WFACUDA_UNORDERED_UPLOADS") ? -1 : (int64_t)c;
            wfacuda_batch *b = wfacuda_batch_upload(sub, cnt, seq_bytes, q_off + a, q_len + a, t_off + a, t_len + a);
            ctx->h2d_turns.pass(c);            /* whatever happened in there, later chunks may go */
            sub->h2d_ticket = -1;
            const double w1 = now_ms();
            int rc = b ? wfacuda_batch_run(sub, b) : (sub->last_rc ? sub->last_rc : WFACUDA_E_CUDA);
            const double w2 = now_ms();
            t_up[k] += w1 - w0; t_run[k] += w2 - w1;
            if (rc == 0) {
                const uint64_t tot = b->ops_total, base = cursor.fetch_add(tot);
                const bool fits = ops && base + tot <= ops_capacity;
                rc = wfacuda_batch_download(sub, b, results + a, fits ? ops + base : nullptr, fits ? tot : 0, ops_off ? ops_off + a : nullptr);
                if (rc == 0 && ops_off) for (uint64_t i = 0; i < cnt; i++) ops_off[a + i] += base;
            }
            t_down[k] += now_ms() - w2;
            if (getenv("WFACUDA_DEBUG")) fprintf(stderr, "[wfacuda]   chunk %2llu worker %d: start %.2f uploaded %.2f ran %.2f downloaded %.2f (ms since call)\n", (unsigned long long)c, k, w0 - t_begin, w1 - t_begin, w2 - t_begin, now_ms() - t_begin);
            {
                std::lock_guard<std::mutex> lk(mu);
                add_stats(total, sub->stats);
                if (rc != 0 && first_err.load() == 0) { first_err.store(rc); err_text = sub->err; }
                if (rc != 0) ctx->h2d_turns.pass(n_chunks);
            }
            if (b) wfacuda_batch_free(sub, b);
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k + 1 < K; k++) th.emplace_back(work, k);     /* the calling thread starts last: it takes the last worker's chunks */
    work(K - 1);
    for (auto &t : th) t.join();
    ctx->stats = total;
    ctx->last_ops_total = cursor.load();
    if (getenv("WFACUDA_DEBUG")) {
        fprintf(stderr, "[wfacuda] align_batch: %llu chunks of up to %llu pairs on %d workers, %.2f ms; per worker upload/run/download ms:",
                (unsigned long long)n_chunks, (unsigned long long)chunk_pairs, K, now_ms() - t_begin);
        for (int k = 0; k < K; k++) fprintf(stderr, " %.1f/%.1f/%.1f", t_up[k], t_run[k], t_down[k]);
        fprintf(stderr, "\n");
    }
    if (first_err.load()) return fail(ctx, first_err.load(), "%s", err_text.c_str());
    if (ops && cursor.load() > ops_capacity)
        return fail(ctx, WFACUDA_E_OPS_CAPACITY, "ops buffer holds %llu words, %llu needed", (unsigned long long)ops_capacity, (unsigned long long)cursor.load());
    return 0;
}

/* Contiguous index ranges of equal estimated cost: cost ~ (n+m) * band, band ~ (n+m) without
 * heuristic (work grows with the square of the edit count), constant with wf-adaptive.
 * cuts[0..n_shards]: shard d owns pairs [cuts[d], cuts[d+1]).  Pure host logic. */
int wfacuda_shard_plan(int n_shards, uint64_t n_pairs, const uint32_t *q_len, const uint32_t *t_len,
                       int adaptive, uint64_t *cuts)
{
    if (n_shards < 1 || !cuts || (n_pairs && (!q_len || !t_len))) return WFACUDA_E_INVALID;
    std::vector<double> pre(n_pairs + 1, 0.0);
    for (uint64_t i = 0; i < n_pairs; i++) {
        const double nm = (double)q_len[i] + t_len[i];
        pre[i + 1] = pre[i] + (adaptive ? nm : nm * nm) + 64.0;
    }
    cuts[0] = 0; cuts[n_shards] = n_pairs;
    for (int d = 1; d < n_shards; d++)
        cuts[d] = (uint64_t)(std::lower_bound(pre.begin(), pre.end(), pre[n_pairs] * d / n_shards) - pre.begin());
    for (int d = 1; d <= n_shards; d++) cuts[d] = std::max(cuts[d], cuts[d - 1]);
    return 0;
}

/* Estimated work of one pair: the wavefront band grows with the score without
 * heuristic and in semi-global mode (cost ~ (n+m)^2), and is bounded by wf-adaptive (~ n+m). */
static inline double pair_cost(uint32_t n, uint32_t m, bool adaptive, bool global_aln)
{
    const double nm = (double)n + m;
    return ((adaptive && global_aln) ? nm * 64.0 : nm * nm) + 4096.0;
}

/* Length-binned LPT: pairs are binned by n+m (half-octave bins), bins are dealt out from the
 * longest to the shortest, and inside a bin every shard gets ONE run of consecutive pairs, sized so
 * that the shards' estimated loads level out (water-filling).  A batch of equal-length reads ends
 * up as n_shards contiguous index ranges; a mixed batch as at most 64 runs per shard.  shard_of[i]
 * = shard of pair i; shard_cost (optional) = estimated load per shard.  Pure host logic. */
int wfacuda_shard_assign(int n_shards, uint64_t n_pairs, const uint32_t *q_len, const uint32_t *t_len,
                         int adaptive, int global_alignment, uint32_t *shard_of, double *shard_cost)
{
    if (n_shards < 1 || (n_pairs && (!q_len || !t_len || !shard_of))) return WFACUDA_E_INVALID;
    std::vector<double> load(n_shards, 0.0);
    auto bin_of = [&](uint64_t i) {
        const uint64_t nm = (uint64_t)q_len[i] + t_len[i];
        const int lg = 63 - __builtin_clzll(nm | 1);
        return 2 * lg + (int)((nm >> (lg > 0 ? lg - 1 : 0)) & 1);
    };
    double bin_cost[130] = {0.0};
    uint64_t bin_n[130] = {0};
    for (uint64_t i = 0; i < n_pairs; i++) {
        const int b = bin_of(i);
        bin_cost[b] += pair_cost(q_len[i], t_len[i], adaptive != 0, global_alignment != 0); bin_n[b]++;
    }
    /* per bin: which shard is being filled and how much of its share is left */
    struct Fill { int shard; double left; std::vector<double> share; };
    std::vector<Fill> fill(130);
    for (int b = 129; b >= 0; b--) {
        if (!bin_n[b]) continue;
        double total = bin_cost[b];
        for (double l : load) total += l;
        /* water level: shards above it get nothing from this bin */
        std::vector<int> order(n_shards);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int a, int c) { return load[a] < load[c]; });
        double level = 0.0; int used = n_shards;
        for (;;) {
            double sum = bin_cost[b];
            for (int j = 0; j < used; j++) sum += load[order[j]];
            level = sum / used;
            if (used > 1 && load[order[used - 1]] > level) used--; else break;
        }
        Fill &f = fill[b];
        f.share.assign(n_shards, 0.0);
        for (int j = 0; j < used; j++) f.share[order[j]] = std::max(0.0, level - load[order[j]]);
        for (int d = 0; d < n_shards; d++) load[d] += f.share[d];
        f.shard = 0;
        while (f.shard < n_shards - 1 && f.share[f.shard] <= 0.0) f.shard++;
        f.left = f.share[f.shard];
    }
    for (uint64_t i = 0; i < n_pairs; i++) {
        Fill &f = fill[bin_of(i)];
        const double c = pair_cost(q_len[i], t_len[i], adaptive != 0, global_alignment != 0);
        while (f.left < 0.5 * c && f.shard < n_shards - 1) {
            int nx = f.shard + 1;
            while (nx < n_shards - 1 && f.share[nx] <= 0.0) nx++;
            if (f.share[nx] <= 0.0 && nx == n_shards - 1) break;        /* nobody left to take it: stays with this shard */
            f.shard = nx; f.left += f.share[nx];
        }
        shard_of[i] = (uint32_t)f.shard; f.left -= c;
    }
    if (shard_cost) {
        for (int d = 0; d < n_shards; d++) shard_cost[d] = 0.0;
        for (uint64_t i = 0; i < n_pairs; i++) shard_cost[shard_of[i]] += pair_cost(q_len[i], t_len[i], adaptive != 0, global_alignment != 0);
    }
    return 0;
}

/* One batch over several devices (one ctx each): length-binned LPT shards (wfacuda_shard_assign),
 * one host thread per device, each running the chunked pipeline of wfacuda_align_batch on its
 * shard; results come back at the caller's indices.  No collective: pairs are independent
 * (reference contract: one Aligner per goroutine, wfa.go:73-78).  The ops buffer is cut into one
 * region per device in proportion to the shards' bases; ops_off[] are absolute positions, so the
 * regions need not be compacted. */
int wfacuda_align_batch_multi(wfacuda_ctx *const *ctxs, int n_ctx, uint64_t n_pairs, const uint8_t *seq_bytes,
                              const uint64_t *q_off, const uint32_t *q_len, const uint64_t *t_off, const uint32_t *t_len,
                              wfacuda_result *results, uint64_t *ops, uint64_t ops_capacity, uint64_t *ops_off)
{
    if (!ctxs || n_ctx < 1) return fail(nullptr, WFACUDA_E_INVALID, "need at least one ctx");
    for (int d = 0; d < n_ctx; d++) if (!ctxs[d]) return fail(nullptr, WFACUDA_E_INVALID, "ctx %d is NULL", d);
    if (n_ctx == 1) return wfacuda_align_batch(ctxs[0], n_pairs, seq_bytes, q_off, q_len, t_off, t_len, results, ops, ops_capacity, ops_off);
    if (n_pairs && (!seq_bytes || !q_off || !q_len || !t_off || !t_len || !results)) return fail(ctxs[0], WFACUDA_E_INVALID, "NULL input array");
    const wfacuda_config &cfg = ctxs[0]->cfg;
    std::vector<std::vector<uint32_t>> idx(n_ctx);
    std::vector<uint64_t> bases(n_ctx, 0), range_a(n_ctx, 0), range_n(n_ctx, 0);
    /* Reads of one length class (the usual batch): the LPT of wfacuda_shard_assign is n_ctx equal,
     * contiguous index ranges -- found with one cheap pass over the lengths (a few host threads),
     * no per-pair shard table, no index lists. */
    bool one_bin = n_pairs > 0;
    {
        const unsigned T = n_pairs < 200000 ? 1u : std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
        std::vector<int> lo_bin(T, 1 << 30), hi_bin(T, -1);
        std::vector<uint64_t> sum(T * (size_t)n_ctx, 0);
        auto scan = [&](unsigned k) {
            for (int d = 0; d < n_ctx; d++) {
                /* thread k scans its part of every shard's range, so that the per-shard byte sums fall out too */
                const uint64_t a = n_pairs * d / n_ctx & ~31ull, e = d + 1 == n_ctx ? n_pairs : (n_pairs * (d + 1) / n_ctx & ~31ull);
                uint64_t s_ = 0; int lo = lo_bin[k], hi = hi_bin[k];
                for (uint64_t i = a + (e - a) * k / T; i < a + (e - a) * (k + 1) / T; i++) {
                    const uint64_t nm = (uint64_t)q_len[i] + t_len[i];
                    const int lg = 63 - __builtin_clzll(nm | 1), b = 2 * lg + (int)((nm >> (lg > 0 ? lg - 1 : 0)) & 1);
                    lo = std::min(lo, b); hi = std::max(hi, b); s_ += nm + 16;
                }
                lo_bin[k] = lo; hi_bin[k] = hi; sum[k * (size_t)n_ctx + d] = s_;
            }
        };
        std::vector<std::thread> th;
        for (unsigned k = 1; k < T; k++) th.emplace_back(scan, k);
        scan(0);
        for (auto &t : th) t.join();
        for (unsigned k = 0; k < T; k++) if (lo_bin[k] != lo_bin[0] || hi_bin[k] != hi_bin[0] || lo_bin[k] != hi_bin[k]) one_bin = false;
        if (one_bin) for (int d = 0; d < n_ctx; d++) {
            range_a[d] = n_pairs * d / n_ctx & ~31ull;
            range_n[d] = (d + 1 == n_ctx ? n_pairs : (n_pairs * (d + 1) / n_ctx & ~31ull)) - range_a[d];
            for (unsigned k = 0; k < T; k++) bases[d] += sum[k * (size_t)n_ctx + d];
        }
    }
    if (!one_bin) {
        std::vector<uint32_t> shard_of(n_pairs);
        wfacuda_shard_assign(n_ctx, n_pairs, q_len, t_len, cfg.adaptive, cfg.global_alignment, shard_of.data(), nullptr);
        for (uint64_t i = 0; i < n_pairs; i++) { idx[shard_of[i]].push_back((uint32_t)i); bases[shard_of[i]] += (uint64_t)q_len[i] + t_len[i] + 16; }
        for (int d = 0; d < n_ctx; d++) {
            range_n[d] = idx[d].size();
            if (range_n[d] && (uint64_t)idx[d].back() - idx[d].front() + 1 == range_n[d]) { range_a[d] = idx[d].front(); idx[d].clear(); }   /* one run: no gather needed */
        }
    }
    uint64_t bases_total = 0;
    for (uint64_t v : bases) bases_total += v;
    /* ops regions: proportional to the shards' bases */
    std::vector<uint64_t> obase(n_ctx + 1, 0);
    {
        uint64_t run = 0;
        for (int d = 0; d < n_ctx; d++) { obase[d] = bases_total ? (uint64_t)((long double)ops_capacity * run / bases_total) : 0; run += bases[d]; }
        obase[n_ctx] = ops_capacity;
    }
    std::vector<int> rcs(n_ctx, 0);
    std::vector<uint64_t> used(n_ctx, 0);
    std::vector<std::thread> th;
    auto work = [&](int d) {
        const std::vector<uint32_t> &ix = idx[d];
        const uint64_t cnt = range_n[d];
        if (!cnt) { ctxs[d]->stats = wfacuda_stats{}; return; }
        uint64_t *ops_d = ops ? ops + obase[d] : nullptr;
        const uint64_t cap_d = ops ? obase[d + 1] - obase[d] : 0;
        /* the devices of one process share its cores: each pipeline gets its share of them */
        ctxs[d]->core_share = std::max(2u, std::thread::hardware_concurrency() / (unsigned)n_ctx);
        struct Share { wfacuda_ctx *c; ~Share() { c->core_share = 0; } } share{ctxs[d]};
        if (ix.empty()) {
            const uint64_t a = range_a[d];
            rcs[d] = wfacuda_align_batch(ctxs[d], cnt, seq_bytes, q_off + a, q_len + a, t_off + a, t_len + a, results + a, ops_d, cap_d, ops_off ? ops_off + a : nullptr);
            if (rcs[d] == 0 && ops_off) for (uint64_t j = 0; j < cnt; j++) ops_off[a + j] += obase[d];
        } else {
            std::vector<uint64_t> qo(cnt), to(cnt), oo(cnt);
            std::vector<uint32_t> ql(cnt), tl(cnt);
            std::vector<wfacuda_result> rr(cnt);
            for (uint64_t j = 0; j < cnt; j++) { const uint32_t i = ix[j]; qo[j] = q_off[i]; to[j] = t_off[i]; ql[j] = q_len[i]; tl[j] = t_len[i]; }
            rcs[d] = wfacuda_align_batch(ctxs[d], cnt, seq_bytes, qo.data(), ql.data(), to.data(), tl.data(), rr.data(), ops_d, cap_d, oo.data());
            if (rcs[d] == 0) for (uint64_t j = 0; j < cnt; j++) { const uint32_t i = ix[j]; results[i] = rr[j]; if (ops_off) ops_off[i] = oo[j] + obase[d]; }
        }
        used[d] = ctxs[d]->last_ops_total;
    };
    for (int d = 0; d + 1 < n_ctx; d++) th.emplace_back(work, d);
    work(n_ctx - 1);
    for (auto &t : th) t.join();
    int rc = 0;
    for (int d = 0; d < n_ctx; d++) if (rcs[d] && rcs[d] != WFACUDA_E_OPS_CAPACITY && rc == 0) { rc = rcs[d]; g_tls_error = ctxs[d]->err; }
    /* capacity: what the whole buffer must hold so that every device's region does (its share is
     * proportional to its bases); on success, the extent of the buffer in use */
    uint64_t need = 0, extent = 0; bool short_of = false;
    for (int d = 0; d < n_ctx; d++) {
        if (rcs[d] == WFACUDA_E_OPS_CAPACITY) short_of = true;
        if (bases[d]) need = std::max<uint64_t>(need, (uint64_t)((long double)used[d] * bases_total / bases[d]) + (uint64_t)n_ctx + 64);
        if (used[d]) extent = std::max(extent, obase[d] + used[d]);
    }
    for (int d = 0; d < n_ctx; d++) ctxs[d]->last_ops_total = short_of ? need : extent;
    if (rc == 0 && short_of && ops)
        rc = fail(ctxs[0], WFACUDA_E_OPS_CAPACITY, "ops buffer holds %llu words, %llu needed (one region per device)", (unsigned long long)ops_capacity, (unsigned long long)need);
    return rc;
}

} // extern "C"
