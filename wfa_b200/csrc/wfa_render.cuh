/*
 * wfa_render.cuh -- CIGAR strings and alignment text of a whole batch on the GPU.
 *
 * (*AlignmentResult).CIGAR(onlyAignedRegion)            wfa_cigar.go:236-257
 * (*AlignmentResult).AlignmentText(q, t, onlyAigned..)  wfa_cigar.go:261-333
 * trimOps                                               wfa_cigar.go:217-233
 *
 * Input is what a run leaves in HBM: results (n_ops, begin coordinates), the ops pool in
 * completion order with where[pair], the caller's bytes.  Output strings are placed like the
 * ops are: a warp reserves the bytes of its 32 pairs with one atomicAdd per buffer and every
 * pair remembers its offset, nothing is sorted afterwards.
 *   cigar   decimal count + op letter per op, no terminator
 *   text    three lines of equal length L per pair, at off, off + L, off + 2 L:
 *           Q (query bases, '-' under an insertion), A ('|' under a match, ' ' elsewhere),
 *           T (target bases, '-' under a deletion / 'H')
 * With onlyAignedRegion the ops are trimmed to [first M, last M] and the sequences to
 * q[QBegin-1 : QEnd], t[TBegin-1 : TEnd].  An alignment without any M has no aligned region:
 * the reference slices ops[-1:0] there (a run-time panic); here the strings are empty.
 */
#pragma once
#include "wfa_kernels.cuh"

namespace wfak {

struct RenderParams {
    const PairDesc *pairs; const Result *results; const uint64_t *ops_pool; const uint64_t *ops_where;
    const uint8_t *raw;
    uint32_t n_pairs; int only_aligned;
    unsigned long long *cursors;           /* [0] cigar bytes, [1] text bytes (three lines per pair) */
    uint64_t *cigar_off, *text_off; uint32_t *cigar_len, *text_len;
    uint8_t *cigar, *text; uint64_t cigar_cap, text_cap;
};

__device__ __forceinline__ uint32_t dec_digits(uint32_t n)
{
    uint32_t d = 1;
    while (n >= 10u) { n /= 10u; d++; }
    return d;
}

/* ops of the pair the reference would render: [first, last] after the optional trim */
__device__ __forceinline__ void render_range(const RenderParams &R, uint32_t pair, const uint64_t *&ops, uint32_t &first, uint32_t &count)
{
    const Result r = R.results[pair];
    ops = R.ops_pool + R.ops_where[pair]; first = 0; count = 0;
    if (r.status != ST_OK) return;
    count = r.n_ops;
    if (R.only_aligned) {
        int a = -1, b = -1;
        for (uint32_t i = 0; i < r.n_ops; i++) if ((uint32_t)(ops[i] >> 32) == 'M') { a = (int)i; break; }
        for (int i = (int)r.n_ops - 1; i >= 0; i--) if ((uint32_t)(ops[i] >> 32) == 'M') { b = i; break; }
        if (a < 0) { count = 0; return; }
        first = (uint32_t)a; count = (uint32_t)(b - a + 1);
    }
}

/* pass 1: lengths, then one reservation per warp and buffer; one thread per pair */
__global__ void __launch_bounds__(256)
render_measure_kernel(const RenderParams R)
{
    const uint32_t pair = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t clen = 0, tlen = 0;
    if (pair < R.n_pairs) {
        const uint64_t *ops; uint32_t first, count;
        render_range(R, pair, ops, first, count);
        for (uint32_t i = 0; i < count; i++) {
            const uint32_t n = (uint32_t)(ops[first + i] & 0xffffffffu);
            clen += dec_digits(n) + 1u; tlen += n;
        }
    }
    unsigned long long ci = clen, ti = 3ull * tlen;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ci, d), b = __shfl_up_sync(0xffffffffu, ti, d);
        if (lane >= d) { ci += a; ti += b; }
    }
    unsigned long long cb = 0, tb = 0;
    if (lane == 31) { cb = atomicAdd(&R.cursors[0], ci); tb = atomicAdd(&R.cursors[1], ti); }
    cb = __shfl_sync(0xffffffffu, cb, 31); tb = __shfl_sync(0xffffffffu, tb, 31);
    if (pair < R.n_pairs) {
        R.cigar_off[pair] = cb + ci - clen; R.cigar_len[pair] = clen;
        R.text_off[pair] = tb + ti - 3ull * tlen; R.text_len[pair] = tlen;
    }
}

/* pass 2a: CIGAR bytes, one thread per pair (a pair's string is a few dozen bytes for short reads) */
__global__ void __launch_bounds__(256)
render_cigar_kernel(const RenderParams R)
{
    const uint32_t pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= R.n_pairs) return;
    const uint32_t clen = R.cigar_len[pair];
    const uint64_t off = R.cigar_off[pair];
    if (clen == 0 || off + clen > R.cigar_cap) return;
    const uint64_t *ops; uint32_t first, count;
    render_range(R, pair, ops, first, count);
    uint8_t *out = R.cigar + off;
    for (uint32_t i = 0; i < count; i++) {
        const uint64_t op = ops[first + i];
        uint32_t n = (uint32_t)(op & 0xffffffffu);
        const uint32_t d = dec_digits(n);
        for (int j = (int)d - 1; j >= 0; j--) { out[j] = (uint8_t)('0' + n % 10u); n /= 10u; }
        out[d] = (uint8_t)(op >> 32);
        out += d + 1;
    }
}

/* pass 2b: the three text lines, one warp per pair; the lanes share the characters of an op */
__global__ void __launch_bounds__(256)
render_text_kernel(const RenderParams R)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t pr = warp0; pr < R.n_pairs; pr += nwarps) {
        const uint32_t pair = (uint32_t)pr;
        const uint32_t L = R.text_len[pair];
        const uint64_t off = R.text_off[pair];
        if (L == 0 || off + 3ull * L > R.text_cap) continue;
        const uint64_t *ops; uint32_t first, count;
        render_range(R, pair, ops, first, count);
        const PairDesc pd = R.pairs[pair];
        const Result r = R.results[pair];
        const uint8_t *q = R.raw + pd.q_byte + (R.only_aligned ? (uint32_t)(r.qbegin - 1) : 0u);
        const uint8_t *t = R.raw + pd.t_byte + (R.only_aligned ? (uint32_t)(r.tbegin - 1) : 0u);
        uint8_t *Q = R.text + off, *A = Q + L, *T = A + L;
        uint32_t v = 0, h = 0, pos = 0;
        for (uint32_t i = 0; i < count; i++) {
            const uint64_t op = ops[first + i];
            const uint32_t n = (uint32_t)(op & 0xffffffffu), o = (uint32_t)(op >> 32);
            const bool useq = o != 'I', uset = o == 'M' || o == 'X' || o == 'I';
            const uint8_t a = o == 'M' ? (uint8_t)'|' : (uint8_t)' ';
            for (uint32_t j = lane; j < n; j += 32) {
                Q[pos + j] = useq ? q[v + j] : (uint8_t)'-';
                A[pos + j] = a;
                T[pos + j] = uset ? t[h + j] : (uint8_t)'-';
            }
            pos += n; if (useq) v += n; if (uset) h += n;
        }
    }
}

} /* namespace wfak */
