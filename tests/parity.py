"""Shared helpers: run a batch through the GPU C ABI and through the oracle, compare bit-exactly."""
import numpy as np

import oracle_lib
from wfa_b200 import api

FIELDS = ["status", "score", "tbegin", "tend", "qbegin", "qend", "align_len", "matches", "gaps", "gap_regions", "n_ops"]


def make_aligner(mismatch=4, gap_open=6, gap_ext=2, global_alignment=True, adaptive=None, **kw):
    a = api.New(api.Penalties(mismatch, gap_open, gap_ext), api.Options(global_alignment), **kw)
    if adaptive is not None:
        a.AdaptiveReduction(api.AdaptiveReductionOption(adaptive[0], adaptive[1], 1))
    return a


def oracle_batch(batch, threads=8, **cfgkw):
    cfg = oracle_lib.make_config(**cfgkw)
    return oracle_lib.align_batch(cfg, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, threads=threads)


def assert_same(batch, gpu, ref, what=""):
    """gpu/ref = (results, ops, ops_off[, counters]).  TEnd/QEnd are only defined
    when the alignment has a match run (reference leaves stale values otherwise)."""
    gr, gops, goff = gpu[0], gpu[1], gpu[2]
    rr, rops, roff = ref[0], ref[1], ref[2]
    assert len(gr) == len(rr)
    for f in FIELDS:
        bad = np.nonzero(gr[f] != rr[f])[0]
        if len(bad):
            i = int(bad[0])
            q, t = batch.pair(i)
            gc = oracle_lib.ops_to_cigar(gops[int(goff[i]):int(goff[i]) + int(gr["n_ops"][i])]) if gr["status"][i] == 0 else "-"
            rc = oracle_lib.ops_to_cigar(rops[int(roff[i]):int(roff[i]) + int(rr["n_ops"][i])])
            raise AssertionError("%s: %d/%d pairs differ in %s; first pair %d (n=%d m=%d)\n gpu %s %s\n ref %s %s\n q=%r\n t=%r" % (
                what, len(bad), len(gr), f, i, len(q), len(t), {k: int(gr[k][i]) for k in FIELDS}, gc,
                {k: int(rr[k][i]) for k in FIELDS}, rc, q[:200], t[:200]))
    # ops: compare pair by pair (the buffer order of pairs is unspecified for chunked batches)
    gi, ri = ops_in_index_order(gr, gops, goff), ops_in_index_order(rr, rops, roff)
    if not np.array_equal(gi, ri):
        d = int(np.nonzero(gi != ri)[0][0])
        starts = np.concatenate([[0], np.cumsum(np.where(rr["status"] == 0, rr["n_ops"], 0).astype(np.int64))])
        i = int(np.searchsorted(starts, d, side="right") - 1)
        raise AssertionError("%s: ops differ first at word %d (pair %d)\n gpu %s\n ref %s" % (
            what, d, i, oracle_lib.ops_to_cigar(gops[int(goff[i]):int(goff[i]) + int(gr['n_ops'][i])]),
            oracle_lib.ops_to_cigar(rops[int(roff[i]):int(roff[i]) + int(rr['n_ops'][i])])))


def ops_in_index_order(res, ops, off):
    """Concatenate every pair's ops slice in pair-index order."""
    cnt = np.where(res["status"] == 0, res["n_ops"], 0).astype(np.int64)
    total = int(cnt.sum())
    if total == 0:
        return np.zeros(0, np.uint64)
    excl = np.cumsum(cnt) - cnt
    idx = np.repeat(off.astype(np.int64) - excl, cnt) + np.arange(total, dtype=np.int64)
    return np.asarray(ops)[idx]


def check(batch, what="", threads=8, gpu_kw=None, **cfgkw):
    a = make_aligner(**cfgkw, **(gpu_kw or {}))
    try:
        gpu = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
        stats = a.stats()
    finally:
        a.close()
    ref = oracle_batch(batch, threads=threads, **cfgkw)
    assert_same(batch, gpu, ref, what)
    return gpu, ref, stats


def replay_alignments(batch, res, ops, off, penalties=(4, 6, 2)):
    """Size-independent check of global alignments: every op list, replayed on its two
    sequences, consumes both completely ('I' consumes target, 'D' query -- wfa_cigar.go:312-328),
    has equal bases under 'M' and different ones under 'X', and costs what the result says: the
    score lies between the cost with every merged gap run opened once and the cost with every gap
    base opened separately (merged runs hide how often the path re-opened a gap)."""
    x, o, e = penalties
    for i in range(len(res)):
        assert res["status"][i] == 0, i
        q = batch.seq_bytes[int(batch.q_off[i]):int(batch.q_off[i]) + int(batch.q_len[i])]
        t = batch.seq_bytes[int(batch.t_off[i]):int(batch.t_off[i]) + int(batch.t_len[i])]
        w = np.asarray(ops[int(off[i]):int(off[i]) + int(res["n_ops"][i])])
        op, cnt = (w >> np.uint64(32)).astype(np.int64), (w & np.uint64(0xffffffff)).astype(np.int64)
        isM, isX, isI, isD = op == ord("M"), op == ord("X"), op == ord("I"), op == ord("D")
        assert (isM | isX | isI | isD).all(), i
        dq, dt = np.where(isM | isX | isD, cnt, 0), np.where(isM | isX | isI, cnt, 0)
        assert dq.sum() == len(q) and dt.sum() == len(t), i
        q0, t0 = np.cumsum(dq) - dq, np.cumsum(dt) - dt
        mx = isM | isX
        tot = int(cnt[mx].sum())
        base = np.repeat(np.cumsum(cnt[mx]) - cnt[mx], cnt[mx])
        within = np.arange(tot) - base
        qi, ti = np.repeat(q0[mx], cnt[mx]) + within, np.repeat(t0[mx], cnt[mx]) + within
        eq = q[qi] == t[ti]
        assert np.array_equal(eq, np.repeat(isM[mx], cnt[mx])), i
        gaps = cnt[isI | isD]
        lo = x * int(cnt[isX].sum()) + o * len(gaps) + e * int(gaps.sum())
        hi = x * int(cnt[isX].sum()) + (o + e) * int(gaps.sum())
        assert lo <= int(res["score"][i]) <= hi, (i, lo, int(res["score"][i]), hi)
        assert int(res["matches"][i]) <= int(cnt[isM].sum())
