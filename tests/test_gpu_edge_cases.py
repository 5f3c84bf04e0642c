"""More GPU parity: ragged / degenerate inputs, both worker classes in one batch, context reuse,
capacity errors, and size-independent properties at BASELINE sizes."""
import os
import random

import numpy as np
import pytest

import oracle_lib
import parity
from wfa_b200 import api, datagen

pytestmark = pytest.mark.gpu


def _rand(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def test_ragged_degenerate_batch(built_lib):
    rng = random.Random(99)
    pairs = []
    for L in (1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 500, 1500):
        q = _rand(rng, L)
        pairs += [(q, q), (q, q[::-1]), (q, q + _rand(rng, 5)), (_rand(rng, 3) + q, q), (q, _rand(rng, max(1, L // 2))),
                  (b"A" * L, b"A" * (L + 3)), (b"A" * L, b"C" * L), (b"AC" * L, b"CA" * L), (q, b"G"), (b"T", q)]
    batch = datagen.Batch.from_pairs(pairs)
    for glob in (True, False):
        for ad in (None, (10, 50), (2, 3)):
            parity.check(batch, what="ragged glob=%s ad=%s" % (glob, ad), global_alignment=glob, adaptive=ad)


def test_mixed_classes_one_batch(built_lib):
    """Short pairs (WARP worker) and pairs whose wavefront outgrows the widest ring (CTA worker,
    through the class decision and through the ring-overflow hand-over) in one call."""
    rng = random.Random(5)
    pairs = [(_rand(rng, 120), _rand(rng, 130)) for _ in range(200)]            # unrelated: wide wavefronts
    big = datagen.generate(6, 6000, 0.12, config=3)
    pairs += [big.pair(i) for i in range(len(big))]                              # 6 kbp, 12 %: CTA class without heuristic
    q = _rand(rng, 3000)
    pairs += [(q, _rand(rng, 3000))]                                             # unrelated 3 kbp: very wide
    batch = datagen.Batch.from_pairs(pairs)
    gpu, ref, stats = parity.check(batch, what="mixed classes")
    assert stats["pairs_cta"] + stats["pairs_wide"] > 0 and stats["pairs_warp"] > 0     # (the wide ones: WIDE worker, or CTA worker for other penalty shapes)
    parity.check(batch, what="mixed classes adaptive", adaptive=(10, 50))
    parity.check(datagen.Batch.from_pairs(pairs[:150] + pairs[200:203]), what="mixed semi", global_alignment=False)


def test_non_acgt_in_one_sequence_only(built_lib):
    rng = random.Random(8)
    pairs = []
    for _ in range(100):
        q = _rand(rng, 80)
        t = bytearray(q); t[rng.randrange(80)] = ord(rng.choice("Nacgt-*"))
        pairs += [(q, bytes(t)), (bytes(t), q), (q.lower(), q), (q, q)]
    batch = datagen.Batch.from_pairs(pairs)
    gpu, ref, stats = parity.check(batch, what="mixed alphabets")
    assert 0 < stats["pairs_8bit"] < len(pairs)
    parity.check(batch, what="mixed alphabets semi adaptive", global_alignment=False, adaptive=(3, 5))


def test_context_reuse_and_reconfigure(built_lib):
    a = parity.make_aligner()
    try:
        for count, L in ((10, 50), (3000, 100), (1, 5000), (500, 300)):
            b = datagen.generate(count, L, 0.08, config=2, first=count)
            gpu = a.align_arrays(b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len, copy=True)
            parity.assert_same(b, gpu, parity.oracle_batch(b), "reuse %d x %d" % (count, L))
        a.AdaptiveReduction(api.AdaptiveReductionOption(5, 10, 1))               # same ctx, new heuristic
        b = datagen.generate(300, 400, 0.1, config=3)
        gpu = a.align_arrays(b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len, copy=True)
        parity.assert_same(b, gpu, parity.oracle_batch(b, adaptive=(5, 10)), "after AdaptiveReduction")
    finally:
        a.close()


def test_ops_capacity_error_then_retry(built_lib):
    import ctypes as C
    b = datagen.generate(200, 200, 0.1, config=2)
    a = parity.make_aligner()
    try:
        L = a._L
        res = np.zeros(len(b), api.RESULT_DTYPE); off = np.zeros(len(b), np.uint64); ops = np.zeros(8, np.uint64)
        rc = L.wfacuda_align_batch(a._ctx, len(b), b.seq_bytes.ctypes.data, b.q_off.ctypes.data, b.q_len.ctypes.data,
                                   b.t_off.ctypes.data, b.t_len.ctypes.data, res.ctypes.data, ops.ctypes.data, 8, off.ctypes.data)
        assert rc == -4 and b"ops buffer" in L.wfacuda_last_error(a._ctx)
        need = int(L.wfacuda_last_ops_total(a._ctx))
        assert need == int(res["n_ops"].sum()) and (res["status"] == 0).all()     # results are valid already
        ops = np.zeros(need, np.uint64)
        rc = L.wfacuda_align_batch(a._ctx, len(b), b.seq_bytes.ctypes.data, b.q_off.ctypes.data, b.q_len.ctypes.data,
                                   b.t_off.ctypes.data, b.t_len.ctypes.data, res.ctypes.data, ops.ctypes.data, need, off.ctypes.data)
        assert rc == 0
        parity.assert_same(b, (res, ops, off), parity.oracle_batch(b), "after capacity retry")
        # results only (no CIGARs wanted)
        rc = L.wfacuda_align_batch(a._ctx, len(b), b.seq_bytes.ctypes.data, b.q_off.ctypes.data, b.q_len.ctypes.data,
                                   b.t_off.ctypes.data, b.t_len.ctypes.data, res.ctypes.data, None, 0, None)
        assert rc == 0 and np.array_equal(res["score"], parity.oracle_batch(b)[0]["score"])
    finally:
        a.close()


def test_tiny_arena_budget_requeues_then_reports_resources(built_lib):
    """A 1 MB arena cannot hold a 3 kbp / 15 % pair without heuristic: short pairs still align,
    the long one comes back with status 3 (WFACUDA_ERR_RESOURCES), nothing falls back to a CPU."""
    rng = random.Random(4)
    big = datagen.generate(1, 3000, 0.15, config=3)
    pairs = [(_rand(rng, 60), _rand(rng, 60)) for _ in range(50)] + [big.pair(0)]
    batch = datagen.Batch.from_pairs(pairs)
    a = parity.make_aligner(arena_budget_bytes=1 << 20)
    try:
        res, ops, off = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, copy=True)
    finally:
        a.close()
    ref = parity.oracle_batch(batch)
    assert res["status"][-1] == 3 and (res["status"][:-1] == 0).all()
    assert np.array_equal(res["score"][:-1], ref[0]["score"][:-1])


def test_lane_class_in_rounds_when_arena_is_small(built_lib):
    """The LANE class keeps one arena slot per group of 32 pairs until its finish kernel has
    run; a budget that holds only a few dozen slots makes it work through the batch in rounds
    (forward + finish kernel per round), with the same results."""
    batch = datagen.generate(6000, 150, 0.05, config=2)
    gpu, ref, stats = parity.check(batch, what="lane rounds", gpu_kw=dict(arena_budget_bytes=4 << 20))
    assert stats["pairs_lane"] > 5900 and stats["align_launches"] >= 3, stats
    assert stats["cells"] == ref[3]["cells"]
    full = parity.check(batch, what="lane one round")[2]
    assert full["align_launches"] < stats["align_launches"]


def _cigar_properties(batch, res, ops, off, penalties=(4, 6, 2)):
    """Size-independent checks: every CIGAR consumes exactly its query and target, and in
    global mode its gap-affine cost is the reported score."""
    ok = res["status"] == 0
    assert ok.all()
    o = api.ops_in_index_order(res, ops, off)
    code = (o >> np.uint64(32)).astype(np.uint8); cnt = (o & np.uint64(0xFFFFFFFF)).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(res["n_ops"].astype(np.int64))])[:-1]
    def per_pair(v):
        return np.add.reduceat(v, starts)
    isM, isX, isI, isD, isH = (code == ord(c) for c in "MXIDH")
    q_used = per_pair(np.where(isM | isX | isD | isH, cnt, 0)); t_used = per_pair(np.where(isM | isX | isI, cnt, 0))
    assert np.array_equal(q_used, batch.q_len.astype(np.int64)) and np.array_equal(t_used, batch.t_len.astype(np.int64))
    x, go, ge = penalties
    cost = per_pair(np.where(isX, cnt * x, 0) + np.where(isI | isD, go + ge * cnt, 0))
    return cost


def test_full_size_config2_bit_exact(built_lib):
    """BASELINE config 2 at its full size (1M pairs) against the oracle, bit-exact, through the
    chunked pipeline path of wfacuda_align_batch."""
    name = "cfg2_150bp_e5_global"
    batch = datagen.generate_config(name)
    assert len(batch) == 1_000_000
    gpu, ref, stats = parity.check(batch, what=name + " full", threads=32)
    cost = _cigar_properties(batch, *gpu)
    assert np.array_equal(cost, gpu[0]["score"].astype(np.int64))
    assert stats["cells"] == ref[3]["cells"] and stats["ops"] == ref[3]["ops"]


def test_config3_properties_at_scale(built_lib):
    """BASELINE config 3 (1 kbp, 10 %, wf-adaptive): 200k pairs on the GPU; CIGAR properties on
    all of them, bit-exact parity on a 10k prefix (the oracle needs minutes for all)."""
    name = "cfg3_1kbp_e10_global_adaptive"
    batch = datagen.generate_config(name, 200_000)
    a = parity.make_aligner(adaptive=(10, 50))
    try:
        gpu = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, copy=True)
    finally:
        a.close()
    cost = _cigar_properties(batch, *gpu)
    assert np.array_equal(cost, gpu[0]["score"].astype(np.int64))
    sub = batch.slice(0, 10_000)
    ref = parity.oracle_batch(sub, threads=32, adaptive=(10, 50))
    res, ops, off = gpu
    sub_gpu = (res[:10_000], api.ops_in_index_order(res[:10_000], ops, off[:10_000]),
               np.concatenate([[0], np.cumsum(res["n_ops"][:10_000].astype(np.uint64))])[:-1].astype(np.uint64))
    parity.assert_same(sub, sub_gpu, ref, name + " prefix")


def test_page_locked_buffers_and_ops_placement(built_lib):
    """wfacuda_host_alloc: inputs in page-locked memory take the direct-DMA path (chunked pipeline
    included) and give the same results as ordinary arrays; ops_off regions are disjoint, inside
    the used prefix of the buffer, and hold each pair's ops whatever order pairs completed in."""
    batch = datagen.generate(150_000, 150, 0.05, config=2)                  # > 2 chunks: pipelined path
    a = parity.make_aligner()
    try:
        host = [api.pinned_copy(x) for x in (batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)]
        r1, o1, off1 = a.align_arrays(*host, copy=True)
        st = a.stats()
        r2, o2, off2 = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, copy=True)
    finally:
        a.close()
    assert st["pairs_lane"] > 0 and st["h2d_bytes"] >= batch.seq_bytes.nbytes
    for f in parity.FIELDS:
        assert np.array_equal(r1[f], r2[f]), f
    assert np.array_equal(parity.ops_in_index_order(r1, o1, off1), parity.ops_in_index_order(r2, o2, off2))
    n_ops = r1["n_ops"].astype(np.int64)
    order = np.argsort(off1, kind="stable")
    ends = off1[order].astype(np.int64) + n_ops[order]
    assert (ends[:-1] <= off1[order][1:].astype(np.int64)).all() and ends.max() <= len(o1)
    sample = batch_slice(batch, 0, 3000)
    ref = parity.oracle_batch(sample)
    parity.assert_same(sample, (r1[:3000], o1, off1[:3000]), ref, "pinned e2e vs oracle")


def test_pipeline_graded_chunks_wire_descriptors_and_invalid_pairs(built_lib):
    """A page-locked batch large enough for the graded chunk sizes of the pipeline (small first and
    last chunks), whose workers send 20-byte wire descriptors: with empty sequences and non-ACGT
    bytes sprinkled in, results equal the single-ctx path on ordinary arrays and the oracle."""
    batch = datagen.generate(340_000, 150, 0.05, config=2)
    q_len = batch.q_len.copy()
    seq = batch.seq_bytes.copy()
    empties = [0, 5, 6655, 6656, 6657, 46_000, 170_001, 339_999]              # incl. the first chunk boundaries
    for i in empties:
        q_len[i] = 0
    for i in (1, 6660, 200_000, 339_998):
        seq[int(batch.t_off[i]) + 3] = ord("N")
    a = parity.make_aligner()
    try:
        host = [api.pinned_copy(x) for x in (seq, batch.q_off, q_len, batch.t_off, batch.t_len)]
        r1, o1, off1 = a.align_arrays(*host, copy=True)
        st = a.stats()
        os.environ["WFACUDA_NO_PIPELINE"] = "1"
        try:
            r2, o2, off2 = a.align_arrays(seq, batch.q_off, q_len, batch.t_off, batch.t_len, copy=True)
        finally:
            del os.environ["WFACUDA_NO_PIPELINE"]
    finally:
        a.close()
    assert st["pairs_8bit"] == 4 and (r1["status"][empties] == 1).all() and int((r1["status"] != 0).sum()) == len(empties)
    for f in parity.FIELDS:
        assert np.array_equal(r1[f], r2[f]), f
    assert np.array_equal(parity.ops_in_index_order(r1, o1, off1), parity.ops_in_index_order(r2, o2, off2))
    mod = datagen.Batch(seq, batch.q_off, q_len, batch.t_off, batch.t_len)
    for lo, hi in ((0, 2000), (6000, 8000), (338_000, 340_000)):
        sub = mod.slice(lo, hi)
        ref = parity.oracle_batch(sub)
        parity.assert_same(sub, (r1[lo:hi], o1, off1[lo:hi]), ref, "pipeline pairs %d..%d vs oracle" % (lo, hi))


def batch_slice(batch, a, b):
    pairs = [batch.pair(i) for i in range(a, b)]
    return datagen.Batch.from_pairs(pairs)


def test_scattered_pools_are_gathered(built_lib):
    """Pools where a batch's sequences are NOT one dense range: all queries then all targets, and
    targets that are windows into one shared reference (65 536 pairs each, so the call is cut
    into pipeline chunks).  Every chunk gathers its own sequences instead of uploading the range
    between its first and last byte -- which would be nearly the whole pool per chunk: the
    host-to-device bytes must stay near the sum of the sequence lengths."""
    import parity
    rng = np.random.default_rng(17)
    n, L = 65_536, 150
    # (a) queries first, then targets
    q = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(n, L))
    t = q.copy()
    mut = rng.random((n, L)) < 0.04
    t[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(mut.sum()))
    pool = np.concatenate([q.reshape(-1), t.reshape(-1), np.zeros(64, np.uint8)])
    q_off = (np.arange(n, dtype=np.uint64) * L); t_off = q_off + np.uint64(n * L)
    lens = np.full(n, L, np.uint32)
    batch_a = datagen.Batch(pool, q_off, lens, t_off, lens)
    # (b) reads against windows of one shared reference (windows overlap)
    ref_len = 1 << 20
    ref = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=ref_len)
    starts = rng.integers(0, ref_len - L - 8, n)
    reads = ref[starts[:, None] + np.arange(L)[None, :]].copy()
    mut = rng.random((n, L)) < 0.04
    reads[mut] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(mut.sum()))
    pool_b = np.concatenate([ref, reads.reshape(-1), np.zeros(64, np.uint8)])
    batch_b = datagen.Batch(pool_b, np.uint64(ref_len) + np.arange(n, dtype=np.uint64) * L, lens, starts.astype(np.uint64), lens)
    for batch, what in ((batch_a, "queries then targets"), (batch_b, "shared reference")):
        gpu, ref_, stats = parity.check(batch, what=what, threads=os.cpu_count() or 8)
        seq = int(batch.q_len.sum(dtype=np.uint64) + batch.t_len.sum(dtype=np.uint64))
        assert stats["h2d_bytes"] < 1.5 * seq + 64 * len(batch), (what, stats["h2d_bytes"], seq)
