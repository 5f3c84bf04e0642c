"""GPU parity: libwfacuda.so (through the C ABI) vs the CPU oracle, bit-exact."""
import os
import random

import numpy as np
import pytest

import parity
from wfa_b200 import api, datagen

pytestmark = pytest.mark.gpu

README = [  # (kwargs, q, t, cigar, score) -- reference README.md:18-27, 101-124, 231-240, 245-254
    (dict(global_alignment=False, adaptive=(10, 50)), b"Bioinformatics helps Biology", b"We learn bioinformatics to help biologists", "9I1X14M3I4M1D1M1X5M1X3I", 32),
    (dict(adaptive=(10, 50)), b"ACCATACTCG", b"AGGATGCTCG", "1M2X2M1X4M", 12),
    (dict(adaptive=(10, 50)), b"AGCTAGTGTCAATGGCTACTTTTCAGGTCCT", b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT", "1M1X3M1I5M2X8M3I1M1X9M", 36),
    (dict(adaptive=(10, 50)), b"ATTGGAAAATAGGATTGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTCGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTAGCTCGAAGCCCA",
     b"GATTGGAAAATAGGATGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTGCTCGAAGCCCA", "1X1I14M1D39M1D31M1D12M", 36),
]


def test_readme_goldens(built_lib):
    for kw, q, t, cigar, score in README:
        a = parity.make_aligner(**kw)
        r = a.Align(q, t)
        assert (r.CIGAR(False), r.Score) == (cigar, score)
        api.RecycleAligner(a)


def _mutate(rng, s, rate, alpha):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < rate / 3:
            out.append(rng.choice(alpha))
        elif r < 2 * rate / 3:
            pass
        elif r < rate:
            out.append(c); out.append(rng.choice(alpha))
        else:
            out.append(c)
    return bytes(out) or b"A"


def _random_pairs(seed, count, alpha=b"ACGT", maxlen=120):
    rng = random.Random(seed)
    pairs = []
    for it in range(count):
        al = alpha if it % 3 else alpha[:2]
        L = rng.choice([1, 2, 3, 5, 8, 13, 20, 40, 70, maxlen])
        q = bytes(rng.choice(al) for _ in range(L))
        mode = it % 4
        if mode == 0:
            t = _mutate(rng, q, 0.2, al)
        elif mode == 1:
            t = bytes(rng.choice(al) for _ in range(rng.randint(1, L + 10)))
        elif mode == 2:
            t = _mutate(rng, q, 0.05, al) + bytes(rng.choice(al) for _ in range(rng.randint(0, 20)))
        else:
            t = bytes(rng.choice(al) for _ in range(rng.randint(0, 15))) + _mutate(rng, q, 0.3, al)
        pairs.append((q, t))
    return pairs


CONFIGS = [(glob, ad, pen) for glob in (True, False) for ad in (None, (10, 50), (3, 5), (1, 2), (5, 0))
           for pen in ((4, 6, 2), (1, 0, 1), (3, 1, 2), (2, 3, 1), (5, 2, 3), (4, 4, 4), (7, 11, 3))]


@pytest.mark.parametrize("worker", ["auto", "warp", "cta"])
def test_random_small_all_configs(built_lib, worker):
    """auto = LANE worker for global/no-heuristic configs (32 pairs per warp in lockstep), WARP
    worker otherwise; warp = LANE switched off; cta = every pair on the CTA worker."""
    pairs = _random_pairs(7, 300)
    batch = datagen.Batch.from_pairs(pairs)
    flags = {"auto": 0, "warp": api.FLAG_NO_LANE, "cta": api.FLAG_FORCE_CTA}[worker]
    for glob, ad, pen in CONFIGS:
        gpu, ref, stats = parity.check(batch, what="glob=%s ad=%s pen=%s worker=%s" % (glob, ad, pen, worker), mismatch=pen[0],
                                       gap_open=pen[1], gap_ext=pen[2], global_alignment=glob, adaptive=ad, gpu_kw=dict(flags=flags))
        if worker == "auto" and glob and ad is None and pen in ((4, 6, 2), (1, 0, 1), (2, 3, 1)):
            assert stats["pairs_lane"] > 0, stats
        if worker != "auto":
            assert stats["pairs_lane"] == 0, stats
        if glob:
            assert stats["cells"] == ref[3]["cells"], (worker, glob, ad, pen, stats, ref[3])


def test_lane_worker_boundaries(built_lib):
    """LANE worker limits: lengths around 254 (byte offsets), wavefronts wider than its 64-column
    ring (handed to the WARP worker), groups that are not a multiple of 32, lanes that finish at
    very different scores, non-ACGT pairs inside a group (8-bit hand-over)."""
    rng = random.Random(31)
    rnd = lambda n, al=b"ACGT": bytes(rng.choice(al) for _ in range(n))
    pairs = []
    for L in (250, 253, 254, 255, 256, 300):
        q = rnd(L)
        pairs += [(q, q), (q, _mutate(rng, q, 0.03, b"ACGT")), (q[:L - 3], q), (q, q[2:]), (q, rnd(L))]
    for _ in range(40):                                           # unrelated: score > 64 columns can hold
        pairs.append((rnd(rng.randint(60, 200)), rnd(rng.randint(60, 200))))
    for _ in range(150):                                          # related, mixed error rates and lengths
        q = rnd(rng.randint(1, 254))
        pairs.append((q, _mutate(rng, q, rng.choice([0.0, 0.01, 0.05, 0.1, 0.2]), b"ACGT")))
    for _ in range(20):
        q = rnd(100)
        t = bytearray(q); t[rng.randrange(100)] = ord("N")
        pairs.append((q, bytes(t)))
    pairs += [(b"A", b"A"), (b"A", b"C"), (b"AC", b"A"), (b"A", b"ACGTACGT"), (b"ACGTACGTAC", b"A")]
    rng.shuffle(pairs)
    batch = datagen.Batch.from_pairs(pairs)
    for pen in ((4, 6, 2), (1, 0, 1), (3, 1, 2), (2, 3, 1), (5, 2, 3)):
        gpu, ref, stats = parity.check(batch, what="lane boundaries pen=%s" % (pen,), mismatch=pen[0], gap_open=pen[1], gap_ext=pen[2])
        assert stats["pairs_lane"] > 0 and stats["pairs_warp"] > 0 and stats["pairs_8bit"] > 0, stats
        assert stats["cells"] == ref[3]["cells"], (pen, stats, ref[3])
    parity.check(batch, what="lane boundaries, LANE off", gpu_kw=dict(flags=api.FLAG_NO_LANE))


def test_lane_stages_follow_a_changing_workload(built_lib):
    """The LANE class cuts its rows into stages at the quantiles of the PREVIOUS batch's final
    scores and sizes the later stages' slots from that histogram.  Batches whose score
    distribution jumps (low error -> high error -> short reads -> low error) on one aligner must
    stay exact: pairs that find a later stage full are re-queued."""
    a = parity.make_aligner()
    try:
        for it, (L, e) in enumerate(((150, 0.02), (150, 0.12), (60, 0.05), (200, 0.08), (150, 0.02))):
            batch = datagen.generate(6000, L, e, config=2, first=it * 6000)
            for rep in range(2):                                    # second run uses what the first learned
                gpu = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, copy=True)
                st = a.stats()
                ref = parity.oracle_batch(batch) if rep == 0 else ref
                parity.assert_same(batch, gpu, ref, "L=%d e=%.2f run %d" % (L, e, rep))
                assert st["pairs_lane"] > 0 and st["cells"] == ref[3]["cells"], (L, e, rep, st)
    finally:
        a.close()


def test_text_and_8bit_path(built_lib):
    pairs = _random_pairs(11, 200, alpha=b"abcdefgh -N") + [(b"Bioinformatics helps Biology", b"We learn bioinformatics to help biologists"),
                                                            (b"acgt", b"ACGT"), (b"ACGTN", b"ACGTN")]
    batch = datagen.Batch.from_pairs(pairs)
    for glob in (True, False):
        for ad in (None, (3, 5)):
            parity.check(batch, what="text glob=%s ad=%s" % (glob, ad), global_alignment=glob, adaptive=ad)
    # forcing the 8-bit path on ACGT input must not change anything
    batch = datagen.Batch.from_pairs(_random_pairs(12, 200))
    parity.check(batch, what="force8", gpu_kw=dict(flags=api.FLAG_FORCE_8BIT))


def test_semiglobal_literal_equals_early_stop(built_lib):
    batch = datagen.Batch.from_pairs(_random_pairs(13, 300))
    for ad in (None, (10, 50), (3, 5)):
        parity.check(batch, what="semi literal ad=%s" % (ad,), global_alignment=False, adaptive=ad,
                     gpu_kw=dict(flags=api.FLAG_SEMIGLOBAL_LITERAL))


def test_errors_and_empty(built_lib):
    a = parity.make_aligner()
    res, errs = a.AlignBatch([b"", b"ACGT", b"A"], [b"ACGT", b"", b"A"])
    assert errs[0] is api.ErrEmptySeq and errs[1] is api.ErrEmptySeq and errs[2] is None
    assert res[2].CIGAR() == "1M"
    with pytest.raises(api.WfaError):
        a.Align(b"", b"A")
    r, o, off = a.align_arrays(np.zeros(16, np.uint8), [], [], [], [])
    assert len(r) == 0
    api.RecycleAligner(a)
    with pytest.raises(api.WfaError):
        parity.make_aligner().AdaptiveReduction(api.AdaptiveReductionOption(0, 50, 1))


# Pair counts per config: the oracle run on the GPU box's host cores is the only real cost (a few
# seconds each for configs 2, 3 and 5, about a minute for the 16 semi-global pairs of config 4).
@pytest.mark.parametrize("name,count", [("cfg2_150bp_e5_global", 20000), ("cfg3_1kbp_e10_global_adaptive", 100000),
                                        ("cfg4_10kbp_in_12kbp_e5_semiglobal", 16), ("cfg5_100kbp_e15_global_adaptive", 32)])
def test_synthetic_configs(built_lib, name, count):
    c = datagen.CONFIGS[name]
    batch = datagen.generate_config(name, count)
    gpu, ref, stats = parity.check(batch, what=name, threads=os.cpu_count() or 8, global_alignment=c["global_alignment"], adaptive=c["adaptive"])
    # device work counter C must equal the oracle's (roofline numerator)
    if c["global_alignment"]:
        assert stats["cells"] == ref[3]["cells"], (stats, ref[3])
    if c["adaptive"]:
        assert stats["pairs_slim"] > 0.9 * count, stats           # configs 3 and 5 run on the SLIM worker
    if not c["global_alignment"]:
        assert stats["pairs_wide"] == count, stats                # config 4 runs on the WIDE worker (2-CTA clusters)


def test_config5_shard_shape(built_lib):
    """Config 5 as one GPU of eight sees it: 1 250 pairs of 100 kbp.  The oracle needs minutes for
    that many, so: (1) the SLIM worker (8-byte offset cells, codes re-derived by the
    backtrace, sequences through a moving shared-memory window) against the WARP worker (raw words with codes in a shared-memory ring) on all 1 250
    -- two independent forward passes and arena formats; (2) the oracle on a 40-pair sample of the
    same batch; (3) every alignment replayed on its sequences."""
    name = "cfg5_100kbp_e15_global_adaptive"
    batch = datagen.generate_config(name, 1250)
    res = {}
    for flags in (0, api.FLAG_NO_SLIM):
        a = parity.make_aligner(adaptive=(10, 50), flags=flags)
        try:
            res[flags] = a.align_arrays(batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len, copy=True)
            st = a.stats()
            assert (st["pairs_slim"] > 1000) == (flags == 0), st
        finally:
            a.close()
    parity.assert_same(batch, res[0], res[api.FLAG_NO_SLIM], "config 5 shard: REG vs WARP worker")
    idx = np.r_[0:16, 600:612, 1238:1250]
    sub = datagen.Batch(batch.seq_bytes, batch.q_off[idx], batch.q_len[idx], batch.t_off[idx], batch.t_len[idx])
    ref = parity.oracle_batch(sub, threads=os.cpu_count() or 8, adaptive=(10, 50))
    r, o, off = res[0]
    parity.assert_same(sub, (r[idx], o, off[idx]), ref, "config 5 shard: oracle sample")
    parity.replay_alignments(batch, r, o, off, penalties=(4, 6, 2))


def test_slim_worker_boundaries(built_lib):
    """SLIM worker (wfa_slim.cuh): rows that outgrow the ring (the wider instantiation, then the WARP
    worker), targets around the 1 022-base limit of the 32-bit cell word and the 2 048-base window, non-ACGT pairs (8-bit
    hand-over), pairs that end at score 0, sequences longer than the shared-memory window."""
    rng = random.Random(41)
    rnd = lambda n, al=b"ACGT": bytes(rng.choice(al) for _ in range(n))
    pairs = []
    for L in (255, 300, 700, 1000, 1020, 1022, 1023, 1024, 1030, 2040, 2047, 2048, 2049, 2100, 3000, 5000, 9000):
        q = rnd(L)
        pairs += [(q, q), (q, _mutate(rng, q, 0.02, b"ACGT")), (q, _mutate(rng, q, 0.12, b"ACGT")), (q[:L - 5], q), (q, q[7:])]
    for _ in range(60):
        q = rnd(rng.randint(255, 1500))
        pairs.append((q, _mutate(rng, q, rng.choice([0.0, 0.01, 0.05, 0.1, 0.2, 0.3]), b"ACGT")))
    for _ in range(10):                                           # unrelated: wide, high scores
        pairs.append((rnd(rng.randint(255, 600)), rnd(rng.randint(255, 600))))
    for _ in range(10):
        q = rnd(400)
        t = bytearray(q); t[rng.randrange(400)] = ord("N")
        pairs.append((q, bytes(t)))
    rng.shuffle(pairs)
    batch = datagen.Batch.from_pairs(pairs)
    for ad in ((10, 50), (3, 5), (1, 2), (10, 200), None):
        for pen in ((4, 6, 2), (2, 3, 1), (8, 12, 4)):
            gpu, ref, stats = parity.check(batch, what="slim boundaries ad=%s pen=%s" % (ad, pen), adaptive=ad, mismatch=pen[0], gap_open=pen[1], gap_ext=pen[2])
            assert stats["cells"] == ref[3]["cells"], (ad, pen, stats, ref[3])
            if ad in ((10, 50), (3, 5), (1, 2)):
                assert stats["pairs_slim"] > 0 and stats["pairs_8bit"] > 0, stats
    # every rung of the ladder forced: what does not fit goes straight to the WARP worker
    try:
        for p_ in ("4", "8"):
            os.environ["WFACUDA_SLIM_P"] = p_
            gpu, ref, stats = parity.check(batch, what="slim boundaries passes=%s" % p_, adaptive=(10, 50))
            assert stats["pairs_slim"] > 0 and stats["cells"] == ref[3]["cells"], (p_, stats)
    finally:
        os.environ.pop("WFACUDA_SLIM_P", None)
    parity.check(batch, what="slim boundaries, SLIM off", adaptive=(10, 50), gpu_kw=dict(flags=api.FLAG_NO_SLIM))


def test_wide_worker(built_lib):
    """WIDE worker (wfa_wide.cuh): one thread-block cluster per pair, the live rows as 16-bit offsets in
    the (distributed) shared memory of the cluster, halo cells pushed to the neighbouring CTA, the row's
    reductions through per-CTA mailboxes and one cluster barrier per score; backtraces in their own kernel.
    Cluster sizes 1 .. 8 forced on small pairs (segments of 64 diagonals upwards, so that rows cross CTA
    boundaries and halos matter), global and semi-global, every penalty set of the default shape; pairs with
    a non-ACGT byte (8-bit hand-over to the CTA worker); and the worker switched off (CTA worker) as a control."""
    rng = random.Random(53)
    rnd = lambda n, al=b"ACGT": bytes(rng.choice(al) for _ in range(n))
    pairs = _random_pairs(51, 120, maxlen=300)
    for L in (63, 64, 65, 127, 128, 129, 255, 256, 257, 500, 700):
        q = rnd(L)
        pairs += [(q, q), (q, _mutate(rng, q, 0.05, b"ACGT")), (q, _mutate(rng, q, 0.25, b"ACGT")), (q[:L - 3], q), (q, q[5:]),
                  (rnd(L // 2 + 1) + q, q), (q, rnd(L // 3 + 1) + _mutate(rng, q, 0.1, b"ACGT") + rnd(L // 3 + 1)), (q, rnd(L))]
    for _ in range(6):
        q = rnd(200)
        t = bytearray(q); t[rng.randrange(200)] = ord("N")
        pairs.append((q, bytes(t)))
    pairs += [(b"A", b"A"), (b"A", b"C"), (b"AC", b"A"), (b"A", b"ACGTACGT"), (b"ACGTACGTAC", b"A")]
    rng.shuffle(pairs)
    batch = datagen.Batch.from_pairs(pairs)
    try:
        for cl in ("1", "2", "4", "8", None):
            if cl is None:
                os.environ.pop("WFACUDA_WIDE_CLUSTER", None)
            else:
                os.environ["WFACUDA_WIDE_CLUSTER"] = cl
            for glob in (True, False):
                for pen in ((4, 6, 2), (2, 3, 1), (8, 12, 4)):
                    gpu, ref, stats = parity.check(batch, what="wide cluster=%s glob=%s pen=%s" % (cl, glob, pen), global_alignment=glob,
                                                   mismatch=pen[0], gap_open=pen[1], gap_ext=pen[2], gpu_kw=dict(flags=api.FLAG_FORCE_CTA))
                    assert stats["pairs_wide"] > 0.9 * len(batch) and stats["pairs_8bit"] > 0, stats
                    if glob:
                        assert stats["cells"] == ref[3]["cells"], (cl, pen, stats, ref[3])
    finally:
        os.environ.pop("WFACUDA_WIDE_CLUSTER", None)
    for glob in (True, False):
        gpu, ref, stats = parity.check(batch, what="wide off glob=%s" % glob, global_alignment=glob, gpu_kw=dict(flags=api.FLAG_FORCE_CTA | api.FLAG_NO_WIDE))
        assert stats["pairs_wide"] == 0 and stats["pairs_cta"] > 0, stats


def test_seqs_txt_config1(built_lib):
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))
    pairs = [(p["q"].encode(), p["t"].encode()) for p in G["seqs_txt"]]
    assert len(pairs) == 2
    batch = datagen.Batch.from_pairs(pairs)
    parity.check(batch, what="seqs.txt", adaptive=(10, 50))      # CLI defaults (wfa-go.go:96-106)
    a = parity.make_aligner(adaptive=(10, 50))
    r = a.Align(*pairs[0])
    assert r.CIGAR() == "1X1I14M1D39M1D31M1D12M" and r.Score == 36
    assert (r.QBegin, r.QEnd, r.TBegin, r.TEnd, r.AlignLen, r.Matches, r.Gaps, r.GapRegions) == (2, 100, 3, 98, 99, 96, 3, 3)


def test_align_batch_multi(built_lib):
    """wfacuda_align_batch_multi: length-binned LPT shards, every device running the chunked
    pipeline on its shard, results and ops back at the caller's indices, identical to the oracle.
    One ctx per visible device (two ctxs on device 0 when only one GPU is visible): a mixed-length
    batch (non-contiguous shards, gather / scatter) and a uniform one big enough to be pipelined
    (contiguous shards, several chunks per device)."""
    n_dev = min(api.device_count(), 8)
    devs = list(range(n_dev)) if n_dev >= 2 else [0, 0]
    mixed = datagen.Batch.from_pairs(_random_pairs(21, 500, maxlen=300) + _random_pairs(22, 40, maxlen=3000))
    uniform = datagen.generate(150_000, 150, 0.05, config=2)
    for batch, ad, what in ((mixed, (10, 50), "multi mixed"), (mixed, None, "multi mixed no heuristic"), (uniform, None, "multi uniform")):
        algns = [parity.make_aligner(adaptive=ad, device=d) for d in devs]
        try:
            gpu = algns[0].AlignBatchMulti(algns[1:], batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
            pairs_per_dev = [a.stats()["pairs"] for a in algns]
        finally:
            for a in algns:
                a.close()
        assert sum(pairs_per_dev) == len(batch) and min(pairs_per_dev) > 0, pairs_per_dev
        ref = parity.oracle_batch(batch, threads=os.cpu_count() or 8, adaptive=ad)
        parity.assert_same(batch, gpu, ref, what + " on devices %s" % (devs,))
