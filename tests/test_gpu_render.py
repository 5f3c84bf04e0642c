"""GPU rendering of CIGAR strings and alignment text (wfacuda_batch_render)
against the lines the reference prints in its README and against the Python restatement of
wfa_cigar.go:217-333 applied to the oracle's ops."""
import os
import random
import sys

import numpy as np
import pytest

import oracle_lib
import parity
from wfa_b200 import api, datagen

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle  # noqa: E402

pytestmark = pytest.mark.gpu

README_TEXT = [     # (kwargs, q, t, query line, marks, target line, cigar): README.md:18-27, 231-240, 245-254
    (dict(global_alignment=False, adaptive=(10, 50)), b"Bioinformatics helps Biology", b"We learn bioinformatics to help biologists",
     b"---------Bioinformatics ---helps Biology---", b"          ||||||||||||||   |||| | |||||    ", b"We learn bioinformatics to help- biologists",
     "9I1X14M3I4M1D1M1X5M1X3I"),
    (dict(adaptive=(10, 50)), b"AGCTAGTGTCAATGGCTACTTTTCAGGTCCT", b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT",
     b"AGCTA-GTGTCAATGGCTACT---TTTCAGGTCCT", b"| ||| |||||  ||||||||   | |||||||||", b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT",
     "1M1X3M1I5M2X8M3I1M1X9M"),
    (dict(adaptive=(10, 50)), b"ATTGGAAAATAGGATTGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTCGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTAGCTCGAAGCCCA",
     b"GATTGGAAAATAGGATGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTGCTCGAAGCCCA",
     b"A-TTGGAAAATAGGATTGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTCGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTAGCTCGAAGCCCA",
     b"  |||||||||||||| ||||||||||||||||||||||||||||||||||||||| ||||||||||||||||||||||||||||||| ||||||||||||",
     b"GATTGGAAAATAGGAT-GGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTT-GTCGTCCTTACGTTTCCGGAAGGGAGTGGTT-GCTCGAAGCCCA",
     "1X1I14M1D39M1D31M1D12M"),
]


def _render(batch, only_aligned, **kw):
    a = parity.make_aligner(**kw)
    try:
        rb = api.ResidentBatch(a, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
        rb.run()
        res, ops, off = rb.download()
        out = rb.render(onlyAignedRegion=only_aligned, text=True)
        launches = a.stats()["kernel_launches"]
        rb.free()
    finally:
        a.close()
    return res, ops, off, out, launches


def test_readme_text_blocks(built_lib):
    """The three lines and the CIGAR the reference's CLI prints (whole sequences, README.md)."""
    for kw, q, t, ql, al, tl, cigar in README_TEXT:
        batch = datagen.Batch.from_pairs([(q, t)])
        res, ops, off, out, launches = _render(batch, False, **kw)
        assert res["status"][0] == 0 and launches >= 3
        assert out.CIGAR(0) == cigar
        assert out.AlignmentText(0) == (ql, al.ljust(len(ql)), tl)


def _expected(batch, ref, i, only_aligned):
    """CIGAR and text of pair i from the oracle's result, by the Python restatement of wfa_cigar.go."""
    rr, rops, roff = ref[0], ref[1], ref[2]
    if rr["status"][i] != 0:
        return "", (b"", b"", b"")
    ops = [int(x) for x in rops[int(roff[i]):int(roff[i]) + int(rr["n_ops"][i])]]
    if only_aligned and not any(op >> 32 == ord("M") for op in ops):
        return "", (b"", b"", b"")          # the reference panics here (ops[-1:0]); the library returns empty strings
    r = pyoracle.Result()
    r.Ops = ops
    r.QBegin, r.QEnd, r.TBegin, r.TEnd = (int(rr[k][i]) for k in ("qbegin", "qend", "tbegin", "tend"))
    q, t = batch.pair(i)
    return r.CIGAR(only_aligned), r.AlignmentText(q, t, only_aligned)


@pytest.mark.parametrize("only_aligned", [False, True])
@pytest.mark.parametrize("name,kw,gen", [
    ("lane short global", dict(), lambda: datagen.generate(3000, 150, 0.05, config=2)),
    ("warp 1 kbp adaptive", dict(adaptive=(10, 50)), lambda: datagen.generate(300, 1000, 0.10, config=3)),
    ("semi-global", dict(global_alignment=False), lambda: datagen.generate(64, 300, 0.05, window=400, max_start=100, config=4)),
])
def test_render_matches_reference_formatting(built_lib, name, kw, gen, only_aligned):
    batch = gen()
    res, ops, off, out, _ = _render(batch, only_aligned, **kw)
    ref = parity.oracle_batch(batch, **kw)
    parity.assert_same(batch, (res, ops, off), ref, name)
    for i in range(len(batch)):
        cigar, text = _expected(batch, ref, i, only_aligned)
        assert out.CIGAR(i) == cigar, (name, i)
        assert out.AlignmentText(i) == text, (name, i)
    # strings are disjoint regions of the buffers
    order = np.argsort(out.cigar_off, kind="stable")
    ends = out.cigar_off[order].astype(np.int64) + out.cigar_len[order]
    assert (ends[:-1] <= out.cigar_off[order][1:].astype(np.int64)).all() and ends.max() <= len(out.cigar)


def test_render_ragged_text_and_invalid_pairs(built_lib):
    """Arbitrary bytes (8-bit path), degenerate lengths, empty sequences, alignments without a match."""
    rng = random.Random(7)
    pairs = [(b"", b"ACGT"), (b"A", b"C"), (b"AAAA", b"TTTTTTTT"), (b"hello world", b"help the world"), (b"ACGT", b"")]
    for _ in range(200):
        L = rng.choice([1, 2, 3, 7, 30, 90])
        q = bytes(rng.choice(b"ACGTN acgt") for _ in range(L))
        t = bytes(rng.choice(b"ACGTN acgt") for _ in range(rng.randint(1, L + 12)))
        pairs.append((q, t))
    batch = datagen.Batch.from_pairs(pairs)
    for kw in (dict(), dict(global_alignment=False)):
        for only_aligned in (False, True):
            res, ops, off, out, _ = _render(batch, only_aligned, **kw)
            ref = parity.oracle_batch(batch, **kw)
            parity.assert_same(batch, (res, ops, off), ref, "ragged")
            assert res["status"][0] == 1 and out.CIGAR(0) == "" and out.AlignmentText(0) == (b"", b"", b"")
            for i in range(len(batch)):
                cigar, text = _expected(batch, ref, i, only_aligned)
                assert out.CIGAR(i) == cigar and out.AlignmentText(i) == text, (kw, only_aligned, i, pairs[i])
