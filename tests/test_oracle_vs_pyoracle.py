"""Two independent restatements of the reference (C and Python) must agree:
this is what pins reduce, semi-global ties and exotic penalties, for which the
reference holds no golden vector."""
import random

import pytest

import oracle_lib
import pyoracle
from test_gpu_parity import CONFIGS, _random_pairs


def _tuple_c(r):
    return (r["score"], r["cigar"], r["tbegin"], r["tend"], r["qbegin"], r["qend"], r["align_len"], r["matches"], r["gaps"], r["gap_regions"])


def _tuple_py(r):
    return (r.Score, r.CIGAR(), r.TBegin, r.TEnd, r.QBegin, r.QEnd, r.AlignLen, r.Matches, r.Gaps, r.GapRegions)


@pytest.mark.parametrize("glob", [True, False])
def test_random_small(glob):
    pairs = _random_pairs(3, 24, maxlen=60)
    for g, ad, pen in CONFIGS:
        if g != glob:
            continue
        o = oracle_lib.Oracle(mismatch=pen[0], gap_open=pen[1], gap_ext=pen[2], global_alignment=g, adaptive=ad)
        p = pyoracle.Aligner(pen[0], pen[1], pen[2], g, ad)
        for q, t in pairs:
            assert _tuple_c(o.align(q, t)) == _tuple_py(p.Align(q, t)), (g, ad, pen, q, t)


def test_adaptive_triggers_on_longer_pairs():
    """Pairs long enough for reduce to trim (distance spread > MaxDistDiff)."""
    from wfa_b200 import datagen
    b = datagen.generate(6, 400, 0.12, config=3)
    for ad in ((10, 50), (10, 8), (2, 1)):
        o = oracle_lib.Oracle(adaptive=ad)
        o0 = oracle_lib.Oracle(adaptive=None)
        p = pyoracle.Aligner(adaptive=ad)
        trimmed = False
        for i in range(len(b)):
            q, t = b.pair(i)
            rc = o.align(q, t)
            assert _tuple_c(rc) == _tuple_py(p.Align(q, t))
            trimmed |= rc["counters"]["cells"] < o0.align(q, t)["counters"]["cells"]
        assert trimmed, "adaptive reduction never trimmed anything at %s" % (ad,)


def test_errors():
    o = oracle_lib.Oracle()
    assert o.align(b"", b"A")["status"] == 1 and o.align(b"A", b"")["status"] == 1
    assert o.align(b"C", b"C")["cigar"] == "1M" and o.align(b"CG", b"C")["status"] == 0
