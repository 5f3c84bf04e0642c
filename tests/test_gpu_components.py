"""The wavefront store of one pair read back from the GPU (wfacuda_align_components, SURVEY
section 8 f3): every raw cell of M, I, D against the oracle's, the README's M tables through the
Plot restatement, and the README's known-answer trace."""
import io
import json
import os
import random

import pytest

import oracle_lib
import parity
from wfa_b200 import api

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))


def _cell_text(c):
    return "." if c is None else "%s%2d" % (api.wfaArrows[c[1]], c[0])


@pytest.mark.parametrize("tab", G["m_tables"], ids=[t["source"] for t in G["m_tables"]])
def test_readme_m_tables_from_gpu_store(built_lib, tab):
    """Plot of the M component (README.md:104-113, 131-139) from the cells the GPU kept."""
    q, t = tab["q"].encode(), tab["t"].encode()
    stale = {tuple(c) for c in tab["stale_cells"]}
    a = parity.make_aligner(global_alignment=tab["global"], adaptive=(10, 50))
    try:
        res, comps = a.AlignComponents(q, t)
    finally:
        a.close()
    mat = comps.plot_matrix(q, t, "M", notChangeToMatch=False)
    checked = 0
    for v, row in enumerate(tab["rows"]):
        for h, want in enumerate(row):
            if (v + 1, h + 1) in stale:
                continue
            assert _cell_text(mat[v][h]) == want, "cell (%d,%d)" % (v + 1, h + 1)
            checked += 1
    assert checked == len(q) * len(t) - len(stale)
    buf = io.StringIO()
    comps.Plot(q, t, buf, "M")
    lines = buf.getvalue().split("\n")
    assert len(lines) == len(q) + 3 and lines[2].startswith("  1\t%s" % chr(q[0]))


def test_known_answer_trace_from_gpu_store(built_lib):
    """raw words of ACCATACTCG / AGGATGCTCG (README.md:101-124)."""
    a = parity.make_aligner()
    try:
        res, c = a.AlignComponents(b"ACCATACTCG", b"AGGATGCTCG")
    finally:
        a.close()
    assert res.CIGAR(False) == "1M2X2M1X4M" and res.Score == 12
    assert c.M.GetRaw(0, 0) == 1 << 3 | 6 and c.M.GetRaw(4, 0) == 2 << 3 | 5
    assert c.M.GetRaw(8, 0) == 5 << 3 | 5 and c.D.GetRaw(8, -1) == 1 << 3 | 3 and c.I.GetRaw(8, 1) == 2 << 3 | 1
    assert c.M.GetRaw(10, -2) == 1 << 3 | 4 and c.M.GetRaw(10, 2) == 3 << 3 | 2
    assert c.M.GetRaw(12, -1) == 2 << 3 | 5 and c.M.GetRaw(12, 1) == 3 << 3 | 5      # ties -> Mismatch
    assert c.M.GetRaw(12, 0) == 10 << 3 | 5 and c.M.GetRaw(12, 3) == 4 << 3 | 2
    for s in (1, 2, 3, 5, 6, 7, 9, 11):
        assert not c.M.HasScore(s)


def _compare_store(q, t, **kw):
    a = parity.make_aligner(**kw)
    try:
        res, comps = a.AlignComponents(q, t)
    finally:
        a.close()
    o = oracle_lib.Oracle(**kw)
    r = o.align(q, t)
    assert res.Score == r["score"] and res.CIGAR(False) == oracle_lib.ops_to_cigar(r["ops"])
    n_cells = 0
    for ci, comp in enumerate((comps.M, comps.I, comps.D)):
        for s in range(o.max_score() + 1 + 8):          # + the wavefront initComponents seeds at score x
            kr = o.krange(ci, s)
            ks = set(range(kr[0], kr[1] + 1)) if kr else set()
            if comp.HasScore(s):
                lo, hi = comp.KRange(s)
                ks |= set(range(lo, hi + 1))
            for k in ks:
                assert comp.GetRaw(s, k) == o.get_raw(ci, s, k), ("comp %d s %d k %d" % (ci, s, k), q, t, kw)
                n_cells += 1
        assert all(s <= o.max_score() + 8 for s in comp.W)
    o.close()
    return n_cells


def test_whole_store_matches_oracle(built_lib):
    """M, I, D cell by cell (offset and backtrace code) on random pairs: global with and without
    wf-adaptive, semi-global (which keeps every score up to the global corner), text bytes."""
    rng = random.Random(11)
    total = _compare_store(b"CGGCCCCTG", b"CGGCCCCTG", global_alignment=False, adaptive=(10, 50))   # ends at score 0: M[x] is seed-only
    total += _compare_store(b"ACGTACGT", b"TTACGTACGTCC", global_alignment=False)
    for it in range(24):
        L = rng.choice([3, 10, 40, 120, 400])
        alpha = b"ACGT" if it % 4 else b"ACGTN acgt"
        q = bytes(rng.choice(alpha) for _ in range(L))
        t = bytearray(q)
        for _ in range(max(1, L // 12)):
            j = rng.randrange(len(t) + 1)
            r = rng.random()
            if r < 0.4 and j < len(t):
                t[j] = rng.choice(alpha)
            elif r < 0.7:
                t.insert(j, rng.choice(alpha))
            elif j < len(t) and len(t) > 1:
                del t[j]
        t = bytes(t)
        kw = [dict(), dict(adaptive=(10, 50)), dict(global_alignment=False), dict(global_alignment=False, adaptive=(10, 50))][it % 4]
        if not kw.get("global_alignment", True):
            t = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 30))) + t + bytes(rng.choice(alpha) for _ in range(rng.randint(0, 30)))
        total += _compare_store(q, t, **kw)
    assert total > 10_000
