"""The oracle pinned against every golden vector the reference holds for the path
(SURVEY.md section 8c): five README alignments, two README M-component tables."""
import json
import os

import pytest

import oracle_lib
import pyoracle

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))
STAT_FIELDS = ["score", "qbegin", "qend", "tbegin", "tend", "align_len", "matches", "gaps", "gap_regions"]


@pytest.mark.parametrize("vec", G["alignments"], ids=[v["source"] for v in G["alignments"]])
def test_c_oracle_readme_alignment(vec):
    o = oracle_lib.Oracle(global_alignment=vec["global"], adaptive=tuple(vec["adaptive"]))
    r = o.align(vec["q"].encode(), vec["t"].encode())
    assert r["status"] == 0
    for f in STAT_FIELDS:
        assert r[f] == vec[f], f
    if vec.get("cigar_stale"):
        # README block printed by an older release (label "align-region", pre-0.4 API); the
        # current source places the insertion one base later in the CC run
        assert r["cigar"] == vec["cigar_current"]
    else:
        assert r["cigar"] == vec["cigar"]
    # with and without the CLI's default heuristic: these inputs are too short for it to trim
    o2 = oracle_lib.Oracle(global_alignment=vec["global"], adaptive=None)
    assert o2.align(vec["q"].encode(), vec["t"].encode())["cigar"] == r["cigar"]


@pytest.mark.parametrize("vec", G["alignments"], ids=[v["source"] for v in G["alignments"]])
def test_py_oracle_readme_alignment(vec):
    a = pyoracle.Aligner(global_alignment=vec["global"], adaptive=tuple(vec["adaptive"]))
    r = a.Align(vec["q"].encode(), vec["t"].encode())
    got = dict(score=r.Score, qbegin=r.QBegin, qend=r.QEnd, tbegin=r.TBegin, tend=r.TEnd, align_len=r.AlignLen,
               matches=r.Matches, gaps=r.Gaps, gap_regions=r.GapRegions)
    for f in STAT_FIELDS:
        assert got[f] == vec[f], f
    assert r.CIGAR() == (vec["cigar_current"] if vec.get("cigar_stale") else vec["cigar"])


def _render(cell):
    return "." if cell is None else "%s%2d" % (pyoracle.ARROWS[cell[1]], cell[0])


@pytest.mark.parametrize("tab", G["m_tables"], ids=[t["source"] for t in G["m_tables"]])
def test_m_table(tab):
    """Plot (wfa_component_plot.go:41-209) of the M component after Align, cell by cell.
    The C oracle's wavefronts are loaded into the Python Plot restatement, so both
    restatements are checked against the README table."""
    q, t = tab["q"].encode(), tab["t"].encode()
    stale = {tuple(c) for c in tab["stale_cells"]}
    a = pyoracle.Aligner(global_alignment=tab["global"], adaptive=(10, 50))
    a.Align(q, t)
    mat_py = a.plot_matrix(q, t, "M", notChangeToMatch=False)

    # same, but with M/I/D taken from the C oracle
    o = oracle_lib.Oracle(global_alignment=tab["global"], adaptive=(10, 50))
    o.align(q, t)
    b = pyoracle.Aligner(global_alignment=tab["global"], adaptive=(10, 50))
    b.M, b.I, b.D = pyoracle.Component(), pyoracle.Component(), pyoracle.Component()
    for ci, comp in enumerate((b.M, b.I, b.D)):
        for s in range(o.max_score() + 1):
            kr = o.krange(ci, s)
            if kr is None:
                continue
            wf = comp.W[s] = pyoracle.WaveFront()
            wf.Lo, wf.Hi = kr
            for k in range(kr[0], kr[1] + 1):
                raw = o.get_raw(ci, s, k)
                if raw:
                    wf.c[k] = raw
    mat_c = b.plot_matrix(q, t, "M", notChangeToMatch=False)
    assert mat_c == mat_py

    n_checked = 0
    for v, row in enumerate(tab["rows"]):
        for h, want in enumerate(row):
            if (v + 1, h + 1) in stale:
                continue
            assert _render(mat_py[v][h]) == want, "cell (%d,%d)" % (v + 1, h + 1)
            n_checked += 1
    assert n_checked == len(q) * len(t) - len(stale)


def test_known_answer_trace_global():
    """SURVEY.md section 8c known-answer wavefront trace for ACCATACTCG / AGGATGCTCG."""
    o = oracle_lib.Oracle()
    o.align(b"ACCATACTCG", b"AGGATGCTCG")
    M, I, D = 0, 1, 2
    assert o.get_raw(M, 0, 0) == 1 << 3 | 6
    assert o.get_raw(M, 4, 0) == 2 << 3 | 5
    assert o.get_raw(M, 8, 0) == 5 << 3 | 5 and o.get_raw(D, 8, -1) == 1 << 3 | 3 and o.get_raw(I, 8, 1) == 2 << 3 | 1
    assert o.get_raw(M, 10, -2) == 1 << 3 | 4 and o.get_raw(M, 10, 2) == 3 << 3 | 2
    assert o.get_raw(M, 12, -1) == 2 << 3 | 5 and o.get_raw(M, 12, 1) == 3 << 3 | 5      # ties -> Mismatch
    assert o.get_raw(M, 12, 0) == 10 << 3 | 5 and o.get_raw(M, 12, 3) == 4 << 3 | 2
    for s in (1, 2, 3, 5, 6, 7, 9, 11):
        assert o.krange(M, s) is None


def test_replay_property_on_oracle_output():
    """parity.replay_alignments (the size-independent check used at config 5's shard shape) accepts
    what the oracle produces for global alignments and rejects a corrupted op list."""
    import numpy as np
    import pytest
    import parity
    from wfa_b200 import datagen
    batch = datagen.generate(64, 700, 0.12, config=3)
    for ad in ((10, 50), None):
        r, o, off, _ = parity.oracle_batch(batch, adaptive=ad)
        parity.replay_alignments(batch, r, o, off)
    bad = o.copy()
    first_m = int(np.nonzero((bad >> np.uint64(32)) == ord("M"))[0][0])
    bad[first_m] = (np.uint64(ord("X")) << np.uint64(32)) | (bad[first_m] & np.uint64(0xffffffff))
    with pytest.raises(AssertionError):
        parity.replay_alignments(batch, r, bad, off)
