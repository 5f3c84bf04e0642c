"""Report text of the wfa-go compatible CLI against the blocks the
reference prints in its README.  On CPU the formatter is fed with oracle results (test
infrastructure); on a GPU the whole CLI runs through libwfacuda."""
import io
import os
import sys
from contextlib import redirect_stdout

import pytest

import oracle_lib
from wfa_b200 import api, cli

README_BLOCK_1 = """query   ---------Bioinformatics ---helps Biology---
                  ||||||||||||||   |||| | |||||
target  We learn bioinformatics to help- biologists
cigar   9I1X14M3I4M1D1M1X5M1X3I

align-score : 32
match-region: q[2, 27]/28 vs t[11, 38]/42
align-length: 29, matches: 24 (82.76%), gaps: 4, gap regions: 2

"""                                                                         # README.md:18-27 (-g)

README_BLOCK_2 = """query   AGCTA-GTGTCAATGGCTACT---TTTCAGGTCCT
        | ||| |||||  ||||||||   | |||||||||
target  AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT
cigar   1M1X3M1I5M2X8M3I1M1X9M

align-score : 36
match-region: q[1, 31]/31 vs t[1, 35]/35
align-length: 35, matches: 27 (77.14%), gaps: 4, gap regions: 2

"""                                                                         # README.md:231-240


class _FromOracle(api.AlignmentResult):
    def __init__(self, r):
        import numpy as np
        rec = {k: r[k] for k in ("score", "tbegin", "tend", "qbegin", "qend", "align_len", "matches", "gaps", "gap_regions")}
        super().__init__(rec, np.array(r["ops"], dtype=np.uint64))


def _norm(text):
    return "\n".join(line.rstrip() for line in text.split("\n"))


def test_report_text_from_oracle_results():
    q, t = b"Bioinformatics helps Biology", b"We learn bioinformatics to help biologists"
    r = oracle_lib.Oracle(global_alignment=False, adaptive=(10, 50)).align(q, t)
    assert _norm(cli.format_report(_FromOracle(r), q, t)) == _norm(README_BLOCK_1)
    q, t = b"AGCTAGTGTCAATGGCTACTTTTCAGGTCCT", b"AACTAAGTGTCGGTGGCTACTATATATCAGGTCCT"
    r = oracle_lib.Oracle(adaptive=(10, 50)).align(q, t)
    assert _norm(cli.format_report(_FromOracle(r), q, t)) == _norm(README_BLOCK_2)
    # -t: only the aligned region (wfa_cigar.go:217-233, :266-270)
    q, t = b"Bioinformatics helps Biology", b"We learn bioinformatics to help biologists"
    r = _FromOracle(oracle_lib.Oracle(global_alignment=False, adaptive=(10, 50)).align(q, t))
    assert r.CIGAR(True) == "14M3I4M1D1M1X5M"
    Q, A, T = r.AlignmentText(q, t, True)
    assert Q == b"ioinformatics ---helps Biolog" and T == b"ioinformatics to help- biolog" and set(A) <= set(b"| ")


@pytest.mark.gpu
def test_cli_end_to_end(built_lib, tmp_path):
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert cli.main(["-g", "Bioinformatics helps Biology", "We learn bioinformatics to help biologists"]) == 0
    assert _norm(buf.getvalue()) == _norm(README_BLOCK_1)
    buf = io.StringIO()
    with redirect_stdout(buf):          # -t: only the aligned region, strings rendered by the GPU
        assert cli.main(["-g", "-t", "Bioinformatics helps Biology", "We learn bioinformatics to help biologists"]) == 0
    out = buf.getvalue()
    assert "query   ioinformatics ---helps Biolog\n" in out and "target  ioinformatics to help- biolog\n" in out
    assert "cigar   14M3I4M1D1M1X5M\n" in out
    import json
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))
    path = tmp_path / "seqs.txt"
    path.write_text("".join(">%s\n<%s\n" % (p["q"], p["t"]) for p in G["seqs_txt"]))
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert cli.main(["-i", str(path)]) == 0
    out = buf.getvalue()
    assert "cigar   1X1I14M1D39M1D31M1D12M" in out and "match-region: q[2, 100]/100 vs t[3, 98]/98" in out   # README.md:245-254
    assert out.count("align-score") == 2
