"""The input pairs the reference lists in its own test (wfa_test.go:45-141; fixture
tests/golden/wfa_test_pairs.json, made by tests/golden/make_fixture_pairs.py).  The reference test
asserts nothing, so: the two alignments it notes in comments are checked as known answers, and on
all pairs the C oracle, the Python oracle and (with a GPU) libwfacuda must agree -- in TestWFA's own
configuration (upper-cased, global, 4/6/2, wf-adaptive 10/50/1, :31-44, :143-145) and semi-global."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle  # noqa: E402

PAIRS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "wfa_test_pairs.json")))["pairs"]
CONFIGS = [dict(adaptive=(10, 50)), dict(global_alignment=False, adaptive=(10, 50)), dict()]


def _upper(p):
    return p["q"].upper().encode(), p["t"].upper().encode()


def test_known_answers_noted_in_the_reference_test():
    pinned = [p for p in PAIRS if "cigar" in p]
    assert len(pinned) == 2
    o = oracle_lib.Oracle(adaptive=(10, 50))
    for p in pinned:
        q, t = _upper(p)
        assert oracle_lib.ops_to_cigar(o.align(q, t)["ops"]) == p["cigar"], p["source"]
    o.close()


@pytest.mark.parametrize("cfg", CONFIGS, ids=["global-adaptive", "semiglobal-adaptive", "global"])
def test_c_and_python_oracles_agree_on_reference_inputs(cfg):
    o = oracle_lib.Oracle(**cfg)
    for p in PAIRS:
        q, t = _upper(p)
        if len(q) * len(t) > 300_000 and not cfg.get("adaptive"):
            continue                                  # the pure-Python restatement needs minutes without the heuristic
        if len(q) > 1200 and not cfg.get("global_alignment", True):
            continue
        r = o.align(q, t)
        a = pyoracle.Aligner(**cfg)
        pr = a.Align(q, t)
        assert (pr.Score, pr.CIGAR(False), pr.QBegin, pr.QEnd, pr.TBegin, pr.TEnd) == \
               (r["score"], oracle_lib.ops_to_cigar(r["ops"]), r["qbegin"], r["qend"], r["tbegin"], r["tend"]), (p["source"], cfg)
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("upper", [True, False], ids=["upper-cased", "as-written"])
def test_gpu_matches_oracle_on_reference_inputs(built_lib, upper):
    """One batch with all 25 pairs per configuration; as written, the lower-case and text pairs take
    the 8-bit path (the reference compares raw bytes)."""
    import parity
    from wfa_b200 import datagen
    pairs = [_upper(p) if upper else (p["q"].encode(), p["t"].encode()) for p in PAIRS]
    batch = datagen.Batch.from_pairs(pairs)
    for cfg in CONFIGS:
        gpu, ref, st = parity.check(batch, what="wfa_test.go pairs %r" % (cfg,), **cfg)
        assert int((gpu[0]["status"] == 0).sum()) == len(pairs)
        if upper and cfg == CONFIGS[0]:
            for i, p in enumerate(PAIRS):
                if "cigar" in p:
                    a = int(gpu[2][i])
                    assert oracle_lib.ops_to_cigar(gpu[1][a:a + int(gpu[0]["n_ops"][i])]) == p["cigar"], p["source"]
