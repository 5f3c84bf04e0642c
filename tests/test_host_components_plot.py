"""Host logic of the wavefront-store mirror (api.Components: ComponentView + the restatement of
(*Aligner).Plot, wfa_component_plot.go:41-209) without a GPU: the store is filled from the C
oracle in the exact layout wfacuda_align_components hands out (rows of {score, lo, hi,
first_cell} + M/I/D triples over the M range), and the Plot matrix must equal the Python
oracle's for every component and both notChangeToMatch settings."""
import io
import os
import random
import sys

import numpy as np

import oracle_lib
from wfa_b200 import api

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle  # noqa: E402


def _store_from_oracle(o):
    rows, cells = [], []
    for s in range(o.max_score() + 1 + 8):          # + the wavefront initComponents seeds at score x (wfa.go:155-183)
        kr = o.krange(0, s)
        if kr is None:
            continue
        lo, hi = kr
        rows.append((s, lo, hi, 0, len(cells)))
        for k in range(lo, hi + 1):
            cells += [o.get_raw(c, s, k) for c in range(3)]
    return np.array(rows, dtype=api.WAVEFRONT_DTYPE), np.array(cells, dtype=np.uint32)


def test_plot_matrix_equals_python_oracle():
    rng = random.Random(5)
    for it in range(16):
        L = rng.choice([4, 9, 20, 45])
        q = bytes(rng.choice(b"ACGT") for _ in range(L))
        t = bytearray(q)
        for _ in range(max(1, L // 6)):
            j = rng.randrange(len(t) + 1)
            r = rng.random()
            if r < 0.4 and j < len(t):
                t[j] = rng.choice(b"ACGT")
            elif r < 0.7:
                t.insert(j, rng.choice(b"ACGT"))
            elif j < len(t) and len(t) > 1:
                del t[j]
        t = bytes(t)
        glob = it % 2 == 0
        o = oracle_lib.Oracle(global_alignment=glob, adaptive=(10, 50))
        o.align(q, t)
        rows, cells = _store_from_oracle(o)
        comps = api.Components(api.DefaultPenalties, rows, cells)
        p = pyoracle.Aligner(global_alignment=glob, adaptive=(10, 50))
        p.Align(q, t)
        for name in ("M", "I", "D"):
            for keep in (False, True):
                assert comps.plot_matrix(q, t, name, notChangeToMatch=keep) == p.plot_matrix(q, t, name, notChangeToMatch=keep), (q, t, glob, name, keep)
        # the views answer like Component.Get / GetAfterDiff
        for s in (0, 4, 8, 12):
            for k in (-2, -1, 0, 1, 2):
                assert comps.M.GetRaw(s, k) == o.get_raw(0, s, k) and comps.I.GetRaw(s, k) == o.get_raw(1, s, k)
        assert comps.M.GetAfterDiff(2, 4, 0) == (0, 0, False)
        buf = io.StringIO()
        comps.Plot(q, t, buf, "M")
        assert buf.getvalue().count("\n") == len(q) + 2
        o.close()
