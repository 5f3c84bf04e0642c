"""Staged LANE forward: run a resident batch several times on one ctx (the second run on uses the
stage boundaries learned from the first) and compare the last run with the oracle."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import parity
from wfa_b200 import api, datagen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
b = datagen.generate_config("cfg2_150bp_e5_global", n)
a = api.New()
rb = api.ResidentBatch(a, b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)
for it in range(4):
    rb.run(); st = a.stats()
    print("run %d: align %.3f ms, launches %d, lane %d warp %d retries %d" % (it, st["ms_align"], st["kernel_launches"], st["pairs_lane"], st["pairs_warp"], st["retries"]), file=sys.stderr)
gpu = rb.download()
ref = parity.oracle_batch(b, threads=16)
parity.assert_same(b, gpu, ref, "staged")
assert a.stats()["cells"] == ref[3]["cells"]
print("staged parity ok", file=sys.stderr)
rb.free(); a.close()
