"""Distribution of the e2e time of wfacuda_align_batch (config 2, page-locked inputs)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from wfa_b200 import api, datagen
b = datagen.generate_config("cfg2_150bp_e5_global", 1000000)
host = [api.pinned_copy(x) for x in (b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)]
a = api.New()
ts = []
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    t = time.perf_counter(); a.align_arrays(*host); ts.append((time.perf_counter() - t) * 1e3)
print("e2e ms:", " ".join("%.1f" % x for x in ts))
print("median %.1f min %.1f max %.1f" % (np.median(ts[2:]), min(ts[2:]), max(ts[2:])))
a.close()
