"""e2e time of wfacuda_align_batch over pipeline worker count x chunk size (page-locked inputs)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from wfa_b200 import api, datagen
b = datagen.generate_config("cfg2_150bp_e5_global", 1000000)
host = [api.pinned_copy(x) for x in (b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)]
combos = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(8, 80009)]
for workers, chunk in combos:
    os.environ["WFACUDA_PIPE_WORKERS"] = str(workers); os.environ["WFACUDA_CHUNK_PAIRS"] = str(chunk)
    a = api.New()
    ts = []
    for it in range(8):
        t = time.perf_counter(); a.align_arrays(*host); ts.append(time.perf_counter() - t)
    a.close()
    print(workers, chunk, " ".join("%.1f" % (x * 1e3) for x in ts), flush=True)
