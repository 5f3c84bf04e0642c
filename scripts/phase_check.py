"""Phase-separated e2e: all uploads, barrier, all runs, barrier, all downloads (K ctxs, K threads)."""
import os, sys, time, threading
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
from wfa_b200 import api, datagen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 66304
K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
overlap = len(sys.argv) > 3 and sys.argv[3] == "overlap"
full = datagen.generate_config("cfg2_150bp_e5_global", n * K)
host = [api.pinned_copy(x) for x in (full.seq_bytes, full.q_off, full.q_len, full.t_off, full.t_len)]
als = [api.New() for _ in range(K)]
bar = threading.Barrier(K + 1)
marks = {}
def work(i):
    s, qo, ql, to, tl = host
    for rep in range(4):
        bar.wait()
        rb = api.ResidentBatch(als[i], s, qo[i*n:(i+1)*n], ql[i*n:(i+1)*n], to[i*n:(i+1)*n], tl[i*n:(i+1)*n])
        if not overlap: bar.wait()
        rb.run()
        if not overlap: bar.wait()
        rb.download()
        rb.free()
        bar.wait()
th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
for x in th: x.start()
for rep in range(4):
    bar.wait(); t0 = time.perf_counter()
    if not overlap:
        bar.wait(); t1 = time.perf_counter()
        bar.wait(); t2 = time.perf_counter()
    bar.wait(); t3 = time.perf_counter()
    if overlap: print("overlapped: total %.2f ms" % ((t3 - t0) * 1e3))
    else: print("upload %.2f ms, run %.2f ms, download %.2f ms, total %.2f" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
for x in th: x.join()
