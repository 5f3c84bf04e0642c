"""Do chunk-sized resident batches on separate ctxs run well concurrently?  (GPU-only: no PCIe in the timed part)"""
import os, sys, time, threading
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from wfa_b200 import api, datagen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 66304
K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
full = datagen.generate_config("cfg2_150bp_e5_global", n * K)
als = [api.New() for _ in range(K)]
rbs = [api.ResidentBatch(als[i], full.seq_bytes, full.q_off[i*n:(i+1)*n], full.q_len[i*n:(i+1)*n], full.t_off[i*n:(i+1)*n], full.t_len[i*n:(i+1)*n]) for i in range(K)]
for rb in rbs: rb.run(); rb.run()
t = time.perf_counter()
for rb in rbs: rb.run()
print("sequential: %.2f ms for %d chunks of %d" % ((time.perf_counter() - t) * 1e3, K, n))
for rep in range(3):
    th = [threading.Thread(target=rb.run) for rb in rbs]
    t = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    print("concurrent: %.2f ms" % ((time.perf_counter() - t) * 1e3), [round(a.stats()["ms_total_device"], 2) for a in als])
