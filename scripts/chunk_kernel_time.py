"""Device time of one resident batch as a function of its size (how efficient are chunk-sized launches?)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
from wfa_b200 import api, datagen
full = datagen.generate_config("cfg2_150bp_e5_global", 1000000)
a = api.New()
for n in [int(x) for x in sys.argv[1:]] or [33152, 66304, 80009, 132608, 265216, 1000000]:
    rb = api.ResidentBatch(a, full.seq_bytes, full.q_off[:n], full.q_len[:n], full.t_off[:n], full.t_len[:n])
    ts = []
    for it in range(6):
        rb.run(); st = a.stats(); ts.append((st["ms_pack"], st["ms_align"], st["ms_total_device"]))
    rb.free()
    p, al, tot = ts[-1]
    print("%8d pairs: pack %.3f align %.3f total %.3f ms  -> %.2f ms per 1M pairs" % (n, p, al, tot, tot * 1e6 / n), flush=True)
a.close()
