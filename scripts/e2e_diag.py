"""e2e breakdown of wfacuda_align_batch on config 2 (debug timings on stderr)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
os.environ["WFACUDA_DEBUG"] = "1"
import numpy as np
from wfa_b200 import api, datagen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2_150bp_e5_global"
b = datagen.generate_config(wl, n)
cfgc = datagen.CONFIGS[wl]
a = api.New(api.Penalties(4, 6, 2), api.Options(cfgc["global_alignment"]))
if cfgc["adaptive"]:
    a.AdaptiveReduction(api.AdaptiveReductionOption(cfgc["adaptive"][0], cfgc["adaptive"][1], 1))
host = [api.pinned_copy(x) for x in (b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)] if os.environ.get("PINNED", "1") == "1" else [b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len]
for it in range(6):
    t = time.perf_counter(); a.align_arrays(*host); dt = time.perf_counter() - t
    print("ITER %d: %.1f ms" % (it, dt * 1e3), file=sys.stderr, flush=True)
a.close()
# raw host memcpy bandwidth for reference
x = np.empty(320_000_000, dtype=np.uint8); y = np.empty_like(x); x[:] = 1; y[:] = 2
t = time.perf_counter(); y[:] = x; print("numpy memcpy 320MB: %.1f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
