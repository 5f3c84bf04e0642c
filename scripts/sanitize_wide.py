"""Small batches through the WIDE worker (clusters of 1, 2 and 4 CTAs, global and semi-global) for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_wide.py     (also synccheck)
Every result is compared with the oracle."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import parity
from wfa_b200 import api, datagen

for cl in ("1", "2", "4"):
    os.environ["WFACUDA_WIDE_CLUSTER"] = cl
    for glob, batch in ((False, datagen.generate(6, 600, 0.05, window=800, max_start=200, config=4)), (True, datagen.generate(6, 500, 0.2, config=2, first=100))):
        gpu, ref, st = parity.check(batch, what="wide cluster %s glob %s" % (cl, glob), global_alignment=glob, gpu_kw=dict(flags=api.FLAG_FORCE_CTA))
        assert st["pairs_wide"] == len(batch), st
        print("wide cluster=%s global=%s ok: %d pairs, %d launches" % (cl, glob, st["pairs_wide"], st["kernel_launches"]), file=sys.stderr)
