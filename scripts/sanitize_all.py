"""Small batches through every kernel (LANE single stage + staged, WARP 2-bit / 8-bit, CTA, render,
pipelined chunks) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_all.py
    compute-sanitizer --tool racecheck python scripts/sanitize_all.py
Every result is compared with the oracle, so a sanitizer-clean run is also a parity run."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import numpy as np
import parity
from wfa_b200 import api, datagen

def run(name, batch, runs=1, render=False, **kw):
    a = parity.make_aligner(**kw)
    rb = api.ResidentBatch(a, batch.seq_bytes, batch.q_off, batch.q_len, batch.t_off, batch.t_len)
    for _ in range(runs):
        rb.run()
    st = a.stats()
    gpu = rb.download()
    if render:
        out = rb.render(onlyAignedRegion=True, text=True)
        assert len(out.CIGAR(0)) > 0
    rb.free(); a.close()
    ref = parity.oracle_batch(batch, **kw)
    parity.assert_same(batch, gpu, ref, name)
    print("%-28s ok: lane %d warp %d cta %d 8bit %d launches %d" % (name, st["pairs_lane"], st["pairs_warp"], st["pairs_cta"], st["pairs_8bit"], st["kernel_launches"]), file=sys.stderr)

small = int(os.environ.get("SAN_SCALE", "1"))
run("lane staged (3 runs)", datagen.generate(4096 * small, 150, 0.05, config=2), runs=3, render=True)
run("lane + hand-over to warp", datagen.generate(512, 150, 0.25, config=2))
run("warp adaptive 1 kbp", datagen.generate(64, 1000, 0.10, config=3), adaptive=(10, 50), render=True)
run("warp wide ring", datagen.generate(8, 3000, 0.15, config=5), adaptive=(10, 50))
run("cta semi-global", datagen.generate(4, 600, 0.05, window=800, max_start=200, config=4), global_alignment=False, render=True)
b = datagen.generate(256, 100, 0.1, config=2)
seq = b.seq_bytes.copy(); seq[int(b.t_off[3]) + 2] = ord("N"); seq[int(b.q_off[200]) + 5] = ord("n")
run("8-bit path", datagen.Batch(seq, b.q_off, b.q_len, b.t_off, b.t_len))
# pipelined chunks (pinned buffers, wire descriptors)
big = datagen.generate(70_000, 60, 0.05, config=2)
a = parity.make_aligner()
host = [api.pinned_copy(x) for x in (big.seq_bytes, big.q_off, big.q_len, big.t_off, big.t_len)]
for _ in range(2):
    gpu = a.align_arrays(*host, copy=True)
a.close()
sub = big.slice(0, 2000)
parity.assert_same(sub, (gpu[0][:2000], gpu[1], gpu[2][:2000]), parity.oracle_batch(sub), "pipeline")
print("pipeline ok", file=sys.stderr)
