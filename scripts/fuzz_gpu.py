"""Randomised parity: random penalties / mode / heuristic, random batches (related and unrelated
pairs, lengths 1..600, ACGT or arbitrary bytes) through the C ABI against the oracle.
    python scripts/fuzz_gpu.py [seconds] [seed]"""
import os, random, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import parity
from wfa_b200 import datagen

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time(); rounds = pairs_done = 0
while time.time() - t0 < budget:
    pen = rng.choice([(4, 6, 2), (1, 0, 1), (3, 1, 2), (2, 3, 1), (5, 2, 3), (6, 5, 1), (2, 12, 2), (7, 3, 5)])
    kw = dict(mismatch=pen[0], gap_open=pen[1], gap_ext=pen[2], global_alignment=rng.random() < 0.6)
    if rng.random() < 0.5:
        kw["adaptive"] = (rng.choice([1, 5, 10, 30]), rng.choice([5, 20, 50, 200]))
    alpha = b"ACGT" if rng.random() < 0.75 else b"ACGTN acgt"
    maxlen = rng.choice([20, 150, 254, 600])
    pairs = []
    for _ in range(rng.choice([40, 400, 2500])):
        L = rng.randint(1, maxlen)
        q = bytes(rng.choice(alpha) for _ in range(L))
        if rng.random() < 0.2:
            t = bytes(rng.choice(alpha) for _ in range(rng.randint(1, maxlen)))
        else:
            t = bytearray(q); e = rng.choice([0.0, 0.02, 0.1, 0.3])
            for _ in range(int(e * L) + (rng.random() < 0.3)):
                j = rng.randrange(len(t) + 1); r = rng.random()
                if r < 0.4 and j < len(t): t[j] = rng.choice(alpha)
                elif r < 0.7: t.insert(j, rng.choice(alpha))
                elif j < len(t) and len(t) > 1: del t[j]
            t = bytes(t)
            if not kw["global_alignment"] and rng.random() < 0.5:
                t = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 40))) + t + bytes(rng.choice(alpha) for _ in range(rng.randint(0, 40)))
        pairs.append((q, t))
    batch = datagen.Batch.from_pairs(pairs)
    # a third of the rounds on the wide-wavefront workers (WIDE clusters of a random size where they apply, else the CTA worker)
    gpu_kw = {}
    os.environ.pop("WFACUDA_WIDE_CLUSTER", None)
    if rng.random() < 0.34:
        from wfa_b200 import api
        gpu_kw = dict(flags=api.FLAG_FORCE_CTA)
        c = rng.choice([None, "1", "2", "4", "8"])
        if c:
            os.environ["WFACUDA_WIDE_CLUSTER"] = c
    gpu, ref, st = parity.check(batch, what="fuzz %r %r cluster %s" % (kw, gpu_kw, os.environ.get("WFACUDA_WIDE_CLUSTER")), gpu_kw=gpu_kw, **kw)
    wide_rounds = globals().get("wide_rounds", 0) + (1 if st.get("pairs_wide", 0) else 0)
    if kw["global_alignment"]:      # semi-global stops at the first score with a start-cell hit: fewer cells than the literal scan
        assert st["cells"] == ref[3]["cells"], (kw, st["cells"], ref[3]["cells"])
    rounds += 1; pairs_done += len(pairs)
print("fuzz ok: %d rounds (%d of them with pairs on the WIDE worker), %d pairs in %.0f s" % (rounds, globals().get("wide_rounds", 0), pairs_done, time.time() - t0))
