"""Small LANE-worker batch for compute-sanitizer (memcheck / racecheck)."""
import os, sys, random
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import parity
from wfa_b200 import datagen
rng = random.Random(3)
b = datagen.generate(700, 150, 0.05, config=2)
gpu, ref, stats = parity.check(b, what="sanitize lane")
print("ok", stats["pairs_lane"], stats["cells"], ref[3]["cells"])
