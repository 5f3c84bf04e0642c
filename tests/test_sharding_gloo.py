"""N>1 host logic on CPU: world_size-2 gloo run of the sharding/timing plumbing
bench.py uses, and the static sharding rule of wfacuda_align_batch_multi."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as dist
from wfa_b200 import datagen, dist as wdist
import oracle_lib
rank, world, local = wdist.env()
dist.init_process_group("gloo")
P = 64
b = datagen.generate_config("cfg2_150bp_e5_global", P, first=wdist.shard_first(rank, P))
cfg = oracle_lib.make_config()
res, ops, off, ctr = oracle_lib.align_batch(cfg, b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len, threads=1)
times, totals = wdist.reduce_times_and_totals([1.0 + rank, 5.0 - rank], [float(P), float(res["score"].sum())], world)
gathered = [None] * world
dist.all_gather_object(gathered, res["score"].tolist())
if rank == 0:
    print(json.dumps({"times": times, "totals": totals, "scores": gathered}))
dist.destroy_process_group()
''' % (ROOT, ROOT)


def test_two_rank_gloo_shards_are_disjoint_and_reduced(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["times"] == [2.0, 5.0]                    # MAX over ranks
    assert d["totals"][0] == 128.0                     # SUM of pairs
    # the union of the two shards is the first 128 pairs of the stream, in order
    import oracle_lib
    from wfa_b200 import datagen
    b = datagen.generate_config("cfg2_150bp_e5_global", 128)
    res, _, _, _ = oracle_lib.align_batch(oracle_lib.make_config(), b.seq_bytes, b.q_off, b.q_len, b.t_off, b.t_len)
    assert d["scores"][0] + d["scores"][1] == res["score"].tolist()
    assert d["totals"][1] == float(res["score"].sum())


def test_shard_plan(built_lib):
    from wfa_b200 import api
    rng = np.random.default_rng(5)
    q = rng.integers(1, 5000, 10000).astype(np.uint32)
    t = (q + rng.integers(0, 50, 10000)).astype(np.uint32)
    for adaptive in (0, 1):
        for n in (1, 2, 4, 8):
            cuts = api.shard_plan(n, q, t, adaptive)
            assert cuts[0] == 0 and cuts[-1] == len(q) and np.all(np.diff(cuts.astype(np.int64)) >= 0)
            nm = q.astype(np.float64) + t
            cost = (nm if adaptive else nm * nm) + 64.0
            per = [cost[int(cuts[i]):int(cuts[i + 1])].sum() for i in range(n)]
            assert max(per) <= 1.05 * (cost.sum() / n) + cost.max()
    assert list(api.shard_plan(3, [], [], 0)) == [0, 0, 0, 0]
