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


def test_shard_assign_lpt(built_lib):
    """wfacuda_shard_assign: length-binned LPT.  Equal-length reads -> contiguous, equal ranges; a
    mixed batch -> estimated loads within a few percent of each other (a contiguous cut of the same
    batch by pair count is far off), inside a length bin every shard one run of consecutive pairs."""
    import numpy as np
    from wfa_b200 import api
    rng = np.random.default_rng(5)
    for n_shards in (2, 4, 8):
        q = np.full(100_000, 150, np.uint32); t = q + rng.integers(0, 8, len(q)).astype(np.uint32)
        shard_of, cost = api.shard_assign(n_shards, q, t, False)
        assert (np.diff(shard_of.astype(np.int64)) >= 0).all()                    # contiguous ranges
        counts = np.bincount(shard_of, minlength=n_shards)
        assert counts.max() - counts.min() <= 0.01 * len(q), counts
        # mixed: 90 % short reads first, then long ones (sorted input: the worst case for contiguous cuts by count)
        q = np.concatenate([rng.integers(100, 300, 90_000), rng.integers(5_000, 20_000, 9_000), rng.integers(50_000, 100_000, 1_000)]).astype(np.uint32)
        t = (q * rng.uniform(0.95, 1.05, len(q))).astype(np.uint32)
        for adaptive in (False, True):
            shard_of, cost = api.shard_assign(n_shards, q, t, adaptive)
            assert cost.max() / cost.mean() < 1.05, (n_shards, adaptive, cost)
            # inside a length bin (half octaves of n+m) every shard owns one run of consecutive pairs
            nm = q.astype(np.int64) + t
            lg = np.floor(np.log2(nm)).astype(np.int64)
            bins = 2 * lg + ((nm >> (lg - 1)) & 1)
            for b_ in np.unique(bins):
                assert (np.diff(shard_of[bins == b_].astype(np.int64)) >= 0).all(), b_
            naive = np.array([c.sum() for c in np.array_split((q.astype(np.float64) + t) ** (1 if adaptive else 2), n_shards)])
            assert naive.max() / naive.mean() > 1.5
    assert len(api.shard_assign(3, [], [], 0)[0]) == 0
    shard_of, _ = api.shard_assign(4, [10], [12], 0)
    assert len(shard_of) == 1 and shard_of[0] < 4
