"""Multi-GPU plumbing for bench.py: one process per GPU (torchrun), pairs sharded,
no collective on the data path (pairs are independent).  torch.distributed is
used only for the barrier and the max-over-ranks of the step time."""
import os


def env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_first(rank, pairs_per_rank):
    """Weak scaling: rank r aligns pairs [r*P, (r+1)*P) of the deterministic stream."""
    return rank * pairs_per_rank


def reduce_times_and_totals(times, totals, world, device=None):
    """MAX over ranks of `times`, SUM over ranks of `totals` (lists of floats)."""
    if world <= 1:
        return list(times), list(totals)
    import torch
    import torch.distributed as dist
    t = torch.tensor(times, dtype=torch.float64, device=device)
    s = torch.tensor(totals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [float(x) for x in s.tolist()]
