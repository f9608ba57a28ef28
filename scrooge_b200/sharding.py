"""How bench.py (one rank per GPU) and the host API split work: alignments are independent, so ranks / GPUs get
disjoint contiguous shares and nothing but a barrier and a max of the elapsed time is ever exchanged."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def rank_shard(rank: int, world: int, pairs_per_rank: int) -> Tuple[int, int]:
    """Weak scaling: rank r aligns pairs [r*n, (r+1)*n) of the synthetic workload."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rank * pairs_per_rank, (rank + 1) * pairs_per_rank


def split_by_weight(weights: Sequence[int], parts: int) -> List[int]:
    """Contiguous cut points [0, ..., n] giving each part about the same total weight (+1 per item so that
    zero-weight items still spread) -- the rule sg_host_api.cu uses to scatter a call over a context's GPUs."""
    n = len(weights)
    prefix = [0]
    for w in weights:
        prefix.append(prefix[-1] + int(w) + 1)
    total = prefix[-1]
    cuts = [0]
    for k in range(1, parts):
        target = total // parts * k
        lo, hi = cuts[-1], n
        while lo < hi:
            mid = (lo + hi) // 2
            if prefix[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        cuts.append(lo)
    cuts.append(n)
    return cuts


def max_over_ranks(value: float, device=None) -> float:
    """max of a per-rank scalar (the elapsed device time) over the default process group, if there is one."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
