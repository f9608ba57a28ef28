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


def parse_cpulist(text: str) -> List[int]:
    """"0-3,8,10-11" -> [0, 1, 2, 3, 8, 10, 11] (the kernel's cpulist format)."""
    out: List[int] = []
    for part in text.replace("\n", "").split(","):
        part = part.strip()
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return sorted(set(out))


def format_cpulist(cpus: Sequence[int]) -> str:
    return ",".join(str(c) for c in cpus)


def rank_cpus(local_rank: int, local_world: int, allowed: Sequence[int], gpu_local: Sequence[Sequence[int]]) -> List[int]:
    """The CPUs of rank `local_rank` when `local_world` ranks (one GPU each) share a host: the same rule the library uses
    for the GPUs of one context (csrc/sg_host_threads.h assign_cpus) applied across processes.  gpu_local[r] = the CPUs
    the machine reports as close to rank r's GPU (empty = unknown).  Ranks whose GPUs have the same local set split it
    into equal contiguous slices, so no two ranks share a CPU."""
    allowed = sorted(set(allowed))
    eff = []
    for r in range(local_world):
        loc = [c for c in gpu_local[r] if c in set(allowed)] if r < len(gpu_local) else []
        eff.append(loc or list(allowed))
    mine = eff[local_rank]
    same = [r for r in range(local_world) if eff[r] == mine]
    k = same.index(local_rank)
    a, b = len(mine) * k // len(same), len(mine) * (k + 1) // len(same)
    return mine[a:b] if b > a else mine


def gpu_local_cpus(device_index: int) -> List[int]:
    """NUMA-local CPUs of a CUDA device from sysfs ([] when the machine does not say)."""
    import torch
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            return parse_cpulist(f.read())
    except (OSError, AttributeError, ValueError):
        return []
