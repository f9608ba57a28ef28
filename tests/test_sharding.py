"""Host-side work split, on the CPU: the split rule, and the N>1 rank layout of bench.py with world_size 2 over gloo
(every rank generates and checks its own shard; only a max of the elapsed time crosses ranks)."""
import os
import socket

import numpy as np
import pytest

from scrooge_b200 import sharding


def test_split_by_weight_covers_and_balances():
    w = [10000] * 1000 + [150] * 5000 + [0] * 10
    for parts in (1, 2, 3, 8):
        cuts = sharding.split_by_weight(w, parts)
        assert cuts[0] == 0 and cuts[-1] == len(w) and cuts == sorted(cuts) and len(cuts) == parts + 1
        loads = [sum(w[a:b]) + (b - a) for a, b in zip(cuts, cuts[1:])]
        assert max(loads) - min(loads) <= 10001 + max(loads) * 0.01
    assert sharding.split_by_weight([], 4) == [0, 0, 0, 0, 0]
    assert sharding.split_by_weight([5], 4)[-1] == 1


def test_rank_shard_partitions():
    spans = [sharding.rank_shard(r, 4, 1000) for r in range(4)]
    assert spans == [(0, 1000), (1000, 2000), (2000, 3000), (3000, 4000)]
    with pytest.raises(ValueError):
        sharding.rank_shard(4, 4, 10)


def test_rank_cpus_are_disjoint_and_numa_local():
    allowed = list(range(32))
    # unknown locality: 8 ranks share 32 CPUs in equal slices
    sets = [sharding.rank_cpus(r, 8, allowed, [[]] * 8) for r in range(8)]
    assert all(len(s) == 4 for s in sets) and sorted(c for s in sets for c in s) == allowed
    # two NUMA nodes, four GPUs each: every rank stays on its GPU's node
    local = [list(range(0, 16))] * 4 + [list(range(16, 32))] * 4
    sets = [sharding.rank_cpus(r, 8, allowed, local) for r in range(8)]
    assert sorted(c for s in sets for c in s) == allowed
    assert all(set(sets[r]) <= set(local[r]) for r in range(8))
    # a restricted affinity mask wins over the machine's list; a GPU whose node is outside the mask falls back to the mask
    sets = [sharding.rank_cpus(r, 2, [0, 1, 2, 3], [[0, 1, 2, 3, 4, 5], [16, 17]]) for r in range(2)]
    assert all(s and set(s) <= {0, 1, 2, 3} for s in sets)
    assert sharding.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.format_cpulist([4, 5, 9]) == "4,5,9"


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.binding import Oracle
        from scrooge_b200 import synth
        wl = synth.Workload("t", 300, 0.08, synth.PACBIO, 64, 1234)
        a0, a1 = sharding.rank_shard(rank, world, 40)
        text, tlen, reads = synth.pairs_host(wl, a0, a1 - a0)
        T, Q = synth.pairs_as_strings(text, tlen, reads)
        res = Oracle().align_pairs(T, Q)
        elapsed = 1.0 + rank  # stand-in for the device time
        mx = sharding.max_over_ranks(elapsed)
        dist.barrier()
        out.put((rank, a0, a1, [int(x) for x in res.edit], mx))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo(oracle):
    import torch.multiprocessing as mp
    from scrooge_b200 import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the two shards are exactly the single-process batch, in order, and the time is the max over ranks
    wl = synth.Workload("t", 300, 0.08, synth.PACBIO, 64, 1234)
    text, tlen, reads = synth.pairs_host(wl, 0, 80)
    T, Q = synth.pairs_as_strings(text, tlen, reads)
    want = [int(x) for x in oracle.align_pairs(T, Q).edit]
    assert [g[1:3] for g in got] == [(0, 40), (40, 80)]
    assert got[0][3] + got[1][3] == want
    assert all(g[4] == 2.0 for g in got)
