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


def _plan(sglib, sizes, extra, batch, max_batch, min_units, taper):
    off = np.zeros(len(sizes) + 1, dtype=np.uint64)
    np.cumsum(np.asarray(sizes, dtype=np.uint64), out=off[1:])
    cuts = np.zeros(len(sizes) + 2, dtype=np.uint64)
    k = sglib.sg_plan_sub_batches(off.ctypes.data, len(sizes), extra, batch, max_batch, min_units, taper, cuts.ctypes.data, len(cuts))
    return [int(c) for c in cuts[:k]], off


def test_sub_batch_plan_rules_and_taper(sglib):
    """The cuts of a call (sg_plan_sub_batches = what run_all uses): cover [0, n) in order, respect the byte and unit rules,
    and with taper end in pieces of about 1/2, 1/4, 1/8, 1/8 of a full sub-batch."""
    MB = 1 << 20
    n, per = 524_288, 20_102                       # the end-to-end bench call: 10 kbp pairs
    cuts, off = _plan(sglib, [per] * n, 48, 256 * MB, 3072 * MB, 113_664, 0)
    assert cuts[0] == 0 and cuts[-1] == n and cuts == sorted(set(cuts))
    sizes = np.diff(cuts)
    assert all(s == 113_664 for s in sizes[:-1]) and sizes[-1] == n - 4 * 113_664   # the unit rule decides for 20 KB pairs
    tap, _ = _plan(sglib, [per] * n, 48, 256 * MB, 3072 * MB, 113_664, 1)
    assert tap[0] == 0 and tap[-1] == n and tap == sorted(set(tap))
    ts = np.diff(tap)
    # the last sub-batch (69 632 pairs: more than half a full one, so it is not joined with its predecessor) is cut into
    # 1/2, 1/4, 1/8, 1/8
    last = n - 4 * 113_664
    assert list(ts[:4]) == [113_664] * 4 and len(ts) == 8
    assert abs(ts[4] - last / 2) <= 2 and abs(ts[5] - last / 4) <= 2 and abs(ts[6] - last / 8) <= 2 and abs(ts[7] - last / 8) <= 3
    # a short last sub-batch is joined with the one before it first
    tap2, _ = _plan(sglib, [per] * (4 * 113_664 + 1000), 48, 256 * MB, 3072 * MB, 113_664, 1)
    t2 = np.diff(tap2)
    assert list(t2[:3]) == [113_664] * 3 and len(t2) == 7 and abs(t2[3] - (113_664 + 1000) / 2) <= 2
    # short reads: the byte rule decides; nothing exceeds the maximum
    cuts, off = _plan(sglib, [300] * 2_000_000, 48, 256 * MB, 3072 * MB, 113_664, 1)
    w = np.diff(off[cuts].astype(np.int64)) + 48 * np.diff(cuts)
    assert w.max() <= 256 * MB + 348 and cuts[-1] == 2_000_000
    # one huge alignment per sub-batch when a single one exceeds the maximum; empty strings still advance
    cuts, _ = _plan(sglib, [4000 * MB, 0, 0, 5], 48, 256 * MB, 3072 * MB, 4, 0)
    assert cuts[0] == 0 and cuts[1] == 1 and cuts[-1] == 4
    # a tiny call is not tapered
    cuts, _ = _plan(sglib, [100] * 10, 48, 256 * MB, 3072 * MB, 4, 1)
    assert cuts == [0, 10]
