"""The host side of the C ABI on the GPU box: call statistics, launch order, input validation, big genomes, the pinned cache."""
import ctypes as C
import os
import random

import numpy as np
import pytest

from conftest import mutate, rand_seq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def al64():
    import scrooge_b200
    a = scrooge_b200.Aligner(W=64, n_gpus=1)
    yield a
    a.close()


def test_call_stats_account_for_the_bytes(al64):
    """sg_result_stats: what crossed PCIe (ASCII + packed x 4 = the ASCII input, up to chunk padding), the sub-batches, and
    times that fit inside the call."""
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["long_10kbp"]
    text, tlen, reads = synth.pairs_host(wl, 0, 4096)
    tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
    res = al64.align_pairs_blob(tb, toff, qb, qoff)
    st = res.stats()
    ascii_in = tb.nbytes + qb.nbytes
    assert st["n_devices"] == 1 and st["sub_batches"] >= 1 and st["total_ns"] > 0 and 0 < st["kernel_ns"] < st["total_ns"]
    covered = st["h2d_ascii_bytes"] + 4 * st["h2d_packed_bytes"]
    assert ascii_in <= covered <= ascii_in + 64 * (st["sub_batches"] * 2 * (ascii_in // (8 << 20) + 2)) + 4096, (ascii_in, st)
    assert st["upload_ns"] <= st["total_ns"] and st["wait_ns"] <= st["total_ns"]
    assert st["d2h_bytes"] >= 16 * 4096 + int(res.run_offsets[-1])
    assert res.total_ns == st["total_ns"] and res.kernel_ns == st["kernel_ns"]


def test_mixed_lengths_longest_first(oracle):
    """A call with 100 bp ... 30 kbp reads mixed: the launch hands them out longest first (SG_LONGEST_FIRST, on by default),
    results come back in input order and are bit-exact either way."""
    import scrooge_b200
    rng = random.Random(77)
    T, Q = [], []
    for _ in range(600):
        L = rng.choice([100, 150, 1000, 5000, 30000])
        t = rand_seq(rng, L + L // 4 + 70)
        T.append(t)
        Q.append(mutate(rng, t, L, 0.1))
    want = oracle.align_pairs(T, Q, threads=8)
    for flag in ("1", "0"):
        os.environ["SG_LONGEST_FIRST"] = flag
        try:
            a = scrooge_b200.Aligner(W=64, n_gpus=1)
            got = a.align_pairs(T, Q)
            assert list(got.edit_distances) == list(want.edit), flag
            assert got.cigars() == want.cigars, flag
            assert list(got.ref_consumed) == list(want.ref_consumed), flag
            a.close()
        finally:
            del os.environ["SG_LONGEST_FIRST"]


def test_device_api_launch_order(oracle, sglib):
    """sg_dev_align_ordered with a reversed and a random permutation: same per-alignment outputs as input order."""
    import torch
    from scrooge_b200 import device
    from scrooge_b200._lib import check
    rng = random.Random(3)
    T, Q = [], []
    for _ in range(300):
        L = rng.choice([10, 64, 200, 2000])
        t = rand_seq(rng, L + 90)
        T.append(t); Q.append(mutate(rng, t, L, 0.08))
    want = oracle.align_pairs(T, Q, threads=4)
    dev = torch.device("cuda:0")
    tb = torch.tensor(list("".join(T).encode()), dtype=torch.uint8, device=dev)
    qb = torch.tensor(list("".join(Q).encode()), dtype=torch.uint8, device=dev)
    pt, _ = device.pack_2bit(tb)
    pq, _ = device.pack_2bit(qb)
    tl = torch.tensor([len(x) for x in T], dtype=torch.int64, device=dev)
    ql = torch.tensor([len(x) for x in Q], dtype=torch.int64, device=dev)
    ts, qs = torch.cumsum(tl, 0) - tl, torch.cumsum(ql, 0) - ql
    n = len(T)
    cap = 2 * ql + 8
    slab_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cap, 0)])
    p = lambda t: int(t.data_ptr())
    st = int(torch.cuda.current_stream().cuda_stream)
    for order in (None, torch.arange(n - 1, -1, -1, dtype=torch.int32, device=dev),
                  torch.tensor(random.Random(5).sample(range(n), n), dtype=torch.int32, device=dev)):
        da = device.DeviceAligner(64, n, dev, slab_bytes=int(slab_off[-1]))
        o = da.out
        check(sglib.sg_dev_align_ordered(64, 33, p(pt), p(ts), p(tl), p(pq), p(qs), p(ql), n, 0, p(da.slab), p(slab_off), p(da.counter),
                                         p(o.edit), p(o.ref_consumed), p(o.nruns), p(o.status), None, None,
                                         p(order) if order is not None else None, st))
        torch.cuda.synchronize()
        assert o.edit.cpu().tolist() == [int(x) for x in want.edit]
        assert o.ref_consumed.cpu().tolist() == [int(x) for x in want.ref_consumed]


def test_bad_arguments_are_errors_not_crashes(sglib, al64):
    """NULL blobs with bytes to read, decreasing offsets, NULL strings with a length: SG_ERR_BAD_ARG before anything is sized."""
    out = C.c_void_p()
    off = np.array([0, 4, 8], dtype=np.uint64)
    dec = np.array([0, 8, 4], dtype=np.uint64)
    blob = b"ACGTACGT"
    h = al64._h
    assert sglib.sg_align_pairs(h, None, off.ctypes.data, blob, off.ctypes.data, 2, 0, C.byref(out)) == 3
    assert b"NULL" in sglib.sg_last_error()
    assert sglib.sg_align_pairs(h, blob, dec.ctypes.data, blob, off.ctypes.data, 2, 0, C.byref(out)) == 3
    assert b"decrease" in sglib.sg_last_error()
    assert sglib.sg_align_pairs(h, blob, off.ctypes.data, blob, dec.ctypes.data, 2, 0, C.byref(out)) == 3
    ptrs = (C.c_char_p * 2)(b"ACGT", None)
    lens = np.array([4, 4], dtype=np.uint64)
    assert sglib.sg_align_pairs_v(h, C.cast(ptrs, C.c_void_p), lens.ctypes.data, C.cast(ptrs, C.c_void_p), lens.ctypes.data, 2, 0, C.byref(out)) == 3
    # all-empty strings with NULL blobs are fine
    zero = np.zeros(3, dtype=np.uint64)
    assert sglib.sg_align_pairs(h, None, zero.ctypes.data, None, zero.ctypes.data, 2, 0, C.byref(out)) == 0
    assert sglib.sg_result_count(out) == 2
    sglib.sg_result_free(out)
    # mapping: read index / start out of range
    al64.set_reference("ACGT" * 100)
    roff = np.array([0, 4], dtype=np.uint64)
    cs, cr = np.array([0, 401], dtype=np.uint64), np.array([0, 0], dtype=np.uint32)
    assert sglib.sg_align_candidates(h, b"ACGT", roff.ctypes.data, 1, cs.ctypes.data, cr.ctypes.data, 2, 0, C.byref(out)) == 3
    assert b"candidate 1" in sglib.sg_last_error()
    cs, cr = np.array([0, 4], dtype=np.uint64), np.array([0, 1], dtype=np.uint32)
    assert sglib.sg_align_candidates(h, b"ACGT", roff.ctypes.data, 1, cs.ctypes.data, cr.ctypes.data, 2, 0, C.byref(out)) == 3


def test_mapping_on_a_gigabase_genome(oracle):
    """Read mapping against a 1.1 Gbp genome generated on the host: unaligned start_in_reference (mod 4 and mod 16 != 0),
    candidates in the last bases of the genome (insertion tails past its end), one beyond 2^30 bases; the reference's text
    is the genome suffix from the candidate (src/genasm_cpu.cpp:512-514), the oracle gets a window cut 3 L after it."""
    import scrooge_b200
    G = 1_100_000_000
    genome = np.empty(G, dtype=np.uint8)
    scrooge_b200._lib.bench_check(scrooge_b200.bench_lib().sg_synth_genome(1234, 0, G, genome.ctypes.data, None, None))
    rng = random.Random(11)
    L = 700
    starts = [1, 3, 17, 2 ** 30 + 5, 2 ** 30 - 33, G - L - 1, G - L // 2, G - 7, G - 1, G, 999_999_937, 123_456_789]
    reads, cs, cr, texts = [], [], [], []
    for r, s in enumerate(starts):
        src = genome[s:min(G, s + 2 * L)].tobytes().decode()
        q = mutate(rng, src, L, 0.1) if len(src) >= 8 else rand_seq(rng, 40)
        reads.append(q)
        for d in (0, 2):   # the true locus and one shifted by two bases
            c = max(0, s - d)
            cs.append(c); cr.append(r)
            texts.append(genome[c:min(G, c + 3 * L + 64)].tobytes().decode())
    want = oracle.align_pairs(texts, [reads[r] for r in cr], threads=4)
    assert int(np.max(want.ref_consumed)) <= 3 * L   # the cut never limits the oracle where the genome does not
    a = scrooge_b200.Aligner(W=64, n_gpus=1)
    a.set_reference(genome)
    got = a.align_candidates(reads, cs, cr)
    assert list(got.edit_distances) == list(want.edit)
    assert got.cigars() == want.cigars
    assert list(got.ref_consumed) == list(want.ref_consumed)
    a.close()


def test_pinned_cache_is_bounded_and_trimmable(sglib):
    """Result blocks are recycled through a per-process page-locked cache: sg_trim_host_cache releases it, and a context's
    destruction does too (the cache must not outlive its use)."""
    import scrooge_b200
    a = scrooge_b200.Aligner(W=64, n_gpus=1)
    r1 = a.align_pairs(["ACGT" * 50] * 2000, ["ACGT" * 40] * 2000)
    ed = list(r1.edit_distances)
    del r1
    sglib.sg_trim_host_cache()
    r2 = a.align_pairs(["ACGT" * 50] * 2000, ["ACGT" * 40] * 2000)   # re-acquires fresh blocks
    assert list(r2.edit_distances) == ed
    del r2
    a.close()


def test_bad_base_in_an_early_sub_batch_of_a_long_call():
    """A non-ACGT base that only the DEVICE sees (copies-only ingest), in the first of many sub-batches: whichever pipeline
    step notices it -- the feeder's poll during a later upload, the slot's reuse, the end of the call -- the call fails
    with SG_ERR_BAD_BASE and names the pair."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import random, scrooge_b200\n"
        "from conftest import rand_seq, mutate\n"
        "rng = random.Random(4)\n"
        "T = [rand_seq(rng, 1300) for _ in range(3000)]\n"
        "Q = [mutate(rng, t, 1000, 0.05) for t in T]\n"
        "for bad_pair, where in ((100, 'query'), (2900, 'text')):\n"
        "    T2, Q2 = list(T), list(Q)\n"
        "    if where == 'query': Q2[bad_pair] = Q2[bad_pair][:500] + 'N' + Q2[bad_pair][501:]\n"
        "    else: T2[bad_pair] = T2[bad_pair][:77] + 'n' + T2[bad_pair][78:]\n"
        "    al = scrooge_b200.Aligner(W=64, n_gpus=1)\n"
        "    try:\n"
        "        al.align_pairs(T2, Q2)\n"
        "        raise SystemExit('bad base not detected')\n"
        "    except scrooge_b200.ScroogeError as e:\n"
        "        assert e.code == 2 and f'pair {bad_pair}' in str(e) and where in str(e), str(e)\n"
        "    ok = al.align_pairs(T, Q)   # the context is usable afterwards\n"
        "    assert ok.count == 3000\n"
        "    al.close()\n"
        "print('bad base ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SG_BATCH_MB="1", SG_MIN_BATCH_UNITS="64", SG_ASCII_MIN_BYTES="0", SG_CHUNK_KB="16", SG_HOST_THREADS="0")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bad base ok" in r.stdout, r.stderr[-2000:] + r.stdout[-500:]
