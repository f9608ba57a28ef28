"""The alignment kernel's own SOURCE (scrooge_b200/csrc/sg_align_delta.cuh) run on the CPU-only box: tests/sim compiles it
for the host (every CUDA thread a fiber, warp votes as rendezvous, shared memory a host array) and this file checks what
it writes -- distances, consumed prefixes, run bytes, status -- against the oracle, bit for bit, for both window
configurations and both run-emission variants.  tests/kernel_model.py models the ALGORITHM in Python; this runs the CODE.
The GPU parity tests (tests/test_gpu_parity.py) remain the proof for the hardware; the few inline-PTX blocks have C++ twins
under SG_SIM that only this simulation executes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, random_pairs

SIM_DIR = os.path.join(ROOT, "tests", "sim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libsgsim.so")
SIM_SRC = [os.path.join(SIM_DIR, "sim_kernels.cpp"), os.path.join(SIM_DIR, "sim_runtime.cpp")]
SIM_DEPS = SIM_SRC + [os.path.join(SIM_DIR, "sim_runtime.h"), os.path.join(SIM_DIR, "shim", "cuda_runtime.h")] + [
    os.path.join(ROOT, "scrooge_b200", "csrc", f) for f in ("sg_align.cuh", "sg_align_delta.cuh", "sg_align_generic.cuh", "sg_aux.cuh", "sg_bench_aux.cuh", "sg_synth.h")]
CODE = np.full(256, 255, dtype=np.uint8)
for _k, _c in enumerate("ACGT"):
    CODE[ord(_c)] = _k
    CODE[ord(_c.lower())] = _k


@pytest.fixture(scope="module")
def sim():
    os.makedirs(os.path.dirname(SIM_LIB), exist_ok=True)
    if not os.path.exists(SIM_LIB) or any(os.path.getmtime(d) > os.path.getmtime(SIM_LIB) for d in SIM_DEPS):
        subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-shared", "-fPIC", "-DSG_SIM", "-I" + os.path.join(SIM_DIR, "shim"),
                        "-I" + SIM_DIR] + SIM_SRC + ["-o", SIM_LIB], check=True)
    lib = C.CDLL(SIM_LIB)
    lib.sim_delta_align.restype = C.c_int
    lib.sim_delta_align.argtypes = [C.c_int, C.c_int, C.c_uint] + [C.c_void_p] * 6 + [C.c_uint64, C.c_uint32] + [C.c_void_p] * 10
    lib.sim_generic_align.restype = C.c_int
    lib.sim_generic_align.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint] + [C.c_void_p] * 6 + [C.c_uint64, C.c_uint32] + [C.c_void_p] * 9
    lib.sim_pack_2bit.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint]
    lib.sim_scan_runs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.sim_gather_runs.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_uint]
    lib.sim_set_schedule_seed.argtypes = [C.c_uint64]
    lib.sim_check_runs.argtypes = [C.c_void_p] * 2 + [C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_uint]
    return lib


def pack_blob(strings):
    """2 bit/base, 16 bases per little-endian word, base k of a word in bits 2k+1:2k, strings concatenated without
    alignment, 8 zero words of padding."""
    lens = np.array([len(s) for s in strings], dtype=np.uint64)
    start = np.zeros(len(strings), dtype=np.uint64)
    if len(strings) > 1:
        start[1:] = np.cumsum(lens)[:-1]
    codes = CODE[np.frombuffer("".join(strings).encode(), dtype=np.uint8)].astype(np.uint32)
    assert (codes < 4).all()
    n = len(codes)
    words = np.zeros((n + 15) // 16 + 8, dtype=np.uint32)
    padded = np.zeros(((n + 15) // 16) * 16, dtype=np.uint32)
    padded[:n] = codes
    sh = (2 * np.arange(16, dtype=np.uint32))[None, :]
    words[: (n + 15) // 16] = np.bitwise_or.reduce(padded.reshape(-1, 16) << sh, axis=1)
    return words, start, lens


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_sim(sim, W, emit, texts, queries, ctas=2, distance_only=False, order=None, cap_of=None, want_stats=True, generic=None, slab_off=None):
    """One simulated launch.  generic = None: genasm_delta_kernel<W, emit>; generic = (O, gp): genasm_generic_kernel at window
    configuration (W, O), op planes in shared (gp = 0) or global (gp = 1) memory."""
    n = len(texts)
    tw, ts, tl = pack_blob(texts)
    qw, qs, ql = pack_blob(queries)
    cap = np.array([cap_of(len(q)) if cap_of else 2 * len(q) + 8 for q in queries], dtype=np.uint64)
    if emit:
        cap = (cap + np.uint64(3)) & ~np.uint64(3)          # SG_FLAG_RUN_WORDS: slots start and end on 4-byte boundaries
    if slab_off is None:
        slab_off = np.zeros(n + 1, dtype=np.uint64)
        slab_off[1:] = np.cumsum(cap)
    slab = np.full(int(slab_off[-1]) + 16, 0xEE, dtype=np.uint8)
    assert slab.ctypes.data % 16 == 0
    edit = np.full(n, -7, dtype=np.int64)
    rc = np.full(n, 2**63, dtype=np.uint64)
    nruns = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    status = np.full(n, 0xFF, dtype=np.uint8)
    dce = np.zeros(n, dtype=np.uint64) if want_stats else None
    win = np.zeros(n, dtype=np.uint32) if want_stats else None
    counters = np.zeros(8, dtype=np.uint64)
    ordr = None if order is None else np.asarray(order, dtype=np.uint32)
    if generic is None:
        r = sim.sim_delta_align(W, emit, ctas, p(tw), p(ts), p(tl), p(qw), p(qs), p(ql), n, (1 if distance_only else 0) | (2 if emit else 0),
                                p(slab), p(slab_off), p(edit), p(rc), p(nruns), p(status), p(dce), p(win), p(ordr), p(counters))
    else:
        r = sim.sim_generic_align(W, generic[0], generic[1], ctas, p(tw), p(ts), p(tl), p(qw), p(qs), p(ql), n, 1 if distance_only else 0,
                                  p(slab), p(slab_off), p(edit), p(rc), p(nruns), p(status), p(dce), p(win), p(ordr))
    assert r == 0
    assert (slab[int(slab_off[-1]):] == 0xEE).all(), "write past the end of the slab"
    return dict(edit=edit, rc=rc, nruns=nruns, status=status, slab=slab, slab_off=slab_off, dc_entries=dce, windows=win,
                counters=counters)


def runs_text(b):
    """CIGAR text of packed runs; a byte with count 0 stands for 63 more of its op (windows with W - O > 63)."""
    s, carry = [], 0
    for x in b:
        c = int(x) & 63
        if c == 0:
            carry += 63
            continue
        s.append(f"{carry + c}{'=XID'[int(x) >> 6]}")
        carry = 0
    return "".join(s)


def cigar_of(out, a):
    return runs_text(out["slab"][int(out["slab_off"][a]): int(out["slab_off"][a]) + int(out["nruns"][a])])


def check(out, res, n, cigars=True):
    assert (out["status"] == 0).all()
    assert (out["edit"] == np.asarray(res.edit[:n], dtype=np.int64)).all()
    assert (out["rc"] == np.asarray(res.ref_consumed[:n], dtype=np.uint64)).all()
    if cigars:
        for a in range(n):
            assert cigar_of(out, a) == res.cigars[a], a


@pytest.mark.parametrize("emit", [0, 1])
@pytest.mark.parametrize("W", [64, 32])
def test_sim_matches_oracle_mixed(sim, oracle, W, emit):
    """Mixed bag (empty / 1-base / W+-1 lengths, exhausted texts, unrelated pairs, up to 40-80 % error) over several warps
    and CTAs, so that lanes refill from the queue at different times."""
    T, Q = random_pairs(101 + W + emit, 420, [0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 200, 500, 1200],
                        [0, 0.05, 0.15, 0.4, 0.8])
    res = oracle.align_pairs(T, Q, W=W)
    out = run_sim(sim, W, emit, T, Q, ctas=2)
    check(out, res, len(T))
    assert int(out["dc_entries"].sum()) == res.stats["dc_entries"]
    # runs are stored once each: as bytes (variant 0) or inside words (variant 1: one store per four runs + one flush)
    total = int(out["nruns"].sum())
    if emit == 0:
        assert int(out["counters"][0]) == total and int(out["counters"][1]) == 0
    else:
        assert int(out["counters"][0]) == 0
        assert int(out["counters"][1]) == int(((out["nruns"].astype(np.int64) + 3) // 4).sum())


@pytest.mark.parametrize("emit", [0, 1])
@pytest.mark.parametrize("W", [64, 32])
def test_sim_golden_vectors(sim, golden, W, emit):
    for group, items in golden[W]["groups"].items():
        T = [x["text"] for x in items]
        Q = [x["query"] for x in items]
        out = run_sim(sim, W, emit, T, Q, ctas=1)
        for a, x in enumerate(items):
            assert int(out["edit"][a]) == x["edit"] and cigar_of(out, a) == x["cigar"], (group, a)


@pytest.mark.parametrize("emit", [0, 1])
def test_sim_order_distance_only_and_overflow(sim, oracle, emit):
    T, Q = random_pairs(7, 150, [10, 100, 300, 700], [0.05, 0.3])
    res = oracle.align_pairs(T, Q, W=64)
    order = np.argsort([-len(q) for q in Q], kind="stable")
    out = run_sim(sim, 64, emit, T, Q, ctas=1, order=order)
    check(out, res, len(T))
    out = run_sim(sim, 64, emit, T, Q, ctas=3, distance_only=True, want_stats=False)
    check(out, res, len(T), cigars=False)
    # slots too small for some alignments: those report status 5 and 0 runs, the others are intact, nothing is
    # written outside a slot
    out = run_sim(sim, 64, emit, T, Q, ctas=1, cap_of=lambda L: 8)
    for a in range(len(T)):
        want_runs = sum(1 for c in res.cigars[a] if c in "=XID")
        if want_runs <= 8:
            assert out["status"][a] == 0 and cigar_of(out, a) == res.cigars[a]
        else:
            assert out["status"][a] == 5 and out["nruns"][a] == 0
    assert (out["edit"] == np.asarray(res.edit, dtype=np.int64)).all()


def test_sim_unrelated_candidates_store_counts(sim, oracle):
    """Windows that are mostly edits (a read against a random locus: the 7 spurious candidates of the stress mapping
    workload): the word variant issues a quarter of the run stores."""
    T, Q = random_pairs(99, 64, [600], [0.0], unrelated=1.0, short_text=0.0)
    res = oracle.align_pairs(T, Q, W=64)
    a = run_sim(sim, 64, 0, T, Q, ctas=1)
    b = run_sim(sim, 64, 1, T, Q, ctas=1)
    check(a, res, len(T))
    check(b, res, len(T))
    runs_per_window = a["nruns"].sum() / a["windows"].sum()
    assert runs_per_window > 15          # 6.4 at 10 % error
    assert int(b["counters"][1]) * 3.9 < int(a["counters"][0])


def _generic_cases():
    from oracle.binding import EXTRA_CONFIGS
    odd = [(2, 1), (5, 0), (33, 2), (65, 3), (100, 40), (127, 64), (192, 97), (255, 127), (256, 250), (200, 72)]
    return [(W, O, gp) for (W, O) in list(EXTRA_CONFIGS) + [(64, 33), (32, 17)] + odd for gp in (0, 1)]


@pytest.mark.parametrize("W,O,gp", _generic_cases())
def test_sim_generic_kernel_matches_oracle(sim, oracle, W, O, gp):
    """genasm_generic_kernel (any window configuration at run time) on every configuration the reference was rebuilt for, the
    two tuned ones and ten odd shapes, op planes in shared and in global memory: against the oracle (which is pinned to the
    unmodified reference at these configurations, tests/test_oracle.py)."""
    T, Q = random_pairs(1000 + 7 * W + O, 96, [0, 1, 2, W // 2, W - 1, W, W + 1, 2 * W + 1, 100, 300, 700], [0, 0.05, 0.15, 0.4, 0.8])
    res = oracle.align_pairs(T, Q, W=W, O=O)
    out = run_sim(sim, W, 0, T, Q, ctas=2, generic=(O, gp))
    check(out, res, len(T))
    assert int(out["dc_entries"].sum()) == res.stats["dc_entries"] and int(out["windows"].sum()) == res.stats["windows"]


def test_sim_generic_kernel_golden_vectors(sim, golden_wo):
    for (W, O), g in golden_wo.items():
        for group, items in g["groups"].items():
            T = [x["text"] for x in items]
            Q = [x["query"] for x in items]
            out = run_sim(sim, W, 0, T, Q, ctas=1, generic=(O, 0))
            for a, x in enumerate(items):
                assert int(out["edit"][a]) == x["edit"] and cigar_of(out, a) == x["cigar"], (W, O, group, a)


def test_sim_ingest_kernel_matches_numpy(sim):
    """pack_2bit_kernel: SWAR conversion, case folding, tail words, padding words, and the smallest offending position."""
    rng = np.random.default_rng(5)
    for n in (0, 1, 15, 16, 17, 1000, 4099, 70001):
        letters = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
        a = letters[rng.integers(0, 8, n)].copy()
        buf = np.zeros(n + 64, dtype=np.uint8)     # 16-byte aligned source with room for the last vector load
        buf[:n] = a
        n_words = (n + 15) // 16 + 8
        packed = np.full(n_words, 0xDEADBEEF, dtype=np.uint32)
        bad = np.full(1, 2**64 - 1, dtype=np.uint64)
        sim.sim_pack_2bit(p(buf), n, p(packed), n_words, p(bad), 0, 3)
        want, _, _ = pack_blob([a.tobytes().decode()])
        assert np.array_equal(packed, want) and int(bad[0]) == 2**64 - 1, n
        if n >= 17:
            for pos in (0, 16, n - 1, n // 2):
                b2 = buf.copy()
                b2[pos] = ord("N")
                b2[min(n - 1, pos + 5)] = ord("x")
                bad[0] = 2**64 - 1
                sim.sim_pack_2bit(p(b2), n, p(packed), n_words, p(bad), 0, 2)
                assert int(bad[0]) == pos, (n, pos)


@pytest.mark.parametrize("group", [32, 4])
def test_sim_scan_and_gather_match_numpy(sim, group):
    """scan_* + gather_runs_kernel: exclusive scan of the run counts and the gather of every slot into the dense array, at
    every relative misalignment of source and destination (byte-granular slots, word-granular copies)."""
    rng = np.random.default_rng(9 + group)
    for n, hi in ((1, 5), (7, 40), (2049, 9), (5000, 70)):
        nruns = rng.integers(0, hi, n).astype(np.uint32)
        cap = nruns.astype(np.uint64) + rng.integers(0, 6, n).astype(np.uint64)
        slab_off = np.zeros(n + 1, dtype=np.uint64)
        slab_off[1:] = np.cumsum(cap)
        slab = rng.integers(0, 256, int(slab_off[-1]) + 16).astype(np.uint8)
        run_off = np.full(n + 1, 2**63, dtype=np.uint64)
        tmp = np.zeros(n // 2048 + 4, dtype=np.uint64)
        sim.sim_scan_runs(p(nruns), n, p(run_off), p(tmp))
        want_off = np.zeros(n + 1, dtype=np.uint64)
        want_off[1:] = np.cumsum(nruns.astype(np.uint64))
        assert np.array_equal(run_off, want_off)
        total = int(want_off[-1])
        runs = np.full(total + 16, 0x5A, dtype=np.uint8)
        sim.sim_gather_runs(group, p(slab), p(slab_off), p(nruns), p(run_off), n, p(runs), 3)
        want = np.concatenate([slab[int(slab_off[a]): int(slab_off[a]) + int(nruns[a])] for a in range(n)] + [np.zeros(0, np.uint8)])
        assert np.array_equal(runs[:total], want)
        assert (runs[total:] == 0x5A).all(), "wrote past the dense array"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sim_results_do_not_depend_on_the_schedule(sim, oracle, seed):
    """The fibers of a CTA scheduled in a fresh pseudo-random order every sweep: lanes take other alignments from the queue and
    reach the votes in other orders, the results are the same (tuned kernel with both emissions, general kernel, scan + gather)."""
    T, Q = random_pairs(77 + seed, 300, [0, 1, 33, 64, 65, 150, 500, 1200], [0, 0.05, 0.15, 0.4])
    res = oracle.align_pairs(T, Q, W=64)
    sim.sim_set_schedule_seed(seed)
    try:
        for emit in (0, 1):
            check(run_sim(sim, 64, emit, T, Q, ctas=2), res, len(T))
        check(run_sim(sim, 64, 0, T, Q, ctas=3, generic=(33, 0)), res, len(T))
        rng = np.random.default_rng(seed)
        n = 3000
        nruns = rng.integers(0, 30, n).astype(np.uint32)
        slab_off = np.zeros(n + 1, dtype=np.uint64)
        slab_off[1:] = np.cumsum(nruns.astype(np.uint64) + 3)
        slab = rng.integers(0, 256, int(slab_off[-1]) + 16).astype(np.uint8)
        run_off = np.zeros(n + 1, dtype=np.uint64)
        tmp = np.zeros(8, dtype=np.uint64)
        sim.sim_scan_runs(p(nruns), n, p(run_off), p(tmp))
        assert int(run_off[-1]) == int(nruns.sum())
        runs = np.zeros(int(run_off[-1]) + 16, dtype=np.uint8)
        sim.sim_gather_runs(32, p(slab), p(slab_off), p(nruns), p(run_off), n, p(runs), 2)
        want = np.concatenate([slab[int(slab_off[a]): int(slab_off[a]) + int(nruns[a])] for a in range(n)])
        assert np.array_equal(runs[: len(want)], want)
    finally:
        sim.sim_set_schedule_seed(0)


def test_sim_host_api_slab_layout(sim, oracle):
    """The slab layout the host API hands the kernel for blob inputs (sg_host_threads.h: slab_offset_blob, every offset rounded
    up to a multiple of 4 on its own) with the word-storing kernel: every alignment intact, nothing outside its slot."""
    T, Q = random_pairs(4242, 300, [0, 1, 2, 3, 5, 17, 33, 64, 150, 151, 999], [0, 0.1, 0.5])
    res = oracle.align_pairs(T, Q, W=64)
    qpre = np.concatenate([[0], np.cumsum([len(q) for q in Q])]).astype(np.uint64)
    k = np.arange(len(Q) + 1, dtype=np.uint64)
    off = (np.uint64(2) * qpre + np.uint64(12) * k + np.uint64(3)) & ~np.uint64(3)          # slab_offset_blob(q_prefix, k, words = true)
    assert (off % 4 == 0).all() and ((off[1:] - off[:-1]) >= 2 * np.array([len(q) for q in Q], dtype=np.uint64) + 8).all()
    out = run_sim(sim, 64, 1, T, Q, ctas=2, slab_off=off)
    check(out, res, len(T))
    for a in range(len(T)):   # the bytes of a slot beyond its runs and their padding word are untouched
        used = (int(out["nruns"][a]) + 3) & ~3
        assert (out["slab"][int(off[a]) + used: int(off[a + 1])] == 0xEE).all(), a


def test_sim_batch_checker_catches_every_kind_of_inconsistency(sim, oracle):
    """check_runs_kernel (what bench.py runs over the whole timed batch): clean on real results, and each kind of damage to a
    run byte, a distance, a consumed prefix or a query length is counted once per damaged alignment."""
    T, Q = random_pairs(31, 200, [1, 33, 64, 150, 700], [0, 0.1, 0.4], short_text=0.0)
    T, Q = zip(*[(t, q) for t, q in zip(T, Q) if len(q)])
    out = run_sim(sim, 64, 1, list(T), list(Q), ctas=2)
    n = len(T)
    run_off = np.zeros(n + 1, dtype=np.uint64)
    run_off[1:] = np.cumsum(out["nruns"].astype(np.uint64))
    runs = np.concatenate([out["slab"][int(out["slab_off"][a]): int(out["slab_off"][a]) + int(out["nruns"][a])] for a in range(n)])
    qlen = np.array([len(q) for q in Q], dtype=np.uint64)

    def count_bad(runs=runs, qlen=qlen, edit=out["edit"], rc=out["rc"], max_count=31):
        bad = np.zeros(1, dtype=np.uint64)
        sim.sim_check_runs(p(np.ascontiguousarray(runs)), p(run_off), n, p(qlen), p(edit), p(rc), max_count, p(bad), 2)
        return int(bad[0])

    assert count_bad() == 0
    r = runs.copy(); r[int(run_off[3])] ^= 1                      # a count off by one
    assert count_bad(runs=r) == 1
    r = runs.copy(); r[int(run_off[5])] &= 0xC0                   # a zero-length run
    assert count_bad(runs=r) == 1
    r = runs.copy(); r[int(run_off[7])] = (r[int(run_off[7])] & 0xC0) | 40   # longer than a window's walk
    assert count_bad(runs=r) == 1
    r = runs.copy(); r[int(run_off[9])] ^= 0x40                   # '=' <-> 'X' (or 'I' <-> 'D'): the edit count (or the lengths) change
    assert count_bad(runs=r) == 1
    e = out["edit"].copy(); e[11] += 1; e[12] -= 1
    assert count_bad(edit=e) == 2
    c = out["rc"].copy(); c[13] += 1
    assert count_bad(rc=c) == 1
    q = qlen.copy(); q[0] += 1; q[n - 1] += 1
    assert count_bad(qlen=q) == 2
    assert count_bad(max_count=15) > 0                            # the 32/17 limit on 64/33 results
