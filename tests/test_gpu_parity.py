"""Parity of the CUDA path with the oracle, through the C ABI (host API and device API).  Bit-exact:
edit distance, CIGAR string, consumed reference prefix."""
import numpy as np
import pytest

from conftest import mutate, rand_seq, random_pairs
from oracle.binding import CONFIGS, EXTRA_CONFIGS, RefCpu, cigar_ref_consumed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aligners(sglib):
    import scrooge_b200
    assert sglib.sg_device_count() > 0, "GPU tests need a CUDA device"
    return {64: scrooge_b200.Aligner(W=64, n_gpus=1), 32: scrooge_b200.Aligner(W=32, n_gpus=1)}


def check_against_oracle(oracle, aligner, T, Q, W, O=None):
    want = oracle.align_pairs(T, Q, W=W, O=O, threads=4)
    got = aligner.align_pairs(T, Q)
    ed, rc, cg = got.edit_distances, got.ref_consumed, got.cigars()
    assert got.count == len(T)
    for k in range(len(T)):
        assert int(ed[k]) == int(want.edit[k]), (k, T[k], Q[k])
        assert cg[k] == want.cigars[k], (k, T[k], Q[k])
        assert int(rc[k]) == int(want.ref_consumed[k]) == cigar_ref_consumed(cg[k]), k
    return got


@pytest.mark.parametrize("W", [64, 32])
def test_golden_groups(aligners, golden, W):
    for name, g in golden[W]["groups"].items():
        got = aligners[W].align_pairs([x["text"] for x in g], [x["query"] for x in g])
        ed, cg = got.edit_distances, got.cigars()
        for k, x in enumerate(g):
            assert int(ed[k]) == x["edit"], (name, k)
            assert cg[k] == x["cigar"], (name, k)
            assert got.cigar(k) == x["cigar"]


@pytest.mark.parametrize("W", [64, 32])
def test_golden_mapping(aligners, golden, W):
    m = golden[W]["mapping"]
    cs = [s for l in m["locations"] for s in l]
    cr = [r for r, l in enumerate(m["locations"]) for _ in l]
    al = aligners[W]
    al.set_reference(m["genome"])
    got = al.align_candidates(m["reads"], cs, cr)
    assert [int(x) for x in got.edit_distances] == m["edit"]
    assert got.cigars() == m["cigar"]
    assert [int(x) for x in got.ref_consumed] == [cigar_ref_consumed(c) for c in m["cigar"]]


def test_reference_interface_mirror(aligners, golden):
    """align_all(texts, queries) and align_all(Genome, reads) give the same strings (reference
    src/tests.cu:273-333: all entry points must agree)."""
    import scrooge_b200 as sb
    for x in golden[64]["groups"]["differential_tests_cu"]:
        a = sb.align_all([x["text"]], [x["query"]])
        b = sb.align_all(sb.Genome(x["text"]), [sb.Read("test", x["query"], [sb.CandidateLocation(0)])])
        assert a[0].cigar == b[0].cigar == x["cigar"] and a[0].edit_distance == b[0].edit_distance == x["edit"]


@pytest.mark.parametrize("W", [64, 32])
def test_random_mixed(oracle, aligners, W):
    T, Q = random_pairs(100 + W, 6000, [0, 1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 150, 151, 400, 1000],
                        [0.0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6])
    check_against_oracle(oracle, aligners[W], T, Q, W)


def test_long_reads_config3_shape(oracle, aligners):
    """10 kbp / 10 % PacBio-like pairs from the benchmark generator (BASELINE.json configs[2])."""
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["long_10kbp"]
    text, tlen, reads = synth.pairs_host(wl, 0, 300)
    T, Q = synth.pairs_as_strings(text, tlen, reads)
    got = check_against_oracle(oracle, aligners[64], T, Q, 64)
    assert got.kernel_ns > 0


def test_short_reads_config2_shape(oracle, aligners):
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["short_150bp"]
    text, tlen, reads = synth.pairs_host(wl, 0, 20000)
    T, Q = synth.pairs_as_strings(text, tlen, reads)
    check_against_oracle(oracle, aligners[64], T, Q, 64)
    check_against_oracle(oracle, aligners[32], T, Q, 32)


def test_many_short_pairs_host_paths(oracle, aligners):
    """800 000 x 150 bp pairs in one call (two sub-batches): large enough that the host side splits its per-alignment
    descriptor loops over threads, cuts sub-batches by binary search and lets the results of both sub-batches land in
    the one pinned result block.  Checked against the oracle on a sample spread over the whole call, and on EVERY pair
    through the sequence-independent CIGAR properties (numpy)."""
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["short_150bp"]
    n = 800_000
    text, tlen, reads = synth.pairs_host(wl, 0, n)
    tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
    got = aligners[64].align_pairs_blob(tb, toff, qb, qoff)
    ed, rc, ro, runs = got.edit_distances, got.ref_consumed, got.run_offsets.astype(np.int64), got.runs
    assert len(ed) == n and int(ro[-1]) == len(runs)
    cnt, op = (runs & 63).astype(np.int64), runs >> 6
    assert cnt.min() >= 1 and cnt.max() <= 31
    csum = lambda sel: np.concatenate(([0], np.cumsum(np.where(sel, cnt, 0))))
    for sel, want in ((op != 3, np.full(n, 150)), (op != 2, rc.astype(np.int64)), (op != 0, ed)):
        c = csum(sel)
        assert np.array_equal(c[ro[1:]] - c[ro[:-1]], want)
    idx = list(range(0, n, 997)) + [n - 1]
    T, Q = synth.pairs_as_strings(text[idx], tlen[idx], reads[idx])
    want = oracle.align_pairs(T, Q, threads=4)
    assert [int(ed[i]) for i in idx] == list(want.edit)
    assert [got.cigar(i) for i in idx] == list(want.cigars)
    assert [int(rc[i]) for i in idx] == [int(x) for x in want.ref_consumed]


def test_very_long_and_unrelated(oracle, aligners):
    rng = __import__("random").Random(9)
    t = rand_seq(rng, 120000)
    T = [t, rand_seq(rng, 30000), t[:50000], "ACGT" * 5000]
    Q = [mutate(rng, t, 100000, 0.15), rand_seq(rng, 20000), mutate(rng, t, 60000, 0.05), "ACGT" * 4000 + "TTTT"]
    check_against_oracle(oracle, aligners[64], T, Q, 64)


def test_distance_only(oracle, aligners):
    T, Q = random_pairs(3, 500, [100, 150, 1000], [0.05, 0.1])
    want = oracle.align_pairs(T, Q)
    got = aligners[64].align_pairs(T, Q, distance_only=True)
    assert list(got.edit_distances) == list(want.edit)
    assert list(got.ref_consumed) == list(want.ref_consumed)
    assert len(got.run_offsets) == 0


def test_bad_base_reports_pair(aligners):
    import scrooge_b200
    with pytest.raises(scrooge_b200.ScroogeError) as e:
        aligners[64].align_pairs(["ACGT", "ACGTACGTNACGT", "AC"], ["ACG", "ACGT", "A"])
    assert e.value.code == 2 and "pair 1" in str(e.value) and "position 8" in str(e.value)
    with pytest.raises(scrooge_b200.ScroogeError):
        aligners[64].set_reference("ACGTXACGT")


def test_size_mismatch_and_missing_reference(aligners):
    import scrooge_b200
    with pytest.raises(ValueError):
        aligners[64].align_pairs(["A"], [])
    fresh = scrooge_b200.Aligner(W=64, n_gpus=1)
    with pytest.raises(scrooge_b200.ScroogeError) as e:
        fresh.align_candidates(["ACGT"], [0], [0])
    assert e.value.code == 6


def test_empty_batch(aligners):
    got = aligners[64].align_pairs([], [])
    assert got.count == 0 and got.cigars() == []


def test_pack_kernel_matches_numpy(sglib):
    import torch
    from scrooge_b200 import device
    rng = np.random.default_rng(0)
    for n in (0, 1, 15, 16, 17, 4097, 1_000_003):
        codes = rng.integers(0, 4, size=n, dtype=np.uint8)
        letters = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
        ascii_np = letters[codes + 4 * rng.integers(0, 2, size=n, dtype=np.uint8)]
        a = torch.from_numpy(ascii_np.copy()).cuda() if n else torch.empty(0, dtype=torch.uint8, device="cuda")
        packed, bad = device.pack_2bit(a)
        torch.cuda.synchronize()
        assert int(bad.item()) == -1
        words = packed.cpu().numpy().view(np.uint32)
        pad = np.zeros((len(words)) * 16, dtype=np.uint64)
        pad[:n] = codes
        expect = (pad.reshape(-1, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)
        assert np.array_equal(words, expect), n
    a = torch.from_numpy(np.frombuffer(b"ACGT" * 10 + b"ACNT", dtype=np.uint8).copy()).cuda()
    _, bad = device.pack_2bit(a)
    assert int(bad.item()) == 42


def test_device_api_resident_inputs(oracle, sglib):
    """The path bench.py times: generator, ingest, aligner and compaction all on device-resident buffers."""
    import torch
    from scrooge_b200 import device, synth
    wl = synth.Workload("t", 1000, 0.10, synth.PACBIO, 64, 4242)
    n = 4096
    dev = torch.device("cuda:0")
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, wl.read_len, wl.err, wl.ratio, wl.slack, dev)
    h_text, h_tlen, h_reads = synth.pairs_host(wl, 0, n)
    assert np.array_equal(text.cpu().numpy(), h_text) and np.array_equal(tlen.cpu().numpy().astype(np.uint64), h_tlen)
    assert np.array_equal(reads.cpu().numpy(), h_reads)
    stride, L = text.shape[1], wl.read_len
    ptext, bad_t = device.pack_2bit(text.view(-1))
    pquery, bad_q = device.pack_2bit(reads.view(-1))
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    cap = 2 * L + 8
    slab_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(64, n, dev, slab_bytes=n * cap)
    out = da.align(ptext, idx * stride, tlen, pquery, idx * L, torch.full((n,), L, dtype=torch.int64, device=dev), slab_off)
    run_off, runs = da.compact(slab_off)
    torch.cuda.synchronize()
    assert int(bad_t.item()) == -1 and int(bad_q.item()) == -1 and int(out.status.max().item()) == 0
    T, Q = synth.pairs_as_strings(h_text, h_tlen, h_reads)
    want = oracle.align_pairs(T, Q, threads=4)
    assert np.array_equal(out.edit.cpu().numpy(), want.edit)
    assert np.array_equal(out.ref_consumed.cpu().numpy().astype(np.uint64), want.ref_consumed)
    assert int(out.dc_entries.sum().item()) == want.stats["dc_entries"]  # the roofline's algorithmic work counter
    assert int(out.windows.sum().item()) == want.stats["windows"]        # unit of the delta kernel's own work
    # the whole-batch consistency checker bench.py uses: clean on these results, and it does notice a corrupted run
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    assert device.check_runs(runs, run_off, qlen, out, 64) == 0
    broken = runs.clone()
    broken[int(run_off[5].item())] ^= 0x01   # one count off by one in alignment 5
    broken[int(run_off[9].item())] = 0       # a zero-length run in alignment 9
    assert device.check_runs(broken, run_off, qlen, out, 64) == 2
    ro, rr = run_off.cpu().numpy(), runs.cpu().numpy()
    ops = "=XID"
    for k in (0, 1, 17, n - 1):
        seg = rr[ro[k]:ro[k + 1]]
        assert "".join(f"{int(b) & 63}{ops[int(b) >> 6]}" for b in seg) == want.cigars[k]
    # distance-only gives the same distances without a slab
    da2 = device.DeviceAligner(64, n, dev)
    out2 = da2.align(ptext, idx * stride, tlen, pquery, idx * L, torch.full((n,), L, dtype=torch.int64, device=dev),
                     None, distance_only=True)
    torch.cuda.synchronize()
    assert np.array_equal(out2.edit.cpu().numpy(), want.edit)


def test_cigar_properties_at_scale(oracle, aligners):
    """Size-independent property (reference validateCigarString, src/tests.cu:106-169) on a batch too large for
    string-by-string oracle comparison to be the only check: every CIGAR is a valid transformation whose edit
    count equals the reported distance."""
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["long_10kbp"]
    text, tlen, reads = synth.pairs_host(wl, 1000, 2000)
    tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
    got = aligners[64].align_pairs_blob(tb, toff, qb, qoff)
    ed, cg = got.edit_distances, got.cigars()
    T, Q = synth.pairs_as_strings(text, tlen, reads)
    for k in range(0, 2000, 7):
        assert oracle.validate_cigar(cg[k], T[k], Q[k], int(ed[k])) == 0, k
    assert 800 < float(np.mean(ed)) < 1100


def test_multi_gpu_context_matches_single(oracle, sglib):
    """Host-side scatter over several GPUs (no collective): identical results for any GPU count."""
    import scrooge_b200
    if sglib.sg_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    T, Q = random_pairs(31, 3000, [0, 50, 150, 1000, 3000], [0.02, 0.1, 0.2])
    one = scrooge_b200.Aligner(W=64, n_gpus=1).align_pairs(T, Q)
    al = scrooge_b200.Aligner(W=64, n_gpus=2)
    assert al.num_devices == 2
    two = al.align_pairs(T, Q)
    assert list(one.edit_distances) == list(two.edit_distances)
    assert one.cigars() == two.cigars()
    assert list(one.ref_consumed) == list(two.ref_consumed)
    want = oracle.align_pairs(T, Q, threads=4)
    assert two.cigars() == want.cigars
    # mapping mode: the packed reference is replicated per GPU
    rng = __import__("random").Random(5)
    genome = rand_seq(rng, 20000)
    reads, cs, cr = [], [], []
    for r in range(200):
        s = rng.randrange(0, 19000)
        reads.append(mutate(rng, genome[s:], 500, 0.1))
        for d in (0, 3):
            cs.append(max(0, s - d)); cr.append(r)
    al.set_reference(genome)
    got = al.align_candidates(reads, cs, cr)
    want = oracle.align_candidates(genome, reads, cs, cr, threads=4)
    assert list(got.edit_distances) == list(want.edit) and got.cigars() == want.cigars


@pytest.mark.parametrize("dc,forefront,host_pack", [("delta", "smem", "0"), ("delta", "smem", "1"), ("delta", "smem", "hybrid"),
                                                    ("delta", "smem", "adaptive"), ("delta", "smem", "adaptive_dma"),
                                                    ("delta", "smem", "adaptive_host"),
                                                    ("rows", "smem", "0"), ("rows", "tmem", "1"), ("rows", "smem", "1"),
                                                    ("rows", "tmem", "0")])
def test_kernel_and_ingest_variants(dc, forefront, host_pack):
    """The alignment kernel exists in two bit-identical formulations (SG_DC=delta: column-wise +-1 deltas, the default;
    SG_DC=rows: the reference's row-wise threshold vectors, with the forefront in shared memory or in tensor memory) and
    the host API has two ingest paths (pack on the device or on the host); the defaults depend on W and on the host's
    core count, so run the combinations, each in a fresh process (the choices are made once per process / context).
    "hybrid" = host packing with half of every blob's bytes sent as ASCII and packed on the device; "adaptive" = host
    threads and the copy engine share a blob chunk by chunk (the default for large blobs; forced here on the small test
    inputs with 1 KB chunks)."""
    import subprocess
    import sys
    code = (
        "import json, sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import scrooge_b200\n"
        "from oracle.binding import Oracle\n"
        "from conftest import random_pairs\n"
        "o = Oracle()\n"
        "for W in (64, 32):\n"
        "    g = json.load(open(f'tests/golden/golden_w{W}.json'))['groups']\n"
        "    T = [x['text'] for v in g.values() for x in v]; Q = [x['query'] for v in g.values() for x in v]\n"
        "    T2, Q2 = random_pairs(7 + W, 3000, [0, 1, 31, 32, 33, 64, 65, 150, 400, 2000], [0, 0.05, 0.1, 0.3])\n"
        "    T += T2; Q += Q2\n"
        "    got = scrooge_b200.Aligner(W=W, n_gpus=1).align_pairs(T, Q)\n"
        "    want = o.align_pairs(T, Q, W=W, threads=4)\n"
        "    assert list(got.edit_distances) == list(want.edit) and got.cigars() == want.cigars, W\n"
        "    assert list(got.ref_consumed) == list(want.ref_consumed)\n"
        "    al = scrooge_b200.Aligner(W=W, n_gpus=1)\n"
        "    gv = al.align_pairs_v(T, Q)\n"
        "    assert list(gv.edit_distances) == list(want.edit) and gv.cigars() == want.cigars, ('vectored', W)\n"
        "    m = json.load(open(f'tests/golden/golden_w{W}.json'))['mapping']\n"
        "    cs = [s for l in m['locations'] for s in l]; cr = [r for r, l in enumerate(m['locations']) for _ in l]\n"
        "    al.set_reference(m['genome'])\n"
        "    for res in (al.align_candidates(m['reads'], cs, cr), al.align_candidates_v(m['reads'], cs, cr)):\n"
        "        assert [int(x) for x in res.edit_distances] == m['edit'] and res.cigars() == m['cigar'], ('mapping', W)\n"
        "try:\n"
        "    scrooge_b200.Aligner(W=64, n_gpus=1).align_pairs_v(['ACGT', 'ACGTACGTNACGT'], ['ACG', 'ACGT'])\n"
        "    raise SystemExit('bad base not detected (vectored)')\n"
        "except scrooge_b200.ScroogeError as e:\n"
        "    assert e.code == 2 and 'pair 1' in str(e) and 'position 8' in str(e), str(e)\n"
        "try:\n"
        "    scrooge_b200.Aligner(W=64, n_gpus=1).align_pairs(['ACGT', 'ACGTACGTNACGT'], ['ACG', 'ACGT'])\n"
        "    raise SystemExit('bad base not detected')\n"
        "except scrooge_b200.ScroogeError as e:\n"
        "    assert e.code == 2 and 'pair 1' in str(e) and 'position 8' in str(e), str(e)\n"
        "for T, Q, where in ((['ACGT' * 200, 'ACGT' * 100 + 'N' + 'ACGT' * 100], ['ACGT', 'ACGT'], ('pair 1', 'position 400')),\n"
        "                    (['ACGN' + 'ACGT' * 400, 'ACGT' * 200], ['ACGT', 'ACGT'], ('pair 0', 'position 3')),\n"
        "                    (['ACGT' * 300, 'ACGT' * 300], ['ACGT' * 200, 'ACGT' * 100 + 'acgtn' + 'ACGT' * 100], ('pair 1', 'position 404'))):\n"
        "    try:\n"
        "        scrooge_b200.Aligner(W=64, n_gpus=1).align_pairs(T, Q)\n"
        "        raise SystemExit('bad base not detected (long strings)')\n"
        "    except scrooge_b200.ScroogeError as e:\n"
        "        assert e.code == 2 and all(x in str(e) for x in where), str(e)\n"
        "print('variant ok')\n"
    )
    import os
    env = dict(os.environ, SG_DC=dc, SG_FOREFRONT=forefront)
    if host_pack.startswith("adaptive"):   # the default for large blobs, forced here on the small test inputs with 1 KB chunks
        env.update(SG_INGEST="adaptive", SG_ASCII_MIN_BYTES="0", SG_CHUNK_KB="1")
        if host_pack == "adaptive_dma":    # no packer threads: every chunk crosses as ASCII
            env.update(SG_HOST_THREADS="0")
        if host_pack == "adaptive_host":   # no ASCII copies: the packer threads take every chunk
            env.update(SG_DMA_DEPTH="0")
    else:
        env.update(SG_HOST_PACK="1" if host_pack == "hybrid" else host_pack)
    if host_pack == "hybrid":
        env.update(SG_ASCII_PCT="50", SG_ASCII_MIN_BYTES="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("W,O", EXTRA_CONFIGS)
def test_window_configurations(oracle, golden_wo, W, O):
    """Any window configuration at run time (genasm_generic_kernel, sg_ctx_create_wo) where the reference needs a rebuild
    with -DCLI_W/-DCLI_K/-DCLI_O: goldens made by the unmodified reference at that (W, O), both interfaces, then the
    oracle on a mixed bag (empty / 1-base / W+-1 lengths, exhausted texts, unrelated pairs), distance-only mode, and --
    when it travelled to this box -- the unmodified reference itself on the same inputs."""
    import scrooge_b200
    al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
    gw = golden_wo[(W, O)]
    for name, g in gw["groups"].items():
        got = al.align_pairs([x["text"] for x in g], [x["query"] for x in g])
        ed, cg = got.edit_distances, got.cigars()
        for k, x in enumerate(g):
            assert int(ed[k]) == x["edit"] and cg[k] == x["cigar"], (name, k)
    m = gw["mapping"]
    cs = [s for l in m["locations"] for s in l]
    cr = [r for r, l in enumerate(m["locations"]) for _ in l]
    al.set_reference(m["genome"])
    got = al.align_candidates(m["reads"], cs, cr)
    assert [int(x) for x in got.edit_distances] == m["edit"] and got.cigars() == m["cigar"]
    T, Q = random_pairs(300 + 131 * W + O, 2500, [0, 1, 2, 3, 15, 16, 17, W - 1, W, W + 1, 2 * W + 1, 150, 400, 1000, 3000],
                        [0.0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6])
    got = check_against_oracle(oracle, al, T, Q, W, O)
    d = al.align_pairs(T, Q, distance_only=True)
    assert list(d.edit_distances) == list(got.edit_distances) and list(d.ref_consumed) == list(got.ref_consumed)
    if RefCpu.available(W, O):
        ref = RefCpu(W, O).align_pairs(T[:800], Q[:800], threads=4)
        assert list(ref.edit) == [int(x) for x in got.edit_distances[:800]] and ref.cigars == got.cigars()[:800]


def test_generic_kernel_equals_tuned_kernels():
    """SG_GENERIC=1 sends 64/33 and 32/17 through genasm_generic_kernel as well: same bytes out as the tuned kernels (and
    as the oracle), incl. 10 kbp pairs of the benchmark generator and the DC-entry / window counters."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import numpy as np, scrooge_b200\n"
        "from scrooge_b200 import synth\n"
        "from oracle.binding import Oracle\n"
        "from conftest import random_pairs\n"
        "o = Oracle()\n"
        "for W in (64, 32):\n"
        "    T, Q = random_pairs(900 + W, 3000, [0, 1, 31, 32, 33, 63, 64, 65, 150, 400, 2000], [0, 0.05, 0.1, 0.3, 0.6])\n"
        "    if W == 64:\n"
        "        text, tlen, reads = synth.pairs_host(synth.WORKLOADS['long_10kbp'], 0, 64)\n"
        "        T2, Q2 = synth.pairs_as_strings(text, tlen, reads)\n"
        "        T += T2; Q += Q2\n"
        "    got = scrooge_b200.Aligner(W=W, n_gpus=1).align_pairs(T, Q)\n"
        "    want = o.align_pairs(T, Q, W=W, threads=4)\n"
        "    assert list(got.edit_distances) == list(want.edit) and got.cigars() == want.cigars, W\n"
        "    assert list(got.ref_consumed) == list(want.ref_consumed)\n"
        "print('generic ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, SG_GENERIC="1"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "generic ok" in r.stdout, r.stderr[-2000:]


def test_window_limits_are_errors(sglib):
    import scrooge_b200
    for W, O in ((257, 129), (64, 64), (256, 127), (1, 0), (64, -1)):
        with pytest.raises(scrooge_b200.ScroogeError) as e:
            scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
        assert e.value.code == 3


@pytest.mark.parametrize("planes", ["global", "smem"])
def test_generic_kernel_plane_placement(planes):
    """The general kernel's op planes forced into global memory (default only for four-word windows) and into shared
    memory, for every extra window configuration: goldens of the unmodified reference + random pairs against the oracle."""
    import os
    import subprocess
    import sys
    code = (
        "import json, sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')\n"
        "import scrooge_b200\n"
        "from oracle.binding import Oracle, EXTRA_CONFIGS\n"
        "from conftest import random_pairs\n"
        "o = Oracle()\n"
        "for W, O in EXTRA_CONFIGS:\n"
        "    g = json.load(open(f'tests/golden/golden_w{W}_o{O}.json'))['groups']\n"
        "    T = [x['text'] for v in g.values() for x in v]; Q = [x['query'] for v in g.values() for x in v]\n"
        "    want = [x['cigar'] for v in g.values() for x in v]\n"
        "    T2, Q2 = random_pairs(17 + W + O, 1500, [0, 1, W - 1, W, W + 1, 150, 400, 2000], [0, 0.05, 0.1, 0.3, 0.6])\n"
        "    al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)\n"
        "    got = al.align_pairs(T + T2, Q + Q2)\n"
        "    ref = o.align_pairs(T2, Q2, W=W, O=O, threads=4)\n"
        "    cg = got.cigars()\n"
        "    assert cg[:len(want)] == want and cg[len(want):] == ref.cigars, (W, O)\n"
        "    assert list(got.edit_distances[len(want):]) == list(ref.edit) and list(got.ref_consumed[len(want):]) == list(ref.ref_consumed)\n"
        "print('planes ok')\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, SG_GENERIC_PLANES=planes), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "planes ok" in r.stdout, r.stderr[-2000:]


ODD_WINDOWS = [(2, 1), (3, 1), (5, 0), (33, 2), (63, 0), (65, 3), (100, 40), (127, 64),
               (96, 16), (127, 3), (100, 0), (128, 64),   # W - O > 63: runs longer than one run byte holds
               (128, 0), (192, 97), (255, 127), (200, 72), (256, 250)]   # up to the widest window and the longest walk


@pytest.mark.parametrize("W,O", ODD_WINDOWS)
def test_odd_window_configurations(oracle, W, O):
    """Window sizes that are not a multiple of 32 (or tiny), overlaps at both ends of the range: against the oracle, which
    agrees with the unmodified reference rebuilt at each of these (W, O) (checked in the build container; these builds are
    not shipped) and with the kernel model (tests/test_kernel_model.py::test_odd_windows_model)."""
    import scrooge_b200
    al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
    T, Q = random_pairs(700 + 131 * W + O, 1200, [0, 1, 2, 3, W - 1, W, W + 1, 2 * W + 1, 150, 400], [0.0, 0.05, 0.1, 0.3, 0.6])
    got = check_against_oracle(oracle, al, T, Q, W, O)
    d = al.align_pairs(T, Q, distance_only=True)
    assert list(d.edit_distances) == list(got.edit_distances)


def test_device_api_wide_runs(oracle, sglib):
    """W - O > 63 on the device layer: runs longer than 63 arrive as bytes with count 0 ("63 more") + the rest; the
    whole-batch checker accepts them, and joined they give the oracle's CIGAR (low error: many long '=' runs)."""
    import torch
    from scrooge_b200 import device, synth
    W, O = 128, 1
    wl = synth.Workload("t", 1000, 0.02, synth.PACBIO, W, 777, O=O)
    n = 2048
    dev = torch.device("cuda:0")
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, wl.read_len, wl.err, wl.ratio, wl.slack, dev)
    h_text, h_tlen, h_reads = synth.pairs_host(wl, 0, n)
    stride, L = text.shape[1], wl.read_len
    ptext, _ = device.pack_2bit(text.view(-1))
    pquery, _ = device.pack_2bit(reads.view(-1))
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    cap = 2 * L + 8
    slab_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    da = device.DeviceAligner(W, n, dev, slab_bytes=n * cap, O=O)
    out = da.align(ptext, idx * stride, tlen, pquery, idx * L, qlen, slab_off)
    run_off, runs = da.compact(slab_off)
    torch.cuda.synchronize()
    assert int(out.status.max().item()) == 0
    T, Q = synth.pairs_as_strings(h_text, h_tlen, h_reads)
    want = oracle.align_pairs(T, Q, W=W, O=O, threads=4)
    assert np.array_equal(out.edit.cpu().numpy(), want.edit)
    assert np.array_equal(out.ref_consumed.cpu().numpy().astype(np.uint64), want.ref_consumed)
    assert device.check_runs(runs, run_off, qlen, out, W, O) == 0
    ro, rr = run_off.cpu().numpy(), runs.cpu().numpy()
    assert int((rr[: ro[-1]] & 63 == 0).sum()) > 0, "the workload was meant to produce runs longer than 63"
    for k in range(0, n, 97):
        s, carry = [], 0
        for b in rr[ro[k]:ro[k + 1]]:
            c = int(b) & 63
            if c == 0:
                carry += 63
                continue
            s.append(f"{carry + c}{'=XID'[int(b) >> 6]}")
            carry = 0
        assert "".join(s) == want.cigars[k], k


def test_random_window_configurations(oracle):
    """Twenty random (W, O) from the whole supported range with mixed random pairs each (tools/fuzz_windows.py is the long
    version: 380 configurations x 600-1000 pairs ran clean, profiles/r01_fuzz_windows.txt)."""
    import random
    import scrooge_b200
    rng = random.Random(31337)
    for _ in range(20):
        W = rng.choice([rng.randint(2, 256), rng.choice([31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256])])
        O = rng.randint(max(0, W - 128), W - 1)
        T, Q = random_pairs(rng.randrange(1 << 30), 300, [0, 1, 2, 3, W - 1, W, W + 1, 2 * W + 1, 3 * W, 150, 400, 1500],
                            [0.0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6])
        al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
        check_against_oracle(oracle, al, T, Q, W, O)
        al.close()
