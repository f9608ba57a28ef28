"""Measurement legs shared by bench.py (`extra_configs`) and tools/bench_configs.py: BASELINE.json's non-headline
configurations -- configs[1] short reads (W64/O33 and the reference's short-read setting W32/O17), configs[3] read mapping
on a replicated 3 Gbp genome with 8 candidates per read (+ SURVEY 8d's stress variant, 1 true + 7 uniform-random loci),
configs[4] sweep points -- each device-timed with inputs resident in HBM, end to end through the host C ABI, and checked
against the oracle on a sample of its own inputs.  The oracle is imported here as the CHECKER only; nothing it computes is
timed or shipped."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker only: never timed)
from scrooge_b200 import device, synth  # noqa: E402
from scrooge_b200._lib import bench_check, check as sg_check  # noqa: E402
from scrooge_b200.api import Result  # noqa: E402


def lib():
    return scrooge_b200.lib()


dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
p = lambda t: int(t.data_ptr())
stream = lambda: int(torch.cuda.current_stream().cuda_stream)


class _PerWord(dict):
    def __init__(self, per_word):
        super().__init__()
        self.per_word = per_word

    def __missing__(self, W):
        return self.per_word * ((W + 31) // 32)


OPS = _PerWord(7)          # reference formulation: INT32 ops per R[d][i] entry: 7 per 32-bit word
COL_OPS = _PerWord(10)     # delta kernel: INT32 ops per window column (see bench.py): 10 per 32-bit word


def runs_to_cigar(seg):
    """packed run bytes -> CIGAR text; a byte with count 0 stands for 63 more of the same op (W - O > 63, SG_RUN_COUNT)"""
    out, carry = [], 0
    for b in seg:
        c = int(b) & 63
        if c == 0:
            carry += 63
            continue
        out.append(f"{carry + c}{'=XID'[int(b) >> 6]}")
        carry = 0
    return "".join(out)


def time_steps(fn, steps=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def pairs_point(wl, n, distance_only, peak_gops, check=256):
    """n pairs of workload wl, inputs resident in HBM; kernel-only and ingest+align+compaction timings."""
    L, W, O = wl.read_len, wl.W, wl.overlap
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, L, wl.err, wl.ratio, wl.slack, dev)
    stride = text.shape[1]
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    tstart, qstart = idx * stride, idx * L
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    cap = (2 * L + 8 + 3) & ~3   # slots on 4-byte boundaries: runs stored as whole words (SG_FLAG_RUN_WORDS), as the host API's launches do
    slab_off = None if distance_only else torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(W, n, dev, slab_bytes=0 if distance_only else n * cap, O=O)
    ptext, bad_t = device.pack_2bit(text.view(-1))
    pquery, bad_q = device.pack_2bit(reads.view(-1))

    def kernel():
        da.align(ptext, tstart, tlen, pquery, qstart, qlen, slab_off, distance_only=distance_only, run_words=not distance_only)

    kernel()
    runs = None
    if not distance_only:
        runs = da.compact(slab_off)[1]

    def step():
        scrooge_b200._lib.check(lib().sg_dev_pack_2bit(p(text), n * stride, p(ptext), p(bad_t), stream()))
        scrooge_b200._lib.check(lib().sg_dev_pack_2bit(p(reads), n * L, p(pquery), p(bad_q), stream()))
        kernel()
        if not distance_only:
            da.compact(slab_off, runs)

    ms_kernel = time_steps(kernel)
    ms_step = time_steps(step)
    assert int(bad_t) == -1 and int(bad_q) == -1 and int(da.out.status.max()) == 0
    entries = int(da.out.dc_entries.sum())
    windows = int(da.out.windows.sum())
    # parity on a sample of this very batch
    k = min(check, n)
    h_text, h_tlen, h_reads = synth.pairs_host(wl, 0, k)
    T, Q = synth.pairs_as_strings(h_text, h_tlen, h_reads)
    want = Oracle().align_pairs(T, Q, W=W, O=O, threads=8)
    ok = bool(np.array_equal(da.out.edit[:k].cpu().numpy(), want.edit)) and \
        bool(np.array_equal(da.out.ref_consumed[:k].cpu().numpy().astype(np.uint64), want.ref_consumed))
    if not distance_only:
        ro = da.run_off[: k + 1].cpu().numpy()
        rr = runs[: int(ro[-1])].cpu().numpy()
        for a in range(0, k, 8):
            ok = ok and runs_to_cigar(rr[ro[a]:ro[a + 1]]) == want.cigars[a]
    gops = windows * W * COL_OPS[W] / (ms_kernel / 1e3) / 1e9
    ref_gops = entries * OPS[W] / (ms_kernel / 1e3) / 1e9
    out = {"workload": wl.name, "read_len": L, "error_rate": wl.err, "W": W, "O": O, "pairs": n, "mode": "distance_only" if distance_only else "full_cigar",
           "alignments_per_s_kernel": n / (ms_kernel / 1e3), "alignments_per_s_step": n / (ms_step / 1e3), "kernel_ms": ms_kernel,
           "step_ms": ms_step, "gcups_kernel": n / (ms_kernel / 1e3) * L * L / 1e9, "dc_entries_per_alignment": entries / n,
           "windows_per_alignment": windows / n, "int32_frac": gops / peak_gops, "reference_formulation_ratio": ref_gops / peak_gops,
           "mean_edit_distance": float(da.out.edit.double().mean()),
           "parity": {"checked": k, "bit_exact": ok}}
    del text, reads, ptext, pquery, da, runs
    torch.cuda.empty_cache()
    return out




def pairs_e2e(wl, n, steps=2, render=False):
    """The same workload end to end through sg_align_pairs: pinned host ASCII blobs in, distances + packed runs out (and,
    with render, CIGAR text for every alignment through sg_result_render_all), copies inside the timed region."""
    text, tlen, reads = synth.pairs_host(wl, 0, n)
    tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
    del text, reads
    tb_pin, qb_pin = torch.from_numpy(tb).pin_memory(), torch.from_numpy(qb).pin_memory()
    al = scrooge_b200.Aligner(W=wl.W, O=wl.overlap, device_ids=[torch.cuda.current_device()])
    res = None
    for _ in range(2):   # two untimed calls: the second one still allocates result blocks (the first result is alive while it runs)
        res = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
    t0 = time.perf_counter()
    for _ in range(steps):
        res = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
        if render:
            res.cigar_text()
    dt = (time.perf_counter() - t0) / steps
    st = res.stats()
    out = {"value": n / dt, "unit": "alignments/s", "pairs_per_step": n, "ms_per_step": dt * 1e3, "rendered": render,
           "h2d_bytes_per_step": st["h2d_ascii_bytes"] + st["h2d_packed_bytes"] + st["h2d_other_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
           "input_ascii_bytes_per_step": int(tb.nbytes + qb.nbytes)}
    ed = np.asarray(res.edit_distances)
    al.close()
    return out, ed


def pairs_leg(wl, n, peak_gops, n_e2e, check=2048):
    """One extra_configs entry of bench.py for a pair workload: device-timed step + kernel, e2e, parity sample."""
    pt = pairs_point(wl, n, False, peak_gops, check=check)
    e2e, ed = pairs_e2e(wl, n_e2e)
    # the end-to-end distances of the first pairs against the oracle as well (same generator, same pairs)
    k = min(check, n_e2e)
    h_text, h_tlen, h_reads = synth.pairs_host(wl, 0, k)
    T, Q = synth.pairs_as_strings(h_text, h_tlen, h_reads)
    want = Oracle().align_pairs(T, Q, W=wl.W, O=wl.overlap, threads=8)
    e2e["parity"] = {"checked": k, "bit_exact": bool(np.array_equal(ed[:k], want.edit))}
    return {"workload": wl.name, "config": {"read_len": wl.read_len, "error_rate": wl.err, "W": wl.W, "O": wl.overlap, "pairs": n},
            "value": pt["alignments_per_s_step"], "unit": "alignments/s", "kernel_alignments_per_s": pt["alignments_per_s_kernel"],
            "ms_per_step": pt["step_ms"], "kernel_ms": pt["kernel_ms"], "roofline_frac": pt["int32_frac"], "gcups": pt["gcups_kernel"],
            "e2e": e2e, "parity": pt["parity"]}


def mapping_point(G, n_reads, stress, peak, sub_batch=1_000_000, steps=2, e2e_reads=0, ncand=8):
    """configs[3]: reads x ncand candidate locations on a replicated, packed genome of G bases resident in HBM.  Returns
    the device-timed entry; with e2e_reads > 0 also the end-to-end leg through sg_set_reference + sg_align_candidates
    (host genome and host read blob in pinned memory, distances + runs out) on the first e2e_reads reads."""
    L, W = 10000, 64
    seed = synth.BASE_SEED + 4
    t0 = time.time()
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    bench_check(scrooge_b200.bench_lib().sg_synth_genome(seed, 0, G, None, p(genome), stream()))
    reads = torch.empty((n_reads, L), dtype=torch.uint8, device=dev)
    pos = torch.empty(n_reads, dtype=torch.int64, device=dev)
    bench_check(scrooge_b200.bench_lib().sg_synth_reads(seed + 1, 0, n_reads, L, 0.10, 6, 50, 54, p(genome), G, p(reads), p(pos), 1, stream()))
    pgenome, bad_g = device.pack_2bit(genome)   # one packed copy of the reference, resident in HBM
    preads, bad_r = device.pack_2bit(reads.view(-1))
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    # candidates: read-major; candidate 0 is the true start, 1..7 are the true start jittered by up to +-16 bases
    gcpu = torch.Generator(device="cpu").manual_seed(7)
    jitter = torch.randint(-16, 17, (n_reads, ncand), generator=gcpu, dtype=torch.int64)
    jitter[:, 0] = 0
    cstart2d = (pos[:, None] + jitter.to(dev)).clamp_(min=0)
    if stress:
        # SURVEY 8d's stress variant: candidate 0 = the true start, 1..7 = uniform random loci (unrelated text: window
        # distances around 32, four times the work of a true candidate in the reference's formulation)
        rnd = torch.randint(0, G - 3 * L, (n_reads, ncand), generator=gcpu, dtype=torch.int64).to(dev)
        rnd[:, 0] = cstart2d[:, 0]
        cstart2d = rnd
    cstart = cstart2d.reshape(-1).contiguous()
    cread = torch.arange(n_reads, dtype=torch.int64, device=dev).repeat_interleave(ncand)
    n = n_reads * ncand
    tlen = (G - cstart).contiguous()
    qstart = (cread * L).contiguous()
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    cap = 2 * L + 8
    sub = min(n, sub_batch)   # alignments per launch: bounds the run slab (20 KB capacity per 10 kbp alignment)
    slab_off = torch.arange(sub + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(W, sub, dev, slab_bytes=sub * cap)
    runs = torch.empty(sub * (12000 if stress else 3000), dtype=torch.uint8, device=dev)
    assert cap % 4 == 0   # SG_FLAG_RUN_WORDS: slots on 4-byte boundaries

    def timed(words):
        """The whole candidate list, `steps` times, with the kernel storing its runs as bytes (False) or as whole words (True,
        SG_FLAG_RUN_WORDS -- what sg_align_candidates launches); the first pass is untimed and keeps a sample."""
        kev, keep = [], {}

        def step(record=False):
            for b0 in range(0, n, sub):
                b1 = min(n, b0 + sub)
                assert b1 - b0 == sub, "reads x candidates must be a multiple of the sub-batch"
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                da.align(pgenome, cstart[b0:b1], tlen[b0:b1], preads, qstart[b0:b1], qlen[b0:b1], slab_off, run_words=words)
                e1.record()
                da.compact(slab_off, runs)
                if record:
                    kev.append((e0, e1))
                if b0 == 0 and not keep:
                    torch.cuda.synchronize()
                    keep["edit"] = da.out.edit[:2048].cpu().numpy().copy()
                    keep["refc"] = da.out.ref_consumed[:2048].cpu().numpy().astype(np.uint64).copy()
                    keep["ro"] = da.run_off[:2049].cpu().numpy().copy()
                    keep["runs"] = runs[: int(keep["ro"][-1])].cpu().numpy().copy()
                    keep["entries"] = int(da.out.dc_entries.sum())
                    keep["windows"] = int(da.out.windows.sum())
                    assert int(da.run_off[-1]) <= runs.numel() and int(da.out.status.max()) == 0

        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(record=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, sum(a.elapsed_time(b) for a, b in kev) / steps, keep

    # both ways of storing runs; the entry's value is the one sg_align_candidates launches (words), the other is reported beside it
    by_emit = {"bytes": timed(False), "words": timed(True)}
    emit = "words"
    ms_step, ms_kernel, keep = by_emit[emit]
    other = by_emit["words" if emit == "bytes" else "bytes"][2]
    same = all(np.array_equal(keep[k], other[k]) for k in ("edit", "refc", "ro", "runs")) and keep["entries"] == other["entries"]
    assert int(bad_g) == -1 and int(bad_r) == -1
    # parity: the first 32 reads x 8 candidates against the oracle, on a window of the genome that contains them
    k_reads = 32
    hs = cstart[: k_reads * ncand].cpu().numpy()
    lo, hi = int(hs.min()), int(hs.max()) + 3 * L
    # the oracle's text is the genome suffix; alignments only ever touch the first ~L*1.3 bases, so a suffix cut
    # 3L after the last candidate gives identical results
    hreads = [bytes(r).decode() for r in reads[:k_reads].cpu().numpy()]
    if stress:   # the candidates are spread over the genome: one 3L-base text per candidate, same cut argument
        texts = [genome[int(x):min(G, int(x) + 3 * L)].cpu().numpy().tobytes().decode() for x in hs]
        want = Oracle().align_pairs(texts, [hreads[c // ncand] for c in range(k_reads * ncand)], W=W, threads=8)
        assert int(want.ref_consumed.max()) < 3 * L - W
    else:
        gwin = genome[lo:min(G, hi)].cpu().numpy().tobytes().decode()
        want = Oracle().align_candidates(gwin, hreads, [int(x) - lo for x in hs], [c // ncand for c in range(k_reads * ncand)], threads=8)
    ok = bool(np.array_equal(keep["edit"][: k_reads * ncand], want.edit)) and bool(np.array_equal(keep["refc"][: k_reads * ncand], want.ref_consumed))
    for a in range(k_reads * ncand):
        ok = ok and runs_to_cigar(keep["runs"][keep["ro"][a]:keep["ro"][a + 1]]) == want.cigars[a]
    entries_per = keep["entries"] / sub
    out = {"workload": "mapping_10kbp_1true_7random" if stress else "mapping_10kbp_8cand", "genome_bases": G, "reads": n_reads,
           "candidates_per_read": ncand, "alignments": n, "read_len": L, "error_rate": 0.10, "W": W, "sub_batch": sub,
           "alignments_per_s_kernel": n / (ms_kernel / 1e3), "alignments_per_s_step": n / (ms_step / 1e3), "kernel_ms": ms_kernel,
           "step_ms": ms_step, "gcups_kernel": n / (ms_kernel / 1e3) * L * L / 1e9, "dc_entries_per_alignment": entries_per,
           "int32_frac": keep["windows"] / sub * n * W * COL_OPS[W] / (ms_kernel / 1e3) / 1e9 / peak,
           "reference_formulation_ratio": entries_per * n * OPS[W] / (ms_kernel / 1e3) / 1e9 / peak, "packed_genome_mb": pgenome.numel() * 4 / 1e6,
           "generate_and_pack_s": gen_s, "windows_per_alignment": keep["windows"] / sub,
           "mean_edit_distance_first_2048": float(np.mean(keep["edit"])), "true_start_mean_edit": float(np.mean(keep["edit"][0::ncand])),
           "run_emission": emit, "kernel_ms_by_run_emission": {k: v[1] for k, v in by_emit.items()},
           "parity": {"checked": k_reads * ncand, "bit_exact": ok and same, "both_run_emissions_identical_on": 2048}}
    del da, runs, slab_off, pgenome, preads
    if e2e_reads:
        # ---- end to end: host genome + host reads -> sg_set_reference (packed copy resident per GPU) + sg_align_candidates
        ne = min(e2e_reads, n_reads)
        h_genome = torch.empty(G, dtype=torch.uint8, pin_memory=True).copy_(genome)
        h_reads = torch.empty((ne, L), dtype=torch.uint8, pin_memory=True).copy_(reads[:ne])
        h_cstart = cstart[: ne * ncand].cpu().numpy().astype(np.uint64)
        h_cread = np.repeat(np.arange(ne, dtype=np.uint32), ncand)
        roff = np.arange(ne + 1, dtype=np.uint64) * np.uint64(L)
        del genome, reads
        torch.cuda.empty_cache()
        al = scrooge_b200.Aligner(W=W, device_ids=[torch.cuda.current_device()])
        t0 = time.perf_counter()
        sg_check(lib().sg_set_reference(al._h, p(h_genome), G))
        t_ref = time.perf_counter() - t0
        res, best = None, 1e9
        for _ in range(2 + steps):   # best of: the first two calls allocate the page-locked result blocks
            h = C.c_void_p()
            t0 = time.perf_counter()
            sg_check(lib().sg_align_candidates(al._h, p(h_reads), roff.ctypes.data, ne, h_cstart.ctypes.data, h_cread.ctypes.data, ne * ncand, 0, C.byref(h)))
            dt = time.perf_counter() - t0
            res = Result(h)
            best = min(best, dt)
        st = res.stats()
        ed = np.asarray(res.edit_distances)[: k_reads * ncand]
        cg = [res.cigar(a) for a in range(k_reads * ncand)]
        out["e2e"] = {"value": ne * ncand / best, "unit": "alignments/s", "reads_per_step": ne, "alignments_per_step": ne * ncand,
                      "ms_per_step": best * 1e3, "set_reference_ms": t_ref * 1e3, "set_reference_gbs": G / t_ref / 1e9,
                      "h2d_bytes_per_step": st["h2d_ascii_bytes"] + st["h2d_packed_bytes"] + st["h2d_other_bytes"],
                      "d2h_bytes_per_step": st["d2h_bytes"],
                      "api": "sg_set_reference once (not in the step) + sg_align_candidates (pinned host read blob + candidate arrays in, "
                             "distances + packed CIGAR runs out)",
                      "parity": {"checked": k_reads * ncand,
                                 "bit_exact": bool(np.array_equal(ed, want.edit)) and cg == list(want.cigars) if not stress else None}}
        al.close()
        del h_genome, h_reads
    torch.cuda.empty_cache()
    return out
