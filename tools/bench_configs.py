#!/usr/bin/env python
"""Device-timed measurements of BASELINE.json's non-headline configurations on one GPU (the headline, configs[2], is
bench.py).  Every point is checked against the oracle on a sample of its own inputs.

    python tools/bench_configs.py sweep    [--bases 1e10]   configs[4]: L x error x {distance-only, full CIGAR}
    python tools/bench_configs.py mapping  [--genome 3e9]   configs[3]: reads x 8 candidate locations on a replicated genome
    python tools/bench_configs.py short                     configs[1]: 10M x 150 bp pairs, W=64/O=33 and W=32/O=17
    python tools/bench_configs.py windows  [--pairs 200000] the reference's window sweep (scripts/profile.py:66-100: W with
                                                            O = min(W//2+1, W-1), and O at W=64) on 10 kbp / 10 % pairs

One JSON object per line on stdout.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker only: never timed)
from scrooge_b200 import device, synth  # noqa: E402

class _PerWord(dict):
    def __init__(self, per_word):
        super().__init__()
        self.per_word = per_word

    def __missing__(self, W):
        return self.per_word * ((W + 31) // 32)


OPS = _PerWord(7)          # reference formulation: INT32 ops per R[d][i] entry: 7 per 32-bit word
COL_OPS = _PerWord(10)     # delta kernel: INT32 ops per window column (see bench.py): 10 per 32-bit word
lib = scrooge_b200.lib()
dev = torch.device("cuda:0")
p = lambda t: int(t.data_ptr())
stream = lambda: int(torch.cuda.current_stream().cuda_stream)


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
    cap = 2 * L + 8
    slab_off = None if distance_only else torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(W, n, dev, slab_bytes=0 if distance_only else n * cap, O=O)
    ptext, bad_t = device.pack_2bit(text.view(-1))
    pquery, bad_q = device.pack_2bit(reads.view(-1))

    def kernel():
        da.align(ptext, tstart, tlen, pquery, qstart, qlen, slab_off, distance_only=distance_only)

    kernel()
    runs = None
    if not distance_only:
        runs = da.compact(slab_off)[1]

    def step():
        scrooge_b200._lib.check(lib.sg_dev_pack_2bit(p(text), n * stride, p(ptext), p(bad_t), stream()))
        scrooge_b200._lib.check(lib.sg_dev_pack_2bit(p(reads), n * L, p(pquery), p(bad_q), stream()))
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


def cmd_sweep(args, peak):
    for L in (1000, 2000, 5000, 10000, 20000, 50000, 100000):
        n = max(1024, int(args.bases // L))
        for e in (0.05, 0.10, 0.15):
            for dist in (True, False):
                wl = synth.Workload(f"sweep_{L}bp_{int(e * 100)}pct", L, e, synth.PACBIO, 64, synth.BASE_SEED + 5)
                print(json.dumps(pairs_point(wl, n, dist, peak, check=max(8, min(256, 2_000_000 // L)))), flush=True)


def cmd_windows(args, peak):
    """The axes of the reference's window sweep (scripts/profile.py:66-100 cpu_sweep_wo / cpu_sweep_o; :595-640): W with the
    overlap the reference pairs with it, and the overlap at W = 64.  64/33 and 32/17 run on the tuned kernels, everything
    else on genasm_generic_kernel (64/33 also once on the generic kernel when SG_GENERIC=1 is set by the caller)."""
    n = args.pairs
    points = [(16, 9), (32, 17), (48, 25), (64, 33), (96, 49), (128, 65), (160, 81), (192, 97), (224, 113), (256, 129)] + \
             [(64, O) for O in (0, 1, 8, 16, 24, 40, 48, 56, 63)]
    if args.only:
        points = [tuple(int(x) for x in p.split('/')) for p in args.only.split(',')]
    for W, O in points:
        wl = synth.Workload(f"long_10kbp_w{W}_o{O}", 10000, 0.10, synth.PACBIO, W, synth.BASE_SEED + 3, O=O)
        print(json.dumps(pairs_point(wl, n, False, peak, check=64)), flush=True)
    if not args.only:
        wl = synth.Workload("short_150bp_w48_o25", 150, 0.05, synth.ILLUMINA, 48, synth.BASE_SEED + 2, O=25)
        print(json.dumps(pairs_point(wl, 2_000_000, False, peak, check=2048)), flush=True)


def cmd_short(args, peak):
    for name in ("short_150bp", "short_150bp_w32"):
        for dist in (True, False):
            print(json.dumps(pairs_point(synth.WORKLOADS[name], 10_000_000, dist, peak, check=4096)), flush=True)


def cmd_mapping(args, peak):
    G, n_reads, L, ncand, W = int(args.genome), args.reads, 10000, 8, 64
    seed = synth.BASE_SEED + 4
    t0 = time.time()
    genome = torch.empty(G, dtype=torch.uint8, device=dev)
    scrooge_b200._lib.check(scrooge_b200.bench_lib().sg_synth_genome(seed, 0, G, None, p(genome), stream()))
    reads = torch.empty((n_reads, L), dtype=torch.uint8, device=dev)
    pos = torch.empty(n_reads, dtype=torch.int64, device=dev)
    scrooge_b200._lib.check(scrooge_b200.bench_lib().sg_synth_reads(seed + 1, 0, n_reads, L, 0.10, 6, 50, 54, p(genome), G, p(reads), p(pos), 1, stream()))
    pgenome, bad_g = device.pack_2bit(genome)   # one packed copy of the reference, resident in HBM
    preads, bad_r = device.pack_2bit(reads.view(-1))
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    # candidates: read-major; candidate 0 is the true start, 1..7 are the true start jittered by up to +-16 bases
    gcpu = torch.Generator(device="cpu").manual_seed(7)
    jitter = torch.randint(-16, 17, (n_reads, ncand), generator=gcpu, dtype=torch.int64)
    jitter[:, 0] = 0
    cstart2d = (pos[:, None] + jitter.to(dev)).clamp_(min=0)
    if args.stress:
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
    sub = min(n, args.sub_batch)   # alignments per launch: bounds the run slab (20 KB capacity per 10 kbp alignment)
    slab_off = torch.arange(sub + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(W, sub, dev, slab_bytes=sub * cap)
    runs = torch.empty(sub * (12000 if args.stress else 3000), dtype=torch.uint8, device=dev)
    kev = []
    keep = {}

    def step(record=False):
        for b0 in range(0, n, sub):
            b1 = min(n, b0 + sub)
            assert b1 - b0 == sub, "reads x candidates must be a multiple of the sub-batch"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            da.align(pgenome, cstart[b0:b1], tlen[b0:b1], preads, qstart[b0:b1], qlen[b0:b1], slab_off)
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
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    assert int(bad_g) == -1 and int(bad_r) == -1
    # parity: the first 32 reads x 8 candidates against the oracle, on a window of the genome that contains them
    k_reads = 32
    hp = pos[:k_reads].cpu().numpy()
    hs = cstart[: k_reads * ncand].cpu().numpy()
    lo, hi = int(hs.min()), int(hs.max()) + 3 * L
    # the oracle's text is the genome suffix; alignments only ever touch the first ~L*1.3 bases, so a suffix cut
    # 3L after the last candidate gives identical results
    hreads = [bytes(r).decode() for r in reads[:k_reads].cpu().numpy()]
    if args.stress:   # the candidates are spread over the genome: one 3L-base text per candidate, same cut argument
        texts = [genome[int(x):min(G, int(x) + 3 * L)].cpu().numpy().tobytes().decode() for x in hs]
        want = Oracle().align_pairs(texts, [hreads[c // ncand] for c in range(k_reads * ncand)], W=W, threads=8)
        assert int(want.ref_consumed.max()) < 3 * L - W
    else:
        gwin = genome[lo:min(G, hi)].cpu().numpy().tobytes().decode()
        want = Oracle().align_candidates(gwin, hreads, [int(x) - lo for x in hs], [c // ncand for c in range(k_reads * ncand)], threads=8)
    ok = bool(np.array_equal(keep["edit"][: k_reads * ncand], want.edit)) and bool(np.array_equal(keep["refc"][: k_reads * ncand], want.ref_consumed))
    for a in range(k_reads * ncand):
        s = "".join(f"{int(b) & 63}{'=XID'[int(b) >> 6]}" for b in keep["runs"][keep["ro"][a]:keep["ro"][a + 1]])
        ok = ok and s == want.cigars[a]
    entries_per = keep["entries"] / sub
    print(json.dumps({"workload": "mapping_10kbp_1true_7random" if args.stress else "mapping_10kbp_8cand", "genome_bases": G, "reads": n_reads, "candidates_per_read": ncand, "alignments": n,
                      "read_len": L, "error_rate": 0.10, "W": W, "sub_batch": sub, "alignments_per_s_kernel": n / (ms_kernel / 1e3),
                      "alignments_per_s_step": n / (ms_step / 1e3), "kernel_ms": ms_kernel, "step_ms": ms_step,
                      "gcups_kernel": n / (ms_kernel / 1e3) * L * L / 1e9, "dc_entries_per_alignment": entries_per,
                      "int32_frac": keep["windows"] / sub * n * W * COL_OPS[W] / (ms_kernel / 1e3) / 1e9 / peak,
                      "reference_formulation_ratio": entries_per * n * OPS[W] / (ms_kernel / 1e3) / 1e9 / peak, "packed_genome_mb": pgenome.numel() * 4 / 1e6,
                      "generate_and_pack_s": gen_s, "windows_per_alignment": keep["windows"] / sub,
                      "mean_edit_distance_first_2048": float(np.mean(keep["edit"])), "true_start_mean_edit": float(np.mean(keep["edit"][0::ncand])),
                      "parity": {"checked": k_reads * ncand, "bit_exact": ok}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["sweep", "mapping", "short", "windows"])
    ap.add_argument("--pairs", type=int, default=200_000)
    ap.add_argument("--only", default="", help="windows: comma-separated W/O points instead of the default list")
    ap.add_argument("--bases", type=float, default=1e10)
    ap.add_argument("--genome", type=float, default=3e9)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--sub-batch", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--stress", action="store_true", help="mapping: 1 true + 7 uniform-random candidate loci per read")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak = device.int32_peak(2, 60.0)
    {"sweep": cmd_sweep, "mapping": cmd_mapping, "short": cmd_short, "windows": cmd_windows}[args.what](args, peak)


if __name__ == "__main__":
    main()
