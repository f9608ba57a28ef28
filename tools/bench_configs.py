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
from bench_extra import mapping_point, pairs_point  # noqa: E402  (the measurement legs shared with bench.py)
from scrooge_b200 import device, synth  # noqa: E402

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
    print(json.dumps(mapping_point(int(args.genome), args.reads, args.stress, peak, sub_batch=args.sub_batch, steps=args.steps,
                                   e2e_reads=args.e2e_reads)), flush=True)


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
    ap.add_argument("--e2e-reads", type=int, default=0, help="mapping: also run the first N reads end to end through the host C ABI")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak = device.int32_peak(2, 60.0)
    {"sweep": cmd_sweep, "mapping": cmd_mapping, "short": cmd_short, "windows": cmd_windows}[args.what](args, peak)


if __name__ == "__main__":
    main()
