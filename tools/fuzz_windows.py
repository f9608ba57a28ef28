#!/usr/bin/env python
"""Randomised differential run of the run-time window configurations: random (W, O) within the supported range, random mixed
pairs (empty / tiny / W+-1 lengths, exhausted texts, unrelated pairs), CUDA path through the C ABI against the oracle.
usage: fuzz_windows.py [configs] [pairs per config] [seed]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scrooge_b200  # noqa: E402
from conftest import random_pairs  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker)

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 600
rng = random.Random(int(sys.argv[3]) if len(sys.argv) > 3 else 2026)
o = Oracle()
bad = 0
for c in range(n_cfg):
    W = rng.choice([rng.randint(2, 256), rng.choice([31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 255, 256])])
    O = rng.randint(max(0, W - 128), W - 1)
    T, Q = random_pairs(rng.randrange(1 << 30), n_pairs, [0, 1, 2, 3, W - 1, W, W + 1, 2 * W + 1, 3 * W, 150, 400, 1500],
                        [0.0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6])
    want = o.align_pairs(T, Q, W=W, O=O, threads=8)
    al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
    got = al.align_pairs(T, Q)
    cg = got.cigars()
    ok = list(got.edit_distances) == list(want.edit) and cg == want.cigars and list(got.ref_consumed) == list(want.ref_consumed)
    d = al.align_pairs(T, Q, distance_only=True)
    ok = ok and list(d.edit_distances) == list(want.edit)
    al.close()
    bad += 0 if ok else 1
    print(f"W={W} O={O} pairs={n_pairs} {'ok' if ok else 'MISMATCH'}", flush=True)
print(f"{n_cfg} configurations, {bad} with mismatches")
sys.exit(1 if bad else 0)
