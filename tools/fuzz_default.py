#!/usr/bin/env python
"""Randomised differential run of the two tuned configurations (64/33, 32/17): many mixed random pairs, CUDA path through the
C ABI (blob and vectored entry points) against the oracle.  usage: fuzz_default.py [pairs] [seed]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scrooge_b200  # noqa: E402
from conftest import random_pairs  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 99
o = Oracle()
bad = 0
for W in (64, 32):
    T, Q = random_pairs(seed + W, n, [0, 1, 2, 3, 5, 15, 16, 17, 30, 31, 32, 33, 34, 62, 63, 64, 65, 66, 95, 97, 127, 128, 129, 150, 250, 400,
                                      1000, 2500], [0.0, 0.01, 0.02, 0.05, 0.1, 0.15, 0.2, 0.3, 0.45, 0.6, 0.9])
    want = o.align_pairs(T, Q, W=W, threads=16)
    al = scrooge_b200.Aligner(W=W, n_gpus=1)
    for name, got in (("blob", al.align_pairs(T, Q)), ("vectored", al.align_pairs_v(T, Q))):
        ok = list(got.edit_distances) == list(want.edit) and got.cigars() == want.cigars and list(got.ref_consumed) == list(want.ref_consumed)
        bad += 0 if ok else 1
        print(f"W={W} {name}: {n} pairs {'ok' if ok else 'MISMATCH'}", flush=True)
    al.close()
sys.exit(1 if bad else 0)
