"""Randomised campaign through the host simulation of the kernels (tests/sim): mixed pairs (empty / tiny / W+-1 lengths,
exhausted texts, unrelated pairs, 0-80 % error) through genasm_delta_kernel at both tuned configurations and every run-emission
variant, and through genasm_generic_kernel at random window configurations, each launch compared with the oracle bit for bit.
No GPU needed.   python tools/sim_fuzz.py [rounds] [pairs per launch]"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ctypes as C  # noqa: E402

import test_kernel_sim as tks  # noqa: E402
from conftest import random_pairs  # noqa: E402
from oracle.binding import Oracle, build  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    build()
    sim = tks.sim.__wrapped__()   # the fixture's body: builds tests/sim/_build/libsgsim.so when stale
    oracle = Oracle()
    rng = random.Random(20261017)
    t0, pairs = time.time(), 0
    for r in range(rounds):
        seed = rng.randrange(1 << 30)
        for W in (64, 32):
            T, Q = random_pairs(seed + W, per, [0, 1, 2, 3, W - 1, W, W + 1, 2 * W + 1, 100, 150, 400, 1500, 4000], [0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6, 0.8])
            want = oracle.align_pairs(T, Q, W=W, threads=8)
            for emit in (0, 1):
                out = tks.run_sim(sim, W, emit, T, Q, ctas=rng.choice([1, 2, 3]))
                tks.check(out, want, len(T))
                pairs += len(T)
        W = rng.choice([rng.randint(2, 256), rng.choice([31, 33, 63, 65, 96, 127, 128, 129, 255, 256])])
        O = rng.randint(max(0, W - 128), W - 1)
        T, Q = random_pairs(seed, per // 2, [0, 1, 2, 3, W - 1, W, W + 1, 2 * W + 1, 150, 400, 1500], [0, 0.02, 0.05, 0.1, 0.15, 0.3, 0.6])
        want = oracle.align_pairs(T, Q, W=W, O=O, threads=8)
        for gp in (0, 1):
            out = tks.run_sim(sim, W, 0, T, Q, ctas=2, generic=(O, gp))
            tks.check(out, want, len(T))
            pairs += len(T)
        print(f"round {r + 1}/{rounds}: generic ({W}, {O}) ok; {pairs} simulated alignments checked so far, {time.time() - t0:.0f} s", flush=True)
    print(f"sim fuzz clean: {pairs} simulated alignments bit-exact against the oracle in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
