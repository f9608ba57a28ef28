#!/usr/bin/env python
"""End-to-end throughput of sg_align_pairs (pinned host ASCII blobs in, distances + runs out) for a named workload.
usage: e2e_workloads.py <workload> <pairs> ; SG_DEBUG=1 prints the host-time breakdown of every call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import synth  # noqa: E402

wl = synth.WORKLOADS[sys.argv[1]]
n = int(sys.argv[2])
t0 = time.perf_counter()
text, tlen, reads = synth.pairs_host(wl, 0, n)
tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
del text
tb_pin = torch.from_numpy(tb).pin_memory()
qb_pin = torch.from_numpy(qb).pin_memory()
print(f"inputs ready in {time.perf_counter() - t0:.1f} s: {tb.nbytes / 1e9:.2f} + {qb.nbytes / 1e9:.2f} GB", flush=True)
al = scrooge_b200.Aligner(W=wl.W, device_ids=[0])
al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    r = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
    best = min(best, time.perf_counter() - t0)
print(f"{wl.name} n={n}: {n / best / 1e6:.3f} M aligns/s  ({best * 1e3:.1f} ms, kernel {r.kernel_ns / 1e6:.1f} ms)", flush=True)
