#!/usr/bin/env python
"""End-to-end throughput of sg_align_pairs (pinned host ASCII blobs in, distances + runs out) under different host-side
pipeline settings.  usage: e2e_sweep.py [pairs] ; settings are the SG_* environment knobs read by sg_ctx_create."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
wl = synth.WORKLOADS["long_10kbp"]
text, tlen, reads = synth.pairs_host(wl, 0, n)
tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
del text
tb_pin = torch.from_numpy(tb).pin_memory()
qb_pin = torch.from_numpy(qb).pin_memory()
SETTINGS = [dict(), dict(SG_INGEST="fixed"), dict(SG_CHUNK_KB="2048"), dict(SG_CHUNK_KB="32768"), dict(SG_DMA_DEPTH="2"),
            dict(SG_DMA_DEPTH="8"), dict(SG_CHUNK_KB="4096", SG_DMA_DEPTH="8"), dict(SG_SLOTS="4"),
            dict(SG_HOST_THREADS="15"), dict(SG_HOST_THREADS="8"), dict(SG_HOST_THREADS="2"), dict(SG_HOST_PACK="0")]
KEYS = sorted({k for s in SETTINGS for k in s})
for st in SETTINGS:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(st)
    al = scrooge_b200.Aligner(W=wl.W, device_ids=[0])
    al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        r = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
        best = min(best, time.perf_counter() - t0)
    print(f"{st}  {n / best / 1e6:.3f} M aligns/s  ({best * 1e3:.1f} ms, kernel {r.kernel_ns / 1e6:.1f} ms)", flush=True)
    al.close()
