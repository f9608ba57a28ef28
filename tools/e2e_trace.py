#!/usr/bin/env python
"""One end-to-end call with SG_TRACE=1: the pipeline's event log (stderr) for a named workload.
usage: SG_TRACE=1 python tools/e2e_trace.py [pairs] [workload]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
wl = synth.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "long_10kbp"]
text, tlen, reads = synth.pairs_host(wl, 0, n)
tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
del text, reads
tb_pin, qb_pin = torch.from_numpy(tb).pin_memory(), torch.from_numpy(qb).pin_memory()
al = scrooge_b200.Aligner(W=wl.W, n_gpus=int(os.environ.get("GPUS", "1")))
for k in range(4):
    sys.stderr.write(f"==== call {k}\n")
    sys.stderr.flush()
    t0 = time.perf_counter()
    r = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff)
    dt = time.perf_counter() - t0
    print(f"call {k}: {dt * 1e3:.1f} ms, {n / dt / 1e6:.2f} M alignments/s, stats {r.stats()}", flush=True)
    del r
