#!/usr/bin/env python
"""Device time of the alignment kernel alone on the benchmark workload, for a few batch sizes.
usage: kernel_time.py [workload] [pairs,pairs,...]   (SG_LIB=path selects another build of the library)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import device, synth  # noqa: E402

wl = synth.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "long_10kbp"]
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "104192,312576,1000000").split(",")]
dev = torch.device("cuda", 0)
lib = scrooge_b200.lib()
for n in sizes:
    L = wl.read_len
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, L, wl.err, wl.ratio, wl.slack, dev)
    stride = text.shape[1]
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    tstart, qstart = idx * stride, idx * L
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    cap = 2 * L + 8
    slab_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(wl.W, n, dev, slab_bytes=n * cap)
    ptext, _ = device.pack_2bit(text.view(-1))
    pquery, _ = device.pack_2bit(reads.view(-1))
    del text, reads
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        da.align(ptext, tstart, tlen, pquery, qstart, qlen, slab_off)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    print(f"{os.environ.get('SG_LIB', 'default')}: {wl.name} n={n}: {best:.3f} ms  {n / best / 1e3:.3f} M aligns/s", flush=True)
    del da, ptext, pquery, slab_off
    torch.cuda.empty_cache()
