import os, sys, time
sys.path.insert(0, '.')
import torch, scrooge_b200
from scrooge_b200 import synth
n = 524288
wl = synth.WORKLOADS["long_10kbp"]
text, tlen, reads = synth.pairs_host(wl, 0, n)
tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
del text
tb_pin = torch.from_numpy(tb).pin_memory(); qb_pin = torch.from_numpy(qb).pin_memory()
al = scrooge_b200.Aligner(W=64, device_ids=[0])
for _ in range(3):
    t0 = time.perf_counter(); r = al.align_pairs_blob(tb_pin, toff, qb_pin, qoff); print("py", (time.perf_counter() - t0) * 1e3, r.total_ns / 1e6)
