#!/usr/bin/env python
"""Device time of the ingest kernel (ASCII -> 2 bit/base) on the benchmark's text blob, and a checksum of its output.
usage: [SG_PACK=bulk] pack_time.py [pairs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import device, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
wl = synth.WORKLOADS["long_10kbp"]
dev = torch.device("cuda", 0)
text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, wl.read_len, wl.err, wl.ratio, wl.slack, dev)
flat = text.view(-1)
lib = scrooge_b200.lib()
packed, bad = device.pack_2bit(flat)
torch.cuda.synchronize()
best = 1e9
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scrooge_b200._lib.check(lib.sg_dev_pack_2bit(int(flat.data_ptr()), flat.numel(), int(packed.data_ptr()), int(bad.data_ptr()),
                                                 int(torch.cuda.current_stream().cuda_stream)))
    e1.record()
    torch.cuda.synchronize()
    if it:
        best = min(best, e0.elapsed_time(e1))
nbytes = flat.numel() * 1.25
chk = int(packed.view(torch.int32).to(torch.int64).sum().item())
# an odd-sized blob with a tail and a bad base near the end: the reported position must not depend on the kernel
odd = flat[: 5 * 16384 + 777].clone()
odd[5 * 16384 + 300] = ord("N")
p2, b2 = device.pack_2bit(odd)
torch.cuda.synchronize()
print(f"SG_PACK={os.environ.get('SG_PACK', 'bulk (default)')}: {best:.3f} ms for {flat.numel() / 1e9:.2f} G bases = {nbytes / best / 1e9:.3f} TB/s, "
      f"checksum {chk}, bad position {int(b2.item())}, odd checksum {int(p2.view(torch.int32).to(torch.int64).sum().item())}", flush=True)
