import os, sys, ctypes as C
sys.path.insert(0, '.')
import torch, numpy as np
import scrooge_b200
from scrooge_b200 import device, synth
wl = synth.WORKLOADS["long_10kbp"]; dev = torch.device("cuda", 0); lib = scrooge_b200.lib()
for n in (312576, 1000000):
    L = wl.read_len
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, L, wl.err, wl.ratio, wl.slack, dev)
    stride = text.shape[1]; idx = torch.arange(n, dtype=torch.int64, device=dev)
    tstart, qstart = idx * stride, idx * L; qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    cap = 2 * L + 8; slab_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    da = device.DeviceAligner(wl.W, n, dev, slab_bytes=n * cap)
    ptext, _ = device.pack_2bit(text.view(-1)); pquery, _ = device.pack_2bit(reads.view(-1))
    del text, reads
    out = np.zeros(4, dtype=np.uint64)
    lib.sg_dev_debug_stats(C.c_void_p(out.ctypes.data), 1)
    da.align(ptext, tstart, tlen, pquery, qstart, qlen, slab_off); torch.cuda.synchronize()
    lib.sg_dev_debug_stats(C.c_void_p(out.ctypes.data), 1)
    print(n, out, "generic DC frac", out[1] / (out[0] + out[1]), "generic TB frac", out[3] / (out[2] + out[3]))
    del da, ptext, pquery; torch.cuda.empty_cache()
