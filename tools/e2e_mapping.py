#!/usr/bin/env python
"""End-to-end throughput of the read-mapping interface (BASELINE.json configs[3]) through the C ABI with HOST buffers:
sg_set_reference (3 Gbp of ASCII -> packed genome resident in HBM) + sg_align_candidates (reads blob + candidate
locations in, distances + packed CIGAR runs out).  The first reads' candidates are checked against the oracle.
usage: e2e_mapping.py [genome_bases] [reads]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from oracle.binding import Oracle  # noqa: E402  (checker only: never timed)
from scrooge_b200 import synth  # noqa: E402
from scrooge_b200._lib import check  # noqa: E402
from scrooge_b200.api import Result  # noqa: E402

G = int(float(sys.argv[1])) if len(sys.argv) > 1 else 3_000_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
L, ncand, W = 10000, 8, 64
lib = scrooge_b200.lib()
dev = torch.device("cuda:0")
p = lambda t: int(t.data_ptr())
st = int(torch.cuda.current_stream().cuda_stream)
seed = synth.BASE_SEED + 4
t0 = time.time()
genome = torch.empty(G, dtype=torch.uint8, device=dev)
check(scrooge_b200.bench_lib().sg_synth_genome(seed, 0, G, None, p(genome), st))
reads = torch.empty((n_reads, L), dtype=torch.uint8, device=dev)
pos = torch.empty(n_reads, dtype=torch.int64, device=dev)
check(scrooge_b200.bench_lib().sg_synth_reads(seed + 1, 0, n_reads, L, 0.10, 6, 50, 54, p(genome), G, p(reads), p(pos), 1, st))
h_genome = torch.empty(G, dtype=torch.uint8, pin_memory=True).copy_(genome)
h_reads = torch.empty((n_reads, L), dtype=torch.uint8, pin_memory=True).copy_(reads)
gcpu = torch.Generator(device="cpu").manual_seed(7)
jitter = torch.randint(-16, 17, (n_reads, ncand), generator=gcpu, dtype=torch.int64)
jitter[:, 0] = 0
cstart = (pos.cpu()[:, None] + jitter).clamp_(min=0).reshape(-1).numpy().astype(np.uint64)
cread = np.repeat(np.arange(n_reads, dtype=np.uint32), ncand)
roff = (np.arange(n_reads + 1, dtype=np.uint64) * L)
del genome, reads
torch.cuda.empty_cache()
n = n_reads * ncand
print(f"inputs ready in {time.time() - t0:.1f} s: genome {G / 1e9:.1f} Gbp, {n_reads} reads x {ncand} candidates", flush=True)

al = scrooge_b200.Aligner(W=W, device_ids=[0])
t0 = time.perf_counter()
check(lib.sg_set_reference(al._h, p(h_genome), G))
t_ref = time.perf_counter() - t0
print(f"sg_set_reference: {t_ref * 1e3:.0f} ms ({G / t_ref / 1e9:.1f} GB/s of ASCII)", flush=True)
best, res = 1e9, None
for rep in range(3):
    out = C.c_void_p()
    t0 = time.perf_counter()
    check(lib.sg_align_candidates(al._h, p(h_reads), roff.ctypes.data, n_reads, cstart.ctypes.data, cread.ctypes.data, n, 0, C.byref(out)))
    dt = time.perf_counter() - t0
    res = Result(out)
    print(f"sg_align_candidates: {dt * 1e3:.0f} ms, kernel {res.kernel_ns / 1e6:.0f} ms", flush=True)
    best = min(best, dt)
k_reads = 16
hs = cstart[: k_reads * ncand].astype(np.int64)
lo, hi = int(hs.min()), int(hs.max()) + 3 * L
gwin = h_genome[lo:min(G, hi)].numpy().tobytes().decode()
hreads = [bytes(r).decode() for r in h_reads[:k_reads].numpy()]
want = Oracle().align_candidates(gwin, hreads, [int(x) - lo for x in hs], [c // ncand for c in range(k_reads * ncand)], threads=8)
ed = res.edit_distances[: k_reads * ncand]
cg = [res.cigar(a) for a in range(k_reads * ncand)]
ok = bool(np.array_equal(ed, want.edit)) and cg == list(want.cigars)
print(f"mapping e2e: {n / best / 1e6:.2f} M alignments/s ({best * 1e3:.0f} ms for {n} alignments), d2h runs {int(res.run_offsets[-1]) / 1e9:.1f} GB, "
      f"parity on {k_reads * ncand} candidates: {'bit-exact' if ok else 'MISMATCH'}", flush=True)
