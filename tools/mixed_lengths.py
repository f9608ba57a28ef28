#!/usr/bin/env python
"""A batch of MIXED read lengths in one launch: 1 kbp, 10 kbp and 100 kbp pairs with the same number
of bases per class, shuffled.  Kernel time of the mixed launch in input (random) order and handed out longest first
(sg_dev_align_ordered), against the bases-weighted ideal = the sum of the three homogeneous launches' times.
usage: python tools/mixed_lengths.py [bases_per_class]     one JSON line on stdout"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import device, synth  # noqa: E402
from scrooge_b200._lib import check  # noqa: E402

B = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000_000
lib = scrooge_b200.lib()
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
p = lambda t: int(t.data_ptr()) if t is not None else None
st = lambda: int(torch.cuda.current_stream().cuda_stream)


def launch(pt, ts, tl, pq, qs, ql, order=None, reps=3):
    n = ql.numel()
    cap = 2 * ql + 8
    slab_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cap, 0)])
    da = device.DeviceAligner(64, n, dev, slab_bytes=int(slab_off[-1]))
    o = da.out
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.sg_dev_align_ordered(64, 33, p(pt), p(ts), p(tl), p(pq), p(qs), p(ql), n, 0, p(da.slab), p(slab_off), p(da.counter),
                                       p(o.edit), p(o.ref_consumed), p(o.nruns), p(o.status), None, None, p(order), st()))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert int(o.status.max()) == 0
    return best, o.edit.clone()


classes = []
off_t = off_q = 0
packed_t, packed_q, TS, TL, QS, QL = [], [], [], [], [], []
homog = {}
for L in (1000, 10000, 100000):
    n = B // L
    text, tlen, reads = device.synth_pairs_device(synth.BASE_SEED + 9, 0, n, L, 0.10, synth.PACBIO, 64, dev)
    stride = text.shape[1]
    pt, _ = device.pack_2bit(text.view(-1))
    pq, _ = device.pack_2bit(reads.view(-1))
    del text, reads
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    ts, qs, ql = idx * stride, idx * L, torch.full((n,), L, dtype=torch.int64, device=dev)
    ms, edit = launch(pt, ts, tlen, pq, qs, ql)
    homog[L] = {"pairs": n, "kernel_ms": ms, "edit_sum": int(edit.sum())}
    # into the mixed batch: packed blobs are whole words, so a class starts at a multiple of 16 bases
    packed_t.append(pt); packed_q.append(pq)
    TS.append(ts + off_t * 16); TL.append(tlen); QS.append(qs + off_q * 16); QL.append(ql)
    off_t += pt.numel(); off_q += pq.numel()
pt, pq = torch.cat(packed_t), torch.cat(packed_q)
del packed_t, packed_q
ts, tl, qs, ql = torch.cat(TS), torch.cat(TL), torch.cat(QS), torch.cat(QL)
n = ql.numel()
perm = torch.randperm(n, generator=torch.Generator(device="cpu").manual_seed(5)).to(dev)
ts, tl, qs, ql = ts[perm].contiguous(), tl[perm].contiguous(), qs[perm].contiguous(), ql[perm].contiguous()
ms_rand, e_rand = launch(pt, ts, tl, pq, qs, ql)
order = torch.argsort(ql, descending=True, stable=True).to(torch.int32).contiguous()
ms_sorted, e_sorted = launch(pt, ts, tl, pq, qs, ql, order=order)
assert torch.equal(e_rand, e_sorted), "the launch order changed results"
assert int(e_rand.sum()) == sum(h["edit_sum"] for h in homog.values()), "the mixed batch differs from its parts"
ideal = sum(h["kernel_ms"] for h in homog.values())
print(json.dumps({"bases_per_class": B, "classes": homog, "mixed_pairs": n, "ideal_ms_sum_of_homogeneous_launches": ideal,
                  "mixed_input_order_ms": ms_rand, "mixed_longest_first_ms": ms_sorted,
                  "input_order_over_ideal": ms_rand / ideal, "longest_first_over_ideal": ms_sorted / ideal}))
