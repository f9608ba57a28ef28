"""GPU parity of the run-emission variants of the tuned kernels (SG_FLAG_RUN_WORDS: runs stored as whole 32-bit words, what
the host API launches; bytes: the device API's default) against the oracle and against each other, through the device API
and through the host API with SG_EMIT forcing either variant.  The same kernel source is checked against the
oracle on the CPU-only box by tests/test_kernel_sim.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, mapping_case, random_pairs

pytestmark = pytest.mark.gpu
N_PAIRS = 8192   # per device-API case (tests/test_bench_logic_sim.py runs the same test body on the simulated device with fewer)


def _cigars(runs, run_off, idxs):
    return ["".join(f"{int(b) & 63}{'=XID'[int(b) >> 6]}" for b in runs[run_off[k]:run_off[k + 1]]) for k in idxs]


@pytest.mark.parametrize("W,L,err", [(64, 1000, 0.10), (32, 150, 0.05), (64, 777, 0.45)])
def test_device_api_run_words_equals_bytes_and_oracle(oracle, sglib, W, L, err):
    """Slots on 4-byte boundaries, runs stored as words: same distances, consumed prefixes, run counts and run bytes as the
    byte stores, and as the oracle.  L = 777 with 45 % error: capacities that are not multiples of 4 before rounding, windows
    that are mostly edits (many runs per window: words complete inside the run loop) and slots that end on 1..3 pending runs."""
    import torch
    from scrooge_b200 import device, synth
    wl = synth.Workload("t", L, err, synth.PACBIO, W, 777 + L)
    n = N_PAIRS
    dev = torch.device("cuda:0")
    text, tlen, reads = device.synth_pairs_device(wl.seed, 0, n, wl.read_len, wl.err, wl.ratio, wl.slack, dev)
    stride = text.shape[1]
    ptext, _ = device.pack_2bit(text.view(-1))
    pquery, _ = device.pack_2bit(reads.view(-1))
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    qlen = torch.full((n,), L, dtype=torch.int64, device=dev)
    cap = (2 * L + 8 + 3) & ~3
    slab_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * cap
    got = {}
    for words in (False, True):
        da = device.DeviceAligner(W, n, dev, slab_bytes=n * cap)
        da.slab.fill_(0xEE)
        out = da.align(ptext, idx * stride, tlen, pquery, idx * L, qlen, slab_off, run_words=words)
        run_off, runs = da.compact(slab_off)
        torch.cuda.synchronize()
        assert int(out.status.max().item()) == 0
        assert device.check_runs(runs, run_off, qlen, out, W) == 0
        got[words] = (out.edit.cpu().numpy().copy(), out.ref_consumed.cpu().numpy().copy(), out.nruns.cpu().numpy().copy(),
                      run_off.cpu().numpy().copy(), runs.cpu().numpy().copy(), int(out.dc_entries.sum().item()))
    for a, b in zip(got[False], got[True]):
        assert np.array_equal(a, b)
    h_text, h_tlen, h_reads = synth.pairs_host(wl, 0, 1024)
    T, Q = synth.pairs_as_strings(h_text, h_tlen, h_reads)
    want = oracle.align_pairs(T, Q, W=W, threads=4)
    ed, rc, _, ro, rr, _ = got[True]
    assert np.array_equal(ed[:1024], want.edit) and np.array_equal(rc[:1024].astype(np.uint64), want.ref_consumed)
    assert _cigars(rr, ro, range(1024)) == list(want.cigars)


def test_device_api_run_words_needs_aligned_slab(sglib):
    import torch
    from scrooge_b200 import ScroogeError, device
    dev = torch.device("cuda:0")
    n = 4
    z = torch.zeros(64, dtype=torch.int32, device=dev)
    one = torch.ones(n, dtype=torch.int64, device=dev)
    da = device.DeviceAligner(64, n, dev, slab_bytes=1024)
    da.slab = da.slab[1:]           # a slab that starts on an odd address
    with pytest.raises(ScroogeError):
        da.align(z, one * 0, one, z, one * 0, one, torch.arange(n + 1, dtype=torch.int64, device=dev) * 16, run_words=True)


_CHILD = r"""
import json, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import mapping_case, random_pairs
import scrooge_b200
T, Q = random_pairs(5150, 3000, [0, 1, 31, 32, 33, 64, 65, 150, 400, 1500], [0, 0.05, 0.15, 0.4, 0.8])
al = scrooge_b200.Aligner(W=int(sys.argv[2]))
res = al.align_pairs(T, Q)
out = {"pairs": {"edit": [int(x) for x in res.edit_distances], "refc": [int(x) for x in res.ref_consumed], "cigars": res.cigars()}}
genome, reads, cs, cr = mapping_case(99)
al.set_reference(genome)
res = al.align_candidates(reads, cs, cr)
out["mapping"] = {"edit": [int(x) for x in res.edit_distances], "refc": [int(x) for x in res.ref_consumed], "cigars": res.cigars()}
print(json.dumps(out))
"""


@pytest.mark.parametrize("W", [64, 32])
def test_host_api_emit_policies_agree(oracle, W):
    """sg_align_pairs / sg_align_candidates with SG_EMIT=bytes, SG_EMIT=words and the default (words): identical results,
    equal to the oracle's."""
    outs = {}
    for emit in ("bytes", "words", None):
        env = dict(os.environ)
        env.pop("SG_EMIT", None)
        if emit:
            env["SG_EMIT"] = emit
        r = subprocess.run([sys.executable, "-c", _CHILD, ROOT, str(W)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        import json
        outs[emit] = json.loads(r.stdout.strip().splitlines()[-1])
    assert outs["bytes"] == outs["words"] == outs[None]
    T, Q = random_pairs(5150, 3000, [0, 1, 31, 32, 33, 64, 65, 150, 400, 1500], [0, 0.05, 0.15, 0.4, 0.8])
    want = oracle.align_pairs(T, Q, W=W, threads=4)
    got = outs[None]["pairs"]
    assert got["edit"] == [int(x) for x in want.edit] and got["cigars"] == list(want.cigars)
    assert got["refc"] == [int(x) for x in want.ref_consumed]
    genome, reads, cs, cr = mapping_case(99)
    want = oracle.align_candidates(genome, reads, cs, cr, W=W, threads=4)
    got = outs[None]["mapping"]
    assert got["edit"] == [int(x) for x in want.edit] and got["cigars"] == list(want.cigars)
    assert got["refc"] == [int(x) for x in want.ref_consumed]
