"""The measurement legs of bench_extra.py (pairs_point, mapping_point: what bench.py's extra_configs and the GPU test
test_bench_extra_configs run) executed on the CPU-only box with the HOST SIMULATION of the kernels (tests/sim) standing in
for the device: the same Python, the same oracle checks, the same result dictionaries -- so that a slip in that Python shows
here and not first on the GPU box.  Times are meaningless (a fake event clock); parity flags and shapes are not."""
import ctypes as C
import time

import numpy as np
import pytest
import torch

import test_kernel_sim as tks
from test_kernel_sim import sim  # noqa: F401  (the fixture that builds and loads tests/sim/_build/libsgsim.so)


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class FakeStream:
    cuda_stream = 0


@pytest.fixture
def fake_device(monkeypatch, sim):  # noqa: F811
    import bench_extra
    import scrooge_b200
    from scrooge_b200 import device, synth
    real_lib, real_bench = scrooge_b200.lib(), scrooge_b200.bench_lib()
    cpu = torch.device("cpu")

    def ptr(t):
        return C.c_void_p(0 if t is None else int(t.data_ptr()))

    def pack_2bit(ascii_t, n_bases=None):
        n = ascii_t.numel() if n_bases is None else n_bases
        words = int(real_lib.sg_packed_words(n))
        packed = torch.zeros(words, dtype=torch.int32)
        bad = int(real_lib.sg_host_pack_2bit(int(ascii_t.data_ptr()), n, int(packed.data_ptr()), 1))
        return packed, torch.tensor([-1 if bad == 2**64 - 1 else bad], dtype=torch.int64)

    class FakeAligner:
        """device.DeviceAligner on host tensors: sg_dev_align / scan / gather are the simulated kernels."""

        def __init__(self, W, n, dev, slab_bytes=0, O=None):
            assert (W, O) in ((64, None), (32, None), (64, 33), (32, 17))
            self.W, self.n = W, n
            self.out = device.AlignOut(edit=torch.zeros(n, dtype=torch.int64), ref_consumed=torch.zeros(n, dtype=torch.int64),
                                       nruns=torch.zeros(n, dtype=torch.int32), status=torch.zeros(n, dtype=torch.uint8),
                                       dc_entries=torch.zeros(n, dtype=torch.int64), windows=torch.zeros(n, dtype=torch.int32))
            self.slab = torch.zeros(max(slab_bytes, 16), dtype=torch.uint8) if slab_bytes else None
            self.run_off = torch.zeros(n + 1, dtype=torch.int64)

        def align(self, text, text_start, text_len, query, query_start, query_len, slab_off=None, distance_only=False, stats=True,
                  run_words=False):
            o = self.out
            ts, tl, qs, ql = (x.contiguous() for x in (text_start, text_len, query_start, query_len))
            if run_words:
                assert self.slab.data_ptr() % 4 == 0 and int((slab_off % 4).abs().sum()) == 0, "SG_FLAG_RUN_WORDS needs aligned slots"
            rc = sim.sim_delta_align(self.W, 1 if run_words and not distance_only else 0, 2, ptr(text), ptr(ts), ptr(tl), ptr(query), ptr(qs),
                                     ptr(ql), self.n, (1 if distance_only else 0) | (2 if run_words else 0), ptr(self.slab), ptr(slab_off),
                                     ptr(o.edit), ptr(o.ref_consumed), ptr(o.nruns), ptr(o.status), ptr(o.dc_entries if stats else None),
                                     ptr(o.windows if stats else None), None, None)
            assert rc == 0
            return o

        def compact(self, slab_off, runs=None):
            tmp = torch.zeros(self.n // 2048 + 4, dtype=torch.int64)
            sim.sim_scan_runs(ptr(self.out.nruns), self.n, ptr(self.run_off), ptr(tmp))
            if runs is None:
                runs = torch.zeros(max(int(self.run_off[-1]), 1), dtype=torch.uint8)
            sim.sim_gather_runs(32, ptr(self.slab), ptr(slab_off), ptr(self.out.nruns), ptr(self.run_off), self.n, ptr(runs), 2)
            return self.run_off, runs

    def synth_pairs_device(seed, first, n, L, err, ratio, slack, dev):
        wl = synth.Workload("x", L, err, ratio, 64, seed, slack)
        t, tl, r = synth.pairs_host(wl, first, n)
        return torch.from_numpy(t), torch.from_numpy(tl.astype(np.int64)), torch.from_numpy(r)

    class FakeBenchLib:
        def __getattr__(self, name):
            return getattr(real_bench, name)

        def sg_synth_genome(self, seed, first, n, out, d_out, stream):   # "device" memory is host memory here
            return real_bench.sg_synth_genome(seed, first, n, d_out, None, None)

        def sg_synth_reads(self, seed, first, n, L, err, ws, wi, wd, genome, G, reads, pos, on_device, stream):
            return real_bench.sg_synth_reads(seed, first, n, L, err, ws, wi, wd, genome, G, reads, pos, 0, None)

    class FakeLib:
        def __getattr__(self, name):
            return getattr(real_lib, name)

        def sg_dev_pack_2bit(self, d_ascii, n, d_packed, d_bad, stream):
            bad = int(real_lib.sg_host_pack_2bit(d_ascii, n, d_packed, 1))
            C.c_uint64.from_address(d_bad).value = bad
            return 0

    monkeypatch.setattr(bench_extra, "dev", cpu)
    monkeypatch.setattr(bench_extra, "stream", lambda: 0)
    monkeypatch.setattr(bench_extra, "lib", lambda: FakeLib())
    monkeypatch.setattr(scrooge_b200, "bench_lib", lambda: FakeBenchLib())
    monkeypatch.setattr(device, "pack_2bit", pack_2bit)
    monkeypatch.setattr(device, "DeviceAligner", FakeAligner)
    monkeypatch.setattr(device, "synth_pairs_device", synth_pairs_device)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    return bench_extra


def test_pairs_point_on_the_simulated_device(fake_device):
    from scrooge_b200 import synth
    for name in ("short_150bp", "short_150bp_w32"):
        pt = fake_device.pairs_point(synth.WORKLOADS[name], 96, False, 18000.0, check=96)
        assert pt["parity"] == {"checked": 96, "bit_exact": True}
        assert pt["pairs"] == 96 and pt["alignments_per_s_kernel"] > 0 and pt["windows_per_alignment"] > 1 and pt["mode"] == "full_cigar"
    pt = fake_device.pairs_point(synth.WORKLOADS["short_150bp"], 64, True, 18000.0, check=64)
    assert pt["parity"]["bit_exact"] and pt["mode"] == "distance_only"


@pytest.mark.parametrize("stress", [False, True])
def test_mapping_point_on_the_simulated_device(fake_device, stress):
    """Both run emissions are launched, both agree, the sample equals the oracle; the entry carries what bench.py copies."""
    m = fake_device.mapping_point(400_000, 64, stress, 18000.0, sub_batch=256, steps=1, ncand=8)   # two launches per pass
    assert m["parity"]["bit_exact"] is True and m["parity"]["checked"] == 256
    assert m["run_emission"] == "words" and set(m["kernel_ms_by_run_emission"]) == {"bytes", "words"}
    assert m["alignments"] == 512 and m["sub_batch"] == 256 and m["alignments_per_s_step"] > 0
    for k in ("genome_bases", "reads", "candidates_per_read", "read_len", "W", "step_ms", "kernel_ms", "int32_frac", "gcups_kernel", "workload"):
        assert k in m
    if stress:
        assert m["mean_edit_distance_first_2048"] > 3000 > m["true_start_mean_edit"]
    else:
        assert m["mean_edit_distance_first_2048"] < 1500


@pytest.mark.parametrize("W,L,err", [(64, 1000, 0.10), (32, 150, 0.05), (64, 777, 0.45)])
def test_gpu_variant_test_body_on_the_simulated_device(fake_device, monkeypatch, oracle, W, L, err):
    """tests/test_gpu_variants.py's device-API case, the very function the GPU box runs, on the simulated device with fewer
    pairs: its Python is exercised here first (device.check_runs runs the bench library's checking kernel in the simulation)."""
    import test_gpu_variants as tgv
    from scrooge_b200 import device
    monkeypatch.setattr(tgv, "N_PAIRS", 1100)

    def check_runs(runs, run_off, query_len, out, W, O=None):
        bad = torch.zeros(1, dtype=torch.int64)
        tks.sim.__wrapped__().sim_check_runs(C.c_void_p(runs.data_ptr()), C.c_void_p(run_off.data_ptr()), query_len.numel(), C.c_void_p(query_len.data_ptr()),
                                             C.c_void_p(out.edit.data_ptr()), C.c_void_p(out.ref_consumed.data_ptr()),
                                             W - (min(W // 2 + 1, W - 1) if O is None else O), C.c_void_p(bad.data_ptr()), 2)
        return int(bad[0])
    monkeypatch.setattr(device, "check_runs", check_runs)
    monkeypatch.setattr(torch, "device", lambda *a, **k: "cpu")
    tgv.test_device_api_run_words_equals_bytes_and_oracle(oracle, None, W, L, err)


def test_gpu_variant_child_program_compiles():
    import test_gpu_variants as tgv
    compile(tgv._CHILD, "child", "exec")
