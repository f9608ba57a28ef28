"""bench.py prints ONE JSON line with every key the measurement contract names (a reduced batch; the values are checked for
shape, not for speed)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "gpu_launches", "clocks", "e2e", "e2e_rendered", "roofline", "cpu_baseline", "extra_configs"}


@pytest.mark.gpu
def test_bench_line_has_every_contract_key():
    r = subprocess.run([sys.executable, "bench.py", "--pairs", "131072", "--steps", "2", "--warmup", "3", "--no-extra"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["metric"] == "alignments_per_second" and d["unit"] == "alignments/s" and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "long_10kbp" and "model" not in d["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and 0 < d["roofline"]["frac"] < 1
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ceiling", "frac_of_ceiling", "breakdown_per_rank"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0
    # what crossed PCIe is less than the ASCII input whenever the host packed a part of it
    assert d["e2e"]["h2d_bytes_per_step"] <= d["e2e"]["input_ascii_bytes_per_step"] * 1.01
    assert 0 < d["e2e_rendered"]["value"] <= d["e2e"]["value"] * 1.05 and d["e2e_rendered"]["cigar_text_bytes_per_step"] > 0
    assert 0 < d["roofline"]["frac_dual_pipe"] < d["roofline"]["frac"]
    assert d["extra_configs"] is None   # --no-extra; the full line is checked by test_bench_extra_configs
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["gpu_launches"] > 0 and d["parity"]["bit_exact"] is True


@pytest.mark.gpu
def test_bench_extra_configs():
    """The extra_configs legs at reduced sizes: short reads at both window settings and read mapping (device value, end to
    end, parity sample each)."""
    code = ("import json, torch, bench_extra; from scrooge_b200 import device, synth; torch.cuda.set_device(0); "
            "pk = device.int32_peak(2, 20.0); out = [bench_extra.pairs_leg(synth.WORKLOADS[w], 400000, pk, n_e2e=400000, check=1024) "
            "for w in ('short_150bp', 'short_150bp_w32')]; "
            "out.append(bench_extra.mapping_point(50_000_000, 16384, False, pk, sub_batch=16384 * 8, steps=1, e2e_reads=4096)); "
            "out.append(bench_extra.mapping_point(50_000_000, 4096, True, pk, sub_batch=4096 * 8, steps=1)); print(json.dumps(out))")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    legs = json.loads([l for l in r.stdout.splitlines() if l.startswith("[")][-1])
    for leg in legs[:2]:
        assert leg["value"] > 0 and leg["e2e"]["value"] > 0 and leg["parity"]["bit_exact"] and leg["e2e"]["parity"]["bit_exact"]
    assert legs[2]["alignments_per_s_step"] > 0 and legs[2]["parity"]["bit_exact"] and legs[2]["e2e"]["parity"]["bit_exact"]
    assert legs[3]["parity"]["bit_exact"] and legs[3]["mean_edit_distance_first_2048"] > 3 * legs[2]["mean_edit_distance_first_2048"]


def test_reference_arm_line_shape():
    """The reference arm's JSON line (CPU only: runs here too), on a reduced sample."""
    from oracle.binding import RefCpu
    if not RefCpu.available(64):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "512"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference" and d["metric"] == "alignments_per_second" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
