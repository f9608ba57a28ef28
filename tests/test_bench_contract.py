"""bench.py prints ONE JSON line with every key the measurement contract names (a reduced batch; the values are checked for
shape, not for speed)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "gpu_launches", "clocks", "e2e", "roofline", "cpu_baseline"}


@pytest.mark.gpu
def test_bench_line_has_every_contract_key():
    r = subprocess.run([sys.executable, "bench.py", "--pairs", "131072", "--steps", "2", "--warmup", "3"], cwd=ROOT, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["metric"] == "alignments_per_second" and d["unit"] == "alignments/s" and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "long_10kbp" and "model" not in d["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"]) and 0 < d["roofline"]["frac"] < 1
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["gpu_launches"] > 0 and d["parity"]["bit_exact"] is True


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
