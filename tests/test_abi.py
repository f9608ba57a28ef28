"""The C-ABI library on a machine without a GPU: it loads, exports every symbol include/scrooge_b200.h
declares, fails loudly (no CPU fallback) and its host-only helpers work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="scrooge_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(sglib):
    from scrooge_b200._lib import SIGNATURES
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(sglib, n), f"{n} declared in the header but not exported"
        assert n in SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(SIGNATURES) == names


def test_bench_helpers_live_in_their_own_library(sglib):
    """Synthetic generators, peak probes and the batch checker are not product code: they are declared in
    scrooge_b200_bench.h, exported by libscrooge_b200_bench.so and absent from libscrooge_b200.so."""
    import scrooge_b200
    from scrooge_b200._lib import BENCH_SIGNATURES
    bl = scrooge_b200.bench_lib()
    names = declared_symbols("scrooge_b200_bench.h")
    assert sorted(BENCH_SIGNATURES) == names and len(names) >= 7
    for n in names:
        assert hasattr(bl, n), f"{n} declared in scrooge_b200_bench.h but not exported"
        assert not hasattr(sglib, n), f"{n} is a bench helper but libscrooge_b200.so exports it"


def test_no_cpu_fallback(sglib):
    import scrooge_b200
    if sglib.sg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(scrooge_b200.ScroogeError) as e:
        scrooge_b200.Aligner()
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)
    with pytest.raises(scrooge_b200.ScroogeError):
        scrooge_b200.align_all(["ACGT"], ["ACG"])


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under scrooge_b200/ or include/ may import, load or link it."""
    for top in ("scrooge_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    text = open(os.path.join(d, f)).read()
                    for needle in ("oracle/", "libsgoracle", "import oracle", "from oracle", "sgo_", "libscrooge_ref"):
                        assert needle not in text, (f, needle)
    mk = open(os.path.join(ROOT, "Makefile")).read()
    assert "libsgoracle" not in mk and "genasm_oracle" not in mk


def test_synth_generator_is_deterministic_and_shaped(sglib):
    from scrooge_b200 import synth
    wl = synth.WORKLOADS["long_10kbp"]
    t1, l1, r1 = synth.pairs_host(wl, 5, 6)
    t2, l2, r2 = synth.pairs_host(wl, 7, 2)
    assert np.array_equal(t1[2], t2[0]) and np.array_equal(r1[3], r2[1]) and l1[2] == l2[0]
    assert set(np.unique(r1)) <= set(b"ACGT")
    assert all(10000 * 0.9 < x < 10000 * 1.1 + 64 for x in l1)
    assert t1.shape[1] % 16 == 0 and t1.shape[1] >= 10000 + 10000 // 8 + 32 + 64  # L + L/8 + 32 + slack
    # padding after the text is packable
    assert set(np.unique(t1[0, int(l1[0]):])) == {ord("A")}


def test_synth_error_rate(sglib, oracle):
    from scrooge_b200 import synth
    wl = synth.Workload("t", 2000, 0.10, synth.PACBIO, 64, 99)
    text, tlen, reads = synth.pairs_host(wl, 0, 20)
    T, Q = synth.pairs_as_strings(text, tlen, reads)
    res = oracle.align_pairs(T, Q)
    rate = float(np.mean(res.edit)) / 2000
    assert 0.07 < rate < 0.12  # observed distance is a little under e*L (adjacent edits cancel)
    tb, toff, qb, qoff = synth.pairs_as_blobs(text, tlen, reads)
    assert bytes(tb[int(toff[3]):int(toff[4])]).decode() == T[3] and bytes(qb[int(qoff[3]):int(qoff[4])]).decode() == Q[3]


def test_window_configuration_arguments(sglib):
    """sg_default_overlap follows scripts/profile.py:78 (O = min(W//2+1, W-1)); out-of-range windows are argument errors
    before any device is touched."""
    assert [sglib.sg_default_overlap(W) for W in (64, 32, 96, 128, 2, 3)] == [33, 17, 49, 65, 1, 2]
    h = C.c_void_p()
    for W, O in ((257, 129), (64, 64), (256, 127), (1, 0), (64, -1)):
        assert sglib.sg_ctx_create_wo(C.byref(h), None, 1, W, O) == 3, (W, O)
        assert b"window configuration" in sglib.sg_last_error()
    assert sglib.sg_ctx_create(C.byref(h), None, 1, 48) == 3


def test_host_renderer_matches_python(sglib):
    """CIGAR text from packed runs (sg_host_render.cpp: AVX-512 VBMI2 when the CPU has it, scalar otherwise): exact text,
    exact length, nothing written outside it -- around the 64-run block boundaries in particular."""
    sglib.sg_host_runs_text_len.restype = C.c_uint64
    sglib.sg_host_runs_text_len.argtypes = [C.c_void_p, C.c_uint64]
    sglib.sg_host_runs_render.restype = C.c_void_p
    sglib.sg_host_runs_render.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(1)
    for cnt in [1, 2, 21, 22, 42, 43, 63, 64, 65, 127, 128, 129, 191, 192, 193, 1000, 2113]:
        for hi in (64, 10, 32):   # all counts / one-digit counts only / the 64/33 window's range
            runs = ((rng.integers(0, 4, cnt) << 6) | rng.integers(1, hi, cnt)).astype(np.uint8)
            want = "".join(f"{int(b) & 63}{'=XID'[int(b) >> 6]}" for b in runs)
            n = sglib.sg_host_runs_text_len(runs.ctypes.data, cnt)
            assert n == len(want), (cnt, hi)
            out = np.full(n + 160, 0x7E, dtype=np.uint8)
            end = sglib.sg_host_runs_render(runs.ctypes.data, cnt, out.ctypes.data + 80)
            assert end == out.ctypes.data + 80 + n
            assert bytes(out[80:80 + n]).decode() == want, (cnt, hi)
            assert (out[:80] == 0x7E).all() and (out[80 + n:] == 0x7E).all(), ("wrote outside the text", cnt, hi)


def test_streaming_renderer_is_exact_at_every_alignment(sglib):
    """sg_host_runs_render_stream (what sg_result_render_all uses for large blobs: rendered into a cache-resident scratch,
    copied out with non-temporal stores for whole destination lines): exact text, nothing outside it, at every
    misalignment of the destination."""
    sglib.sg_host_runs_render_stream.restype = C.c_void_p
    sglib.sg_host_runs_render_stream.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(3)
    for cnt in [1, 5, 21, 22, 43, 64, 65, 200, 2113]:
        runs = ((rng.integers(0, 4, cnt) << 6) | rng.integers(1, 64, cnt)).astype(np.uint8)
        want = "".join(f"{int(b) & 63}{'=XID'[int(b) >> 6]}" for b in runs).encode()
        scratch = np.zeros(3 * cnt + 4096, dtype=np.uint8)
        for mis in range(64):
            out = np.full(len(want) + 256, 0x7E, dtype=np.uint8)
            dst = ((out.ctypes.data + 63) & ~63) + mis
            end = sglib.sg_host_runs_render_stream(runs.ctypes.data, cnt, dst, scratch.ctypes.data + (mis % 5))
            o0 = dst - out.ctypes.data
            assert end == dst + len(want)
            assert bytes(out[o0:o0 + len(want)]) == want, (cnt, mis)
            assert (out[:o0] == 0x7E).all() and (out[o0 + len(want):] == 0x7E).all(), ("wrote outside the text", cnt, mis)


def test_headers_are_plain_c(tmp_path, sglib):
    """The boundary is a C ABI: both public headers compile as C99 (-pedantic -Werror) and a C program links against the
    libraries and calls them."""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text('#include "scrooge_b200.h"\n#include "scrooge_b200_bench.h"\n#include <stdio.h>\n'
                   'int main(void) { sg_call_stats s; (void)s; printf("v%d devices=%d overlap=%d\\n", sg_version(), sg_device_count(), '
                   'sg_default_overlap(64)); return 0; }\n')
    libdir = os.path.join(ROOT, "scrooge_b200", "lib")
    exe = str(tmp_path / "c_abi")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir,
                    "-lscrooge_b200", "-lscrooge_b200_bench", "-Wl,-rpath," + libdir, "-o", exe], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("v1 devices=") and "overlap=33" in out.stdout, out.stdout + out.stderr
