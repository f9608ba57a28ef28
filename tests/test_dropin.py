"""The reference's own callers, UNMODIFIED, against this library.

oracle/Makefile `dropin` builds /root/reference/src/library_example.cu and src/tests.cu the way the reference's
Makefile does (nvcc -rdc=true over caller + SOURCES, Makefile:8-12,23-38) with src/genasm_gpu.cu replaced by
libscrooge_b200_rdc.a + -lscrooge_b200; the callers include the reference's own headers.  The CPU test builds them
(when /root/reference is present) and checks what they link against; the GPU tests run the prebuilt binaries, which
travel to the GPU box in oracle/_ref/."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
LIBDIR = os.path.join(ROOT, "scrooge_b200", "lib")
HAVE_REF = os.path.exists("/root/reference/src/library_example.cu")


def _nm(path, flags):
    out = subprocess.run(["nm", "-C"] + flags + [path], capture_output=True, text=True, check=True).stdout
    return out


def test_library_exports_what_the_reference_header_declares(sglib):
    """src/genasm_gpu.hpp:5-9: enabled_algorithm_log, two align_all overloads (mangled against the reference's struct
    names Genome / Read / Alignment), and the kernel in the rdc archive."""
    dyn = _nm(os.path.join(LIBDIR, "libscrooge_b200.so"), ["-D", "--defined-only"])
    assert "genasm_gpu::enabled_algorithm_log" in dyn
    assert "genasm_gpu::align_all(Genome&, std::vector<Read, std::allocator<Read> >&, long long*)" in dyn
    assert ("genasm_gpu::align_all(std::vector<std::__cxx11::basic_string<char, std::char_traits<char>, std::allocator<char> >"
            in dyn)
    rdc = _nm(os.path.join(LIBDIR, "libscrooge_b200_rdc.a"), ["--defined-only"])
    assert "genasm_gpu::ascii_to_twobit_strings(int, long long*, char**, char**)" in rdc


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference sources (build container only)")
def test_reference_callers_build_unmodified(sglib):
    from oracle.binding import build_dropin
    assert build_dropin(force=True)
    for name in ("dropin_library_example", "dropin_library_example_gxx", "dropin_tests"):
        exe = os.path.join(REF, name)
        und = _nm(exe, ["-D", "--undefined-only"])
        # the GPU entry points come from the product library, nothing of src/genasm_gpu.cu is in the binary
        assert "genasm_gpu::align_all(Genome&" in und, name
        needed = subprocess.run(["readelf", "-d", exe], capture_output=True, text=True, check=True).stdout
        assert "libscrooge_b200.so" in needed, name
    # the CPU half of the unmodified example runs here; the GPU half must fail loudly (no CPU fallback)
    p = subprocess.run([os.path.join(REF, "dropin_library_example_gxx")], capture_output=True, text=True)
    assert p.stdout.splitlines()[0] == "edit_distance:0 cigar:7="
    import scrooge_b200
    if scrooge_b200.lib().sg_device_count() == 0:
        assert p.returncode != 0 and "no CUDA device" in p.stderr


def test_headers_coexist_with_the_reference_util_hpp(tmp_path):
    """include/genasm_gpu.hpp next to the reference's util.hpp / genasm_cpu.hpp, in both include orders, and alone."""
    cases = {"alone": ('#include "genasm_gpu.hpp"\n', [])}
    if HAVE_REF:
        cases["ref_first"] = ('#include "genasm_cpu.hpp"\n#include "util.hpp"\n#include "genasm_gpu.hpp"\n#include "scrooge_io.hpp"\n',
                              ["-I/root/reference/src"])
        cases["ours_first"] = ('#include "genasm_gpu.hpp"\n#include "genasm_cpu.hpp"\n#include "util.hpp"\n', ["-I/root/reference/src"])
    body = ("int main() { std::vector<std::string> t{\"ACGT\"}, q{\"ACG\"}; Genome_t g; std::vector<Read_t> r;\n"
            "  auto a = genasm_gpu::align_all(t, q); auto b = genasm_gpu::align_all(g, r); return (int)(a.size() + b.size()); }\n")
    for name, (inc, flags) in cases.items():
        src = tmp_path / f"{name}.cpp"
        src.write_text(inc + body)
        # our include directory FIRST on the path: the worst case for a clash
        cmd = ["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "include")] + flags + [str(src)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, f"{name}: {p.stderr[:2000]}"


@pytest.mark.gpu
def test_unmodified_library_example_runs():
    for name in ("dropin_library_example", "dropin_library_example_gxx"):
        exe = os.path.join(REF, name)
        if not os.path.exists(exe):
            pytest.skip("oracle/_ref/dropin_* not prebuilt")
        p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr
        assert p.stdout.splitlines() == ["edit_distance:0 cigar:7="] * 4, p.stdout


@pytest.mark.gpu
def test_unmodified_reference_unit_tests_run():
    """src/tests.cu --unit_tests: bitvector KATs, its own launch of ascii_to_twobit_strings against the reference byte
    layout (src/tests.cu:583-650), CPU and GPU known answers, and CPU == GPU CIGAR equality over both interfaces."""
    exe = os.path.join(REF, "dropin_tests")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_tests not prebuilt")
    p = subprocess.run([exe, "--unit_tests"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    for t in ("ascii_to_two_bit_correctness_test", "cpu_algorithm_correctness_test", "gpu_algorithm_correctness_test",
              "library_interface_correctness_test"):
        assert f"PASSED {t}" in p.stdout, p.stdout
    assert "FAILED" not in p.stdout, p.stdout
