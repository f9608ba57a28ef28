"""What the build produced, checked without a GPU: ptxas' resource report for every kernel (build/*.ptxas.log, written by the
Makefile) and the SASS of the product library -- the claims DESIGN.md makes about registers, spills, occupancy and the
sm_100a instructions in use (bulk-copy staging of the ingest kernel, tensor memory in the row-wise variant)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scrooge_b200", "lib", "libscrooge_b200.so")


@pytest.fixture(scope="module")
def ptxas():
    """kernel name (demangled prefix) -> (registers, spill bytes) from build/sg_device_api.ptxas.log"""
    import scrooge_b200
    if not os.path.exists(LIB):
        scrooge_b200.build()
    path = os.path.join(ROOT, "build", "sg_device_api.ptxas.log")
    if not os.path.exists(path):   # a library without its build directory: recompile the kernels' translation unit
        subprocess.run(["make", "-s", "-C", ROOT, "-W", "scrooge_b200/csrc/sg_device_api.cu", "build/sg_device_api.o"], check=True)
    log = open(path).read()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\w+)' for 'sm_100a'.*?(\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers", log, re.S):
        out[m.group(1)] = (int(m.group(4)), int(m.group(2)) + int(m.group(3)))
    assert len(out) >= 40
    return out


def test_no_kernel_spills(ptxas):
    assert [k for k, (_, spill) in ptxas.items() if spill] == []


def test_alignment_kernels_fit_their_occupancy(ptxas):
    """genasm_delta_kernel runs 6 CTAs x 4 warps per SM: 65 536 registers / 768 threads = 85 per thread at most;
    all run-emission variants included."""
    delta = {k: r for k, (r, _) in ptxas.items() if "genasm_delta_kernel" in k}
    assert len(delta) == 4, sorted(delta)          # W = 64 / 32  x  runs stored as bytes / as words
    assert max(delta.values()) <= 80, delta


def test_sass_has_the_sm100_instructions_the_design_names():
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not installed")
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    fn = None
    seen = {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\w+)", line)
        if m:
            fn = m.group(1)
            seen[fn] = set()
        elif fn:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                seen[fn].add(m.group(1).split(".")[0])
    bulk = [v for k, v in seen.items() if "pack_2bit_bulk_kernel" in k]
    assert bulk and all({"UBLKCP", "SYNCS"} <= v for v in bulk), "ingest kernel without bulk-copy staging / mbarrier waits"
    tmem = [v for k, v in seen.items() if "genasm_align_kernel" in k and {"LDTM", "STTM"} <= v]
    assert len(tmem) == 2, "the tensor-memory forefront variants (W = 64, 32) use tcgen05.ld/st"
    delta = [v for k, v in seen.items() if "genasm_delta_kernel" in k]
    assert len(delta) == 4 and all({"LOP3", "SHF", "PRMT", "LDS", "STS"} <= v and not ({"LDL", "STL"} & v) for v in delta)
