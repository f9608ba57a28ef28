"""The CUDA-free host helpers (scrooge_b200/csrc/sg_host_threads.h: CPU lists, CPUs per GPU, the packer team, the ingest
tuner) as a plain C++ unit-test program: built with g++ and run here, no GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_thread_helpers(tmp_path):
    exe = str(tmp_path / "host_units")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_units.cpp")],
                   check=True, capture_output=True, text=True)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "host units ok" in p.stdout, p.stdout + p.stderr
