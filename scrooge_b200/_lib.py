"""ctypes binding of libscrooge_b200.so (the C ABI declared in include/scrooge_b200.h).

The library is built in-tree by the repository Makefile (``__graft_entry__.build()``).  There is no
Python or CPU fallback: if the shared library is missing, importing this module's ``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# SG_LIB: another build of the same library (A/B timing of kernel variants by tools/kernel_time.py); never a fallback
LIB_PATH = os.environ.get("SG_LIB") or os.path.join(ROOT, "scrooge_b200", "lib", "libscrooge_b200.so")

SG_OK = 0
SG_ERR_CUDA = 1
SG_ERR_BAD_BASE = 2
SG_ERR_BAD_ARG = 3
SG_ERR_OOM = 4
SG_ERR_CIGAR_OVERFLOW = 5
SG_ERR_NO_REFERENCE = 6
SG_FLAG_DISTANCE_ONLY = 1
SG_FLAG_RUN_WORDS = 2

u64, i64, u32, i32, vp, cp = C.c_uint64, C.c_int64, C.c_uint32, C.c_int, C.c_void_p, C.c_char_p
dbl = C.c_double

# name -> (restype, argtypes); every symbol include/scrooge_b200.h declares
SIGNATURES = {
    "sg_last_error": (cp, []),
    "sg_version": (i32, []),
    "sg_device_count": (i32, []),
    "sg_ctx_create": (i32, [C.POINTER(vp), C.POINTER(i32), i32, i32]),
    "sg_ctx_create_wo": (i32, [C.POINTER(vp), C.POINTER(i32), i32, i32, i32]),
    "sg_default_overlap": (i32, [i32]),
    "sg_ctx_window": (i32, [vp]),
    "sg_ctx_overlap": (i32, [vp]),
    "sg_ctx_destroy": (None, [vp]),
    "sg_ctx_num_devices": (i32, [vp]),
    "sg_align_pairs": (i32, [vp, vp, vp, vp, vp, u64, u32, C.POINTER(vp)]),
    "sg_align_pairs_v": (i32, [vp, vp, vp, vp, vp, u64, u32, C.POINTER(vp)]),
    "sg_set_reference": (i32, [vp, vp, u64]),
    "sg_align_candidates": (i32, [vp, vp, vp, u64, vp, vp, u64, u32, C.POINTER(vp)]),
    "sg_align_candidates_v": (i32, [vp, vp, vp, u64, vp, vp, u64, u32, C.POINTER(vp)]),
    "sg_result_count": (u64, [vp]),
    "sg_result_edit_distances": (vp, [vp]),
    "sg_result_ref_consumed": (vp, [vp]),
    "sg_result_run_offsets": (vp, [vp]),
    "sg_result_runs": (vp, [vp]),
    "sg_result_kernel_ns": (i64, [vp]),
    "sg_result_total_ns": (i64, [vp]),
    "sg_result_cigar_len": (u64, [vp, u64]),
    "sg_result_render_cigar": (i64, [vp, u64, vp, u64]),
    "sg_result_entries": (i64, [vp, u64, vp, u64]),
    "sg_result_render_all": (u64, [vp, vp, u64, vp, i32]),
    "sg_result_free": (None, [vp]),
    "sg_result_stats": (i32, [vp, vp]),
    "sg_trim_host_cache": (None, []),
    "sg_plan_sub_batches": (u64, [vp, u64, u64, u64, u64, u64, i32, vp, u64]),
    "sg_host_alloc": (vp, [u64]),
    "sg_host_free": (None, [vp]),
    "sg_host_pack_2bit": (u64, [vp, u64, vp, i32]),
    "sg_host_pack_isa": (i32, []),
    "sg_packed_words": (u64, [u64]),
    "sg_dev_pack_2bit": (i32, [vp, u64, vp, vp, vp]),
    "sg_dev_pack_2bit_ex": (i32, [vp, u64, vp, vp, u32, vp]),
    "sg_dev_align": (i32, [i32, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "sg_dev_align_wo": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "sg_dev_align_ordered": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, u64, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "sg_dev_align_geometry_wo": (i32, [i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "sg_scan_tmp_bytes": (u64, [u64]),
    "sg_dev_scan_runs": (i32, [vp, u64, vp, vp, vp]),
    "sg_dev_gather_runs": (i32, [vp, vp, vp, vp, u64, vp, vp]),
    "sg_dev_gather_runs_sized": (i32, [vp, vp, vp, vp, u64, vp, u64, vp]),
    "sg_dev_align_geometry": (i32, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
}

# name -> (restype, argtypes); every symbol include/scrooge_b200_bench.h declares (libscrooge_b200_bench.so: synthetic
# generators, peak probes, the batch checker -- measurement and test helpers, not part of the product library)
BENCH_SIGNATURES = {
    "sg_bench_last_error": (cp, []),
    "sg_dev_int32_peak": (i32, [i32, dbl, C.POINTER(dbl)]),
    "sg_synth_text_stride": (u64, [u32, u32]),
    "sg_dev_check_runs": (i32, [vp, vp, u64, vp, vp, vp, u32, vp, vp]),
    "sg_synth_pairs_host": (i32, [u64, u64, u64, u32, dbl, u32, u32, u32, u32, vp, u64, vp, vp]),
    "sg_dev_synth_pairs": (i32, [u64, u64, u64, u32, dbl, u32, u32, u32, u32, vp, u64, vp, vp, vp]),
    "sg_synth_genome": (i32, [u64, u64, u64, vp, vp, vp]),
    "sg_synth_reads": (i32, [u64, u64, u64, u32, dbl, u32, u32, u32, vp, u64, vp, vp, i32, vp]),
}

class CallStats(C.Structure):
    """sg_call_stats (include/scrooge_b200.h)."""
    _fields_ = [("total_ns", i64), ("kernel_ns", i64), ("upload_ns", i64), ("pack_thread_ns", i64), ("wait_ns", i64),
                ("host_other_ns", i64), ("h2d_ascii_bytes", u64), ("h2d_packed_bytes", u64), ("h2d_other_bytes", u64),
                ("d2h_bytes", u64), ("n_devices", u32), ("sub_batches", u32), ("host_threads_per_device", u32), ("packers_in_use", u32)]


_lib = None
_bench = None


class ScroogeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"scrooge_b200 error {code}: {message}")
        self.code = code


def build() -> None:
    """Compile libscrooge_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", ROOT, "all"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()). "
                "scrooge_b200 has no Python/CPU fallback for the alignment path."
            )
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def bench_lib() -> C.CDLL:
    """libscrooge_b200_bench.so (include/scrooge_b200_bench.h); independent of the product library."""
    global _bench
    if _bench is None:
        path = os.environ.get("SG_BENCH_LIB") or os.path.join(os.path.dirname(LIB_PATH), "libscrooge_b200_bench.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `make` (or __graft_entry__.build())")
        l = C.CDLL(path)
        for name, (res, args) in BENCH_SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _bench = l
    return _bench


def bench_check(rc: int) -> None:
    if rc != SG_OK:
        raise ScroogeError(rc, bench_lib().sg_bench_last_error().decode(errors="replace"))


def check(rc: int) -> None:
    if rc != SG_OK:
        raise ScroogeError(rc, lib().sg_last_error().decode(errors="replace"))
