"""ctypes bindings for the parity checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package ``scrooge_b200`` never does.

* ``Oracle``   -- our plain-C restatement (oracle/genasm_oracle.c -> oracle/libsgoracle.so)
* ``RefCpu``   -- the UNMODIFIED reference genasm_cpu.cpp behind oracle/ref_shim.cpp
                  (oracle/_ref/libscrooge_ref_w{64,32}.so), when it has been built.
* ``RefGpu``   -- the UNMODIFIED reference genasm_gpu.cu compiled for sm_100a behind oracle/ref_gpu_shim.cpp
                  (oracle/_ref/libscrooge_refgpu_{default,best}.so): same-box GPU comparison point.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libsgoracle.so")
REF_DIR = os.path.join(HERE, "_ref")

# (W, O) pairs the reference is used with: in-file default (src/genasm_cpu.cpp:7-9) and the
# short-read setting (README.md:208, scripts/profile.py:78: O = min(W//2+1, W-1)).
CONFIGS = {64: 33, 32: 17}
# Further (W, O) builds of the unmodified reference (oracle/Makefile EXTRA_WO): the axes of the reference's window sweep
# (scripts/profile.py:66-100 cpu_sweep_wo / cpu_sweep_o, :595-640 accuracy sweeps with W in 32, 64, 96, 128 and
# O = min(W//2+1, W-1) or free), up to W = 256; from (128, 1) on W - O > 63 (runs longer than a run byte's 6-bit count).
EXTRA_CONFIGS = [(64, 20), (64, 48), (64, 1), (48, 25), (32, 8), (16, 9), (96, 49), (128, 65), (128, 100), (128, 1), (64, 0), (256, 129), (160, 81)]


def ref_lib_path(W: int, O: Optional[int] = None) -> str:
    if O is None or CONFIGS.get(W) == O:
        return os.path.join(REF_DIR, f"libscrooge_ref_w{W}.so")
    return os.path.join(REF_DIR, f"libscrooge_ref_w{W}_o{O}.so")


def build(force: bool = False) -> None:
    """Compile the C restatement, and the reference shim when /root/reference is present."""
    if force or not os.path.exists(ORACLE_SO) or (
        os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "genasm_oracle.c"))
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    # the reference's own CUDA kernel for sm_100a (same-box comparison point); ~30 s, only when stale or absent
    src = "/root/reference/src/genasm_gpu.cu"
    out = os.path.join(REF_DIR, "libscrooge_refgpu_best.so")
    if os.path.exists(src) and (force or not os.path.exists(out)
                                or os.path.getmtime(out) < os.path.getmtime(os.path.join(HERE, "ref_gpu_shim.cpp"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "refgpu"], stdout=subprocess.DEVNULL)
    build_dropin(force)


def build_dropin(force: bool = False) -> bool:
    """The reference's own callers (src/library_example.cu, src/tests.cu), unmodified, linked against the product
    library in place of src/genasm_gpu.cu (oracle/Makefile `dropin`).  Needs /root/reference and the built product
    library; returns whether the binaries exist afterwards."""
    outs = [os.path.join(REF_DIR, n) for n in ("dropin_library_example", "dropin_library_example_gxx", "dropin_tests")]
    lib_dir = os.path.join(os.path.dirname(HERE), "scrooge_b200", "lib")
    deps = [os.path.join(lib_dir, "libscrooge_b200.so"), os.path.join(lib_dir, "libscrooge_b200_rdc.a")]
    if os.path.exists("/root/reference/src/library_example.cu") and all(os.path.exists(d) for d in deps):
        stale = force or not all(os.path.exists(o) for o in outs) or \
            min(os.path.getmtime(o) for o in outs) < os.path.getmtime(deps[1])
        if stale:
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin"], stdout=subprocess.DEVNULL)
    return all(os.path.exists(o) for o in outs)


def _blob(strings: Sequence[str | bytes]) -> Tuple[bytes, np.ndarray]:
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


def _u64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _i64p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class AlignResult:
    """edit distances, CIGAR strings, consumed reference prefix lengths, work counters."""

    def __init__(self, edit, cigars, ref_consumed=None, stats=None, core_ns=0):
        self.edit = edit
        self.cigars = cigars
        self.ref_consumed = ref_consumed
        self.stats = stats  # dict(windows, dc_entries, tb_steps) or None
        self.core_ns = core_ns


def cigar_ref_consumed(cigar: str) -> int:
    """#(=, X, D) characters: the consumed reference prefix implied by a CIGAR (src/tests.cu:62-86)."""
    total, num = 0, 0
    for ch in cigar:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            if ch in "=XDM":
                total += num
            num = 0
    return total


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.sgo_align_pairs.restype = C.c_int
        self.lib.sgo_align_candidates.restype = C.c_int
        self.lib.sgo_validate_cigar.restype = C.c_int
        self.lib.sgo_ascii_to_twobit_ref_layout.restype = C.c_int
        self.lib.sgo_max_threads.restype = C.c_int

    def max_threads(self) -> int:
        return int(self.lib.sgo_max_threads())

    def align_pairs_blob(self, tblob: bytes, toff: np.ndarray, qblob: bytes, qoff: np.ndarray,
                         W: int = 64, O: Optional[int] = None, threads: int = 1,
                         want_cigars: bool = True) -> AlignResult:
        O = CONFIGS[W] if O is None else O
        n = len(toff) - 1
        edit = np.zeros(n, dtype=np.int64)
        refc = np.zeros(n, dtype=np.uint64)
        stats = np.zeros(3, dtype=np.uint64)
        cig = C.create_string_buffer(int(4 * int(qoff[-1]) + n + 1))
        ns = C.c_int64(0)
        rc = self.lib.sgo_align_pairs(C.c_int(W), C.c_int(O), tblob, _u64p(toff), qblob, _u64p(qoff),
                                      C.c_uint64(n), C.c_int(threads), _i64p(edit), _u64p(refc), cig,
                                      _u64p(stats), C.byref(ns))
        if rc != 0:
            raise ValueError(f"oracle: sgo_align_pairs failed with {rc}")
        cigars: List[str] = []
        if want_cigars:
            raw = cig.raw
            for p in range(n):
                s = 4 * int(qoff[p]) + p
                e = raw.index(b"\0", s)
                cigars.append(raw[s:e].decode())
        return AlignResult(edit, cigars, refc,
                           dict(windows=int(stats[0]), dc_entries=int(stats[1]), tb_steps=int(stats[2])),
                           int(ns.value))

    def align_pairs(self, texts: Sequence[str], queries: Sequence[str], W: int = 64,
                    O: Optional[int] = None, threads: int = 1) -> AlignResult:
        assert len(texts) == len(queries)
        tblob, toff = _blob(texts)
        qblob, qoff = _blob(queries)
        return self.align_pairs_blob(tblob, toff, qblob, qoff, W, O, threads)

    def align_candidates(self, genome: str | bytes, reads: Sequence[str], cand_start: Sequence[int],
                         cand_read: Sequence[int], W: int = 64, O: Optional[int] = None,
                         threads: int = 1) -> AlignResult:
        O = CONFIGS[W] if O is None else O
        g = genome.encode() if isinstance(genome, str) else bytes(genome)
        rblob, roff = _blob(reads)
        cs = np.asarray(cand_start, dtype=np.uint64)
        cr = np.asarray(cand_read, dtype=np.uint32)
        n = len(cs)
        lens = (roff[1:] - roff[:-1])[cr] if n else np.zeros(0, dtype=np.uint64)
        coff = np.zeros(n + 1, dtype=np.uint64)
        if n:
            coff[1:] = np.cumsum(4 * lens + 1, dtype=np.uint64)
        edit = np.zeros(n, dtype=np.int64)
        refc = np.zeros(n, dtype=np.uint64)
        stats = np.zeros(3, dtype=np.uint64)
        cig = C.create_string_buffer(int(coff[-1]) + 1)
        ns = C.c_int64(0)
        rc = self.lib.sgo_align_candidates(C.c_int(W), C.c_int(O), g, C.c_uint64(len(g)), rblob, _u64p(roff),
                                           C.c_uint64(len(reads)), _u64p(cs),
                                           cr.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(n),
                                           C.c_int(threads), _i64p(edit), _u64p(refc), cig, _u64p(coff),
                                           _u64p(stats), C.byref(ns))
        if rc != 0:
            raise ValueError(f"oracle: sgo_align_candidates failed with {rc}")
        raw = cig.raw
        cigars = []
        for c in range(n):
            s = int(coff[c])
            cigars.append(raw[s:raw.index(b"\0", s)].decode())
        return AlignResult(edit, cigars, refc,
                           dict(windows=int(stats[0]), dc_entries=int(stats[1]), tb_steps=int(stats[2])),
                           int(ns.value))

    def validate_cigar(self, cigar: str, ref: str, read: str, edit_distance: int, ref_start: int = 0) -> int:
        r, q = ref.encode(), read.encode()
        return int(self.lib.sgo_validate_cigar(cigar.encode(), r, C.c_uint64(len(r)), C.c_uint64(ref_start),
                                               q, C.c_uint64(len(q)), C.c_int64(edit_distance)))

    def twobit_ref_layout(self, ascii_str: str) -> bytes:
        a = ascii_str.encode()
        out = C.create_string_buffer((len(a) + 3) // 4 + 1)
        rc = self.lib.sgo_ascii_to_twobit_ref_layout(a, C.c_size_t(len(a)), out)
        if rc != 0:
            raise ValueError("non-ACGT character")
        return out.raw[: (len(a) + 3) // 4]


class RefCpu:
    """The unmodified reference CPU aligner (genasm_cpu::align_all) for one (W, O) build."""

    def __init__(self, W: int = 64, O: Optional[int] = None):
        path = ref_lib_path(W, O)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.W = W
        self.O = CONFIGS[W] if O is None else O
        self.lib = C.CDLL(path)
        for f in ("ref_align_pairs", "ref_align_mapping", "ref_config_w", "ref_config_o", "ref_max_threads"):
            getattr(self.lib, f).restype = C.c_int
        assert self.lib.ref_config_w() == W and self.lib.ref_config_o() == self.O

    @staticmethod
    def available(W: int = 64, O: Optional[int] = None) -> bool:
        return os.path.exists(ref_lib_path(W, O))

    def max_threads(self) -> int:
        return int(self.lib.ref_max_threads())

    def align_pairs_blob(self, tblob: bytes, toff: np.ndarray, qblob: bytes, qoff: np.ndarray,
                         threads: int = 1, want_cigars: bool = True) -> AlignResult:
        n = len(toff) - 1
        edit = np.zeros(n, dtype=np.int64)
        cig = C.create_string_buffer(int(4 * int(qoff[-1]) + n + 1))
        ns = C.c_int64(0)
        rc = self.lib.ref_align_pairs(tblob, _u64p(toff), qblob, _u64p(qoff), C.c_uint64(n), C.c_int(threads),
                                      _i64p(edit), cig, C.byref(ns))
        if rc != 0:
            raise RuntimeError(f"reference align_all returned {rc}")
        cigars = []
        if want_cigars:
            raw = cig.raw
            for p in range(n):
                s = 4 * int(qoff[p]) + p
                cigars.append(raw[s:raw.index(b"\0", s)].decode())
        return AlignResult(edit, cigars, None, None, int(ns.value))

    def align_pairs(self, texts: Sequence[str], queries: Sequence[str], threads: int = 1) -> AlignResult:
        tblob, toff = _blob(texts)
        qblob, qoff = _blob(queries)
        return self.align_pairs_blob(tblob, toff, qblob, qoff, threads)

    def align_mapping(self, genome: str, reads: Sequence[str], locations: Sequence[Sequence[int]],
                      threads: int = 1) -> AlignResult:
        """locations[r] = candidate start positions of read r (read-major output order)."""
        g = genome.encode()
        rblob, roff = _blob(reads)
        begin = np.zeros(len(reads) + 1, dtype=np.uint64)
        begin[1:] = np.cumsum([len(l) for l in locations], dtype=np.uint64)
        starts = np.asarray([s for l in locations for s in l], dtype=np.uint64)
        n = len(starts)
        lens = np.repeat(roff[1:] - roff[:-1], [len(l) for l in locations]).astype(np.uint64)
        coff = np.zeros(n + 1, dtype=np.uint64)
        if n:
            coff[1:] = np.cumsum(4 * lens + 1, dtype=np.uint64)
        edit = np.zeros(n, dtype=np.int64)
        cig = C.create_string_buffer(int(coff[-1]) + 1)
        ns = C.c_int64(0)
        rc = self.lib.ref_align_mapping(g, C.c_uint64(len(g)), rblob, _u64p(roff), C.c_uint64(len(reads)),
                                        _u64p(begin), _u64p(starts), C.c_int(threads), _i64p(edit), cig,
                                        _u64p(coff), C.byref(ns))
        if rc != 0:
            raise RuntimeError(f"reference align_all (mapping) returned {rc}")
        raw = cig.raw
        cigars = [raw[int(coff[c]):raw.index(b"\0", int(coff[c]))].decode() for c in range(n)]
        return AlignResult(edit, cigars, None, None, int(ns.value))


class RefGpu:
    """The unmodified reference GPU aligner (genasm_gpu::align_all, src/genasm_gpu.cu:982-1065) built for sm_100a.
    ``build``: "default" = in-file knobs (SENE+DENT, no ET, 20 blocks/SM); "best" = the paper's headline knobs
    (SENE+DENT+ET, 28 blocks/SM, scripts/plot.py:1277).  It exit()s on CUDA errors, so run it in a subprocess."""

    def __init__(self, build: str = "best"):
        path = os.path.join(REF_DIR, f"libscrooge_refgpu_{build}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.build = build
        self.lib = C.CDLL(path)
        self.lib.refgpu_align_pairs.restype = C.c_int
        self.lib.refgpu_config_w.restype = C.c_int

    @staticmethod
    def available(build: str = "best") -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libscrooge_refgpu_{build}.so"))

    def align_pairs_blob(self, tblob: bytes, toff: np.ndarray, qblob: bytes, qoff: np.ndarray,
                         want_cigars: bool = True) -> AlignResult:
        n = len(toff) - 1
        edit = np.zeros(n, dtype=np.int64)
        cig = C.create_string_buffer(int(4 * int(qoff[-1]) + n + 1))
        ns, total = C.c_int64(0), C.c_int64(0)
        rc = self.lib.refgpu_align_pairs(tblob, _u64p(toff), qblob, _u64p(qoff), C.c_uint64(n), _i64p(edit), cig,
                                         C.byref(ns), C.byref(total))
        if rc != 0:
            raise RuntimeError(f"reference genasm_gpu::align_all returned {rc}")
        cigars = []
        if want_cigars:
            raw = cig.raw
            for p in range(n):
                s = 4 * int(qoff[p]) + p
                cigars.append(raw[s:raw.index(b"\0", s)].decode())
        res = AlignResult(edit, cigars, None, None, int(ns.value))
        res.total_ns = int(total.value)
        return res
