"""Host-side mirror of the reference's library interface on top of the C ABI.

``align_all(texts, queries)`` and ``align_all(reference, reads)`` keep the names, argument order and result
order of the reference's ``genasm_gpu::align_all`` overloads (reference src/genasm_gpu.hpp:7-8,
src/library_example.cu:25-88); ``Aligner`` is the persistent context underneath (one per process is enough,
it owns the GPUs, streams and the replicated packed reference).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import SG_FLAG_DISTANCE_ONLY, ScroogeError, check, lib

OPS = "=XID"


@dataclass
class Alignment:
    """reference Alignment_t (src/util.hpp:38-41) plus the consumed reference prefix."""
    cigar: str
    edit_distance: int
    ref_consumed: int = 0


@dataclass
class CandidateLocation:
    """reference CandidateLocation_t (src/util.hpp:22-30); only start_in_reference is used by the aligner."""
    start_in_reference: int
    strand: bool = True
    chromosome: str = ""
    start_in_chromosome: int = 0


@dataclass
class Read:
    """reference Read_t (src/util.hpp:32-36)."""
    description: str
    content: str
    locations: List[CandidateLocation] = field(default_factory=list)


@dataclass
class Genome:
    """reference Genome_t (src/util.hpp:16-19)."""
    content: str
    chromosome_starts: dict = field(default_factory=dict)


def _blob(strings: Sequence[Union[str, bytes]]) -> Tuple[bytes, np.ndarray]:
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return b"".join(bs), off


class Result:
    """Owns an sg_result; arrays are copied out on access."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        l = lib()
        self.count = int(l.sg_result_count(handle))
        self.kernel_ns = int(l.sg_result_kernel_ns(handle))
        self.total_ns = int(l.sg_result_total_ns(handle))

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().sg_result_free(self._h)
            except Exception:   # interpreter shutdown: the module globals are already gone
                pass
            self._h = None

    def stats(self) -> dict:
        """sg_call_stats of the call that produced this result: where its time and its PCIe bytes went."""
        st = _lib.CallStats()
        check(lib().sg_result_stats(self._h, C.byref(st)))
        return {name: int(getattr(st, name)) for name, _ in st._fields_}

    def _arr(self, ptr, n, dtype):
        if not ptr or n == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).copy()

    @property
    def edit_distances(self) -> np.ndarray:
        return self._arr(lib().sg_result_edit_distances(self._h), self.count, np.int64)

    @property
    def ref_consumed(self) -> np.ndarray:
        return self._arr(lib().sg_result_ref_consumed(self._h), self.count, np.uint64)

    @property
    def run_offsets(self) -> np.ndarray:
        return self._arr(lib().sg_result_run_offsets(self._h), self.count + 1, np.uint64)

    @property
    def runs(self) -> np.ndarray:
        off = self.run_offsets
        total = int(off[-1]) if len(off) else 0
        return self._arr(lib().sg_result_runs(self._h), total, np.uint8)

    def cigar(self, idx: int) -> str:
        l = lib()
        n = int(l.sg_result_cigar_len(self._h, idx))
        buf = C.create_string_buffer(n + 1)
        got = l.sg_result_render_cigar(self._h, idx, buf, n + 1)
        if got < 0:
            raise ScroogeError(_lib.SG_ERR_BAD_ARG, "render_cigar failed")
        return buf.value.decode()

    def cigar_text(self, threads: int = 0, out: Optional[Tuple[np.ndarray, np.ndarray]] = None):
        """All CIGAR texts as one uint8 array + offsets (count+1), rendered by the library's host threads.  `out` = a
        (blob, offsets) pair from an earlier call to reuse (a caller that renders batch after batch keeps its buffer:
        fresh pages for gigabytes of text cost more than the rendering)."""
        off = out[1] if out is not None and len(out[1]) == self.count + 1 else np.zeros(self.count + 1, dtype=np.uint64)
        blob = out[0].base if out is not None and out[0].base is not None else (out[0] if out is not None else None)
        if blob is not None:   # one pass when the buffer is large enough
            total = int(lib().sg_result_render_all(self._h, blob.ctypes.data, blob.size, off.ctypes.data, threads))
            if total <= blob.size:
                return blob[:total], off
        total = int(lib().sg_result_render_all(self._h, None, 0, off.ctypes.data, threads))
        blob = np.empty(max(total + total // 16, 1), dtype=np.uint8)
        lib().sg_result_render_all(self._h, blob.ctypes.data, blob.size, off.ctypes.data, threads)
        return blob[:total], off

    def cigars(self) -> List[str]:
        """All CIGAR strings ("%d%c" per run, reference src/genasm_gpu.cu:881-888)."""
        blob, off = self.cigar_text()
        raw = blob.tobytes()
        return [raw[int(off[i]):int(off[i + 1])].decode() for i in range(self.count)]

    def alignments(self) -> List[Alignment]:
        ed, rc, cg = self.edit_distances, self.ref_consumed, self.cigars()
        return [Alignment(cg[i], int(ed[i]), int(rc[i])) for i in range(self.count)]


class Aligner:
    """sg_ctx wrapper.  W=64 -> O=33 (reference default), W=32 -> O=17 (reference short-read setting); any other
    window configuration the reference can be rebuilt with (-DCLI_W/-DCLI_O, src/genasm_cpu.cpp:22-35) within
    2 <= W <= 256, 0 <= O < W, W-O <= 128 by passing O (default O = min(W//2+1, W-1), scripts/profile.py:78)."""

    def __init__(self, W: int = 64, n_gpus: int = 0, device_ids: Optional[Sequence[int]] = None, O: Optional[int] = None):
        h = C.c_void_p()
        if O is None:
            O = min(W // 2 + 1, W - 1)
        if device_ids is not None:
            ids = (C.c_int * len(device_ids))(*device_ids)
            check(lib().sg_ctx_create_wo(C.byref(h), ids, len(device_ids), W, O))
        else:
            check(lib().sg_ctx_create_wo(C.byref(h), None, n_gpus, W, O))
        self._h = h
        self.W = W
        self.O = O
        self._genome_keepalive = None

    def close(self):
        if getattr(self, "_h", None):
            lib().sg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: the module globals lib() needs may already be gone
            pass

    @property
    def num_devices(self) -> int:
        return int(lib().sg_ctx_num_devices(self._h))

    def align_pairs_blob(self, tblob, toff: np.ndarray, qblob, qoff: np.ndarray, distance_only: bool = False) -> Result:
        """tblob/qblob: bytes or a uint8 numpy array (may be pinned); offsets: uint64 arrays with n+1 entries."""
        n = len(toff) - 1
        assert len(qoff) - 1 == n
        toff = np.ascontiguousarray(toff, dtype=np.uint64)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        out = C.c_void_p()
        check(lib().sg_align_pairs(self._h, _ptr(tblob), toff.ctypes.data, _ptr(qblob), qoff.ctypes.data, n,
                                   SG_FLAG_DISTANCE_ONLY if distance_only else 0, C.byref(out)))
        return Result(out)

    def align_pairs(self, texts: Sequence[str], queries: Sequence[str], distance_only: bool = False) -> Result:
        if len(texts) != len(queries):
            raise ValueError("texts and queries differ in size")  # reference asserts, src/genasm_gpu.cu:984
        tblob, toff = _blob(texts)
        qblob, qoff = _blob(queries)
        return self.align_pairs_blob(tblob, toff, qblob, qoff, distance_only)

    def align_pairs_v(self, texts: Sequence[Union[str, bytes]], queries: Sequence[Union[str, bytes]],
                      distance_only: bool = False) -> Result:
        """Same as align_pairs through the vectored entry point (sg_align_pairs_v): one pointer + length per string,
        nothing is concatenated -- the form the C++ drop-in uses for std::vector<std::string>."""
        if len(texts) != len(queries):
            raise ValueError("texts and queries differ in size")
        tb = [t.encode() if isinstance(t, str) else bytes(t) for t in texts]
        qb = [q.encode() if isinstance(q, str) else bytes(q) for q in queries]
        n = len(tb)
        tp = (C.c_char_p * max(n, 1))(*tb)
        qp = (C.c_char_p * max(n, 1))(*qb)
        tl = np.asarray([len(b) for b in tb], dtype=np.uint64)
        ql = np.asarray([len(b) for b in qb], dtype=np.uint64)
        out = C.c_void_p()
        check(lib().sg_align_pairs_v(self._h, C.cast(tp, C.c_void_p), tl.ctypes.data, C.cast(qp, C.c_void_p), ql.ctypes.data, n,
                                     SG_FLAG_DISTANCE_ONLY if distance_only else 0, C.byref(out)))
        return Result(out)

    def align_candidates_v(self, reads: Sequence[Union[str, bytes]], cand_start: Sequence[int], cand_read: Sequence[int],
                           distance_only: bool = False) -> Result:
        rb = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        rp = (C.c_char_p * max(len(rb), 1))(*rb)
        rl = np.asarray([len(b) for b in rb], dtype=np.uint64)
        cs = np.ascontiguousarray(cand_start, dtype=np.uint64)
        cr = np.ascontiguousarray(cand_read, dtype=np.uint32)
        out = C.c_void_p()
        check(lib().sg_align_candidates_v(self._h, C.cast(rp, C.c_void_p), rl.ctypes.data, len(rb), cs.ctypes.data, cr.ctypes.data,
                                          len(cs), SG_FLAG_DISTANCE_ONLY if distance_only else 0, C.byref(out)))
        return Result(out)

    def set_reference(self, genome: Union[str, bytes, np.ndarray]) -> None:
        g = genome.encode() if isinstance(genome, str) else genome
        check(lib().sg_set_reference(self._h, _ptr(g), len(g)))

    def align_candidates(self, reads: Sequence[str], cand_start: Sequence[int], cand_read: Sequence[int],
                         distance_only: bool = False) -> Result:
        rblob, roff = _blob(reads)
        cs = np.ascontiguousarray(cand_start, dtype=np.uint64)
        cr = np.ascontiguousarray(cand_read, dtype=np.uint32)
        out = C.c_void_p()
        check(lib().sg_align_candidates(self._h, _ptr(rblob), roff.ctypes.data, len(reads), cs.ctypes.data, cr.ctypes.data,
                                        len(cs), SG_FLAG_DISTANCE_ONLY if distance_only else 0, C.byref(out)))
        return Result(out)


def _ptr(buf) -> int:
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if isinstance(buf, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(buf)) if isinstance(buf, bytearray) else C.c_char_p(buf), C.c_void_p).value or 0
    if hasattr(buf, "data_ptr"):  # a (pinned) torch tensor
        return int(buf.data_ptr())
    raise TypeError(type(buf))


_default: dict = {}


def _aligner(W: int, O: Optional[int] = None) -> Aligner:
    key = (W, min(W // 2 + 1, W - 1) if O is None else O)
    if key not in _default:
        _default[key] = Aligner(W=W, O=key[1])
    return _default[key]


def align_all(a, b, W: int = 64, O: Optional[int] = None) -> List[Alignment]:
    """The reference's two overloads in one function.

    ``align_all(texts, queries)``   -- unstructured interface (src/genasm_gpu.hpp:8): lists of strings.
    ``align_all(reference, reads)`` -- read-mapping interface (src/genasm_gpu.hpp:7): a ``Genome`` and a list of
    ``Read``; one result per (read, location), read-major.
    """
    al = _aligner(W, O)
    if isinstance(a, Genome):
        reads: Sequence[Read] = b
        al.set_reference(a.content)
        cs = [loc.start_in_reference for r in reads for loc in r.locations]
        cr = [i for i, r in enumerate(reads) for _ in r.locations]
        return al.align_candidates([r.content for r in reads], cs, cr).alignments()
    return al.align_pairs(a, b).alignments()
