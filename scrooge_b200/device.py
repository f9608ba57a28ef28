"""Device-level entry points (sg_dev_*) on torch tensors.

torch is used only as plumbing: device memory, the current stream and CUDA events.  Every function here
launches the library's own sm_100a kernels on ``torch.cuda.current_stream()`` so that ``torch.cuda.Event``
timing brackets them.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from ._lib import SG_FLAG_DISTANCE_ONLY, SG_FLAG_RUN_WORDS, bench_check, bench_lib, check, lib

U64_MAX = (1 << 64) - 1


def _stream() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else int(t.data_ptr())


def pack_2bit(ascii_dev: torch.Tensor, n_bases: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """ASCII bytes (uint8, device) -> packed 2-bit words (int32 storage, device) and the first-bad-position cell
    (int64 storage; -1 == none).  Replaces reference single_ascii_to_twobit_string (src/genasm_gpu.cu:640-685)."""
    assert ascii_dev.is_cuda and ascii_dev.dtype == torch.uint8 and ascii_dev.is_contiguous()
    n = ascii_dev.numel() if n_bases is None else n_bases
    words = int(lib().sg_packed_words(n))
    packed = torch.empty(words, dtype=torch.int32, device=ascii_dev.device)
    bad = torch.full((1,), -1, dtype=torch.int64, device=ascii_dev.device)  # == UINT64_MAX
    check(lib().sg_dev_pack_2bit(_p(ascii_dev), n, _p(packed), _p(bad), _stream()))
    return packed, bad


@dataclass
class AlignOut:
    edit: torch.Tensor          # int64 [n]
    ref_consumed: torch.Tensor  # int64 storage of uint64 [n]
    nruns: torch.Tensor         # int32 storage of uint32 [n]
    status: torch.Tensor        # uint8 [n]
    dc_entries: torch.Tensor    # int64 storage of uint64 [n]: sum over windows of (d_w+1)*(n_w+1)
    windows: torch.Tensor       # int32 storage of uint32 [n]: windows per alignment


class DeviceAligner:
    """Preallocated outputs + one sg_dev_align launch per call (inputs resident in HBM)."""

    def __init__(self, W: int, n: int, device: torch.device, slab_bytes: int = 0, O: Optional[int] = None):
        self.W, self.n, self.device = W, n, device
        self.O = min(W // 2 + 1, W - 1) if O is None else O
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        self.out = AlignOut(
            edit=torch.empty(n, dtype=torch.int64, device=device),
            ref_consumed=torch.empty(n, dtype=torch.int64, device=device),
            nruns=torch.empty(n, dtype=torch.int32, device=device),
            status=torch.empty(n, dtype=torch.uint8, device=device),
            dc_entries=torch.empty(n, dtype=torch.int64, device=device),
            windows=torch.empty(n, dtype=torch.int32, device=device),
        )
        self.slab = torch.empty(max(slab_bytes, 16), dtype=torch.uint8, device=device) if slab_bytes else None
        self.run_off = torch.empty(n + 1, dtype=torch.int64, device=device)
        self.scan_tmp = torch.empty(int(lib().sg_scan_tmp_bytes(n)), dtype=torch.uint8, device=device)

    def align(self, text: torch.Tensor, text_start: torch.Tensor, text_len: torch.Tensor, query: torch.Tensor,
              query_start: torch.Tensor, query_len: torch.Tensor, slab_off: Optional[torch.Tensor] = None,
              distance_only: bool = False, stats: bool = True, run_words: bool = False) -> AlignOut:
        """stats=False: the optional work counters (dc_entries, windows) are not requested, as in the host API's launches;
        the kernel then skips the window-distance bookkeeping they need.  run_words=True: SG_FLAG_RUN_WORDS (runs stored as
        whole words; every slab offset must be a multiple of 4)."""
        flags = (SG_FLAG_DISTANCE_ONLY if distance_only else 0) | (SG_FLAG_RUN_WORDS if run_words else 0)
        o = self.out
        if not stats:
            check(lib().sg_dev_align_wo(self.W, self.O, _p(text), _p(text_start), _p(text_len), _p(query), _p(query_start), _p(query_len),
                                        self.n, flags, _p(self.slab), _p(slab_off), _p(self.counter), _p(o.edit),
                                        _p(o.ref_consumed), _p(o.nruns), _p(o.status), None, None, _stream()))
            return o
        check(lib().sg_dev_align_wo(self.W, self.O, _p(text), _p(text_start), _p(text_len), _p(query), _p(query_start), _p(query_len),
                                    self.n, flags, _p(self.slab), _p(slab_off), _p(self.counter), _p(o.edit),
                                    _p(o.ref_consumed), _p(o.nruns), _p(o.status), _p(o.dc_entries), _p(o.windows), _stream()))
        return o

    def compact(self, slab_off: torch.Tensor, runs: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """scan + gather; returns (run_off [n+1], dense runs).  `runs` may be a preallocated buffer."""
        o = self.out
        check(lib().sg_dev_scan_runs(_p(o.nruns), self.n, _p(self.run_off), _p(self.scan_tmp), _stream()))
        if runs is None:
            total = int(self.run_off[-1].item())
            runs = torch.empty(max(total, 1), dtype=torch.uint8, device=self.device)
        # hint: the slab capacity per alignment (an upper bound of its runs); short alignments are gathered by four lanes
        hint = int(self.slab.numel() // max(self.n, 1)) if self.slab is not None else 0
        check(lib().sg_dev_gather_runs_sized(_p(self.slab), _p(slab_off), _p(o.nruns), _p(self.run_off), self.n, _p(runs), hint,
                                             _stream()))
        return self.run_off, runs


def check_runs(runs: torch.Tensor, run_off: torch.Tensor, query_len: torch.Tensor, out: "AlignOut", W: int,
               O: Optional[int] = None) -> int:
    """Number of alignments whose compacted runs contradict their query length, consumed reference prefix or edit
    distance (sg_dev_check_runs: the sequence-independent validateCigarString properties on the whole batch)."""
    n = query_len.numel()
    bad = torch.zeros(1, dtype=torch.int64, device=runs.device)
    bench_check(bench_lib().sg_dev_check_runs(_p(runs), _p(run_off), n, _p(query_len), _p(out.edit), _p(out.ref_consumed),
                                  W - (min(W // 2 + 1, W - 1) if O is None else O), _p(bad), _stream()))
    return int(bad.item())


def align_geometry(W: int, O: Optional[int] = None) -> Tuple[int, int, int]:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().sg_dev_align_geometry_wo(W, min(W // 2 + 1, W - 1) if O is None else O, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value  # warps per SM, smem bytes per warp, SMs


def int32_peak(kind: int = 2, ms: float = 50.0) -> float:
    g = C.c_double()
    bench_check(bench_lib().sg_dev_int32_peak(kind, ms, C.byref(g)))
    return g.value


def synth_pairs_device(seed: int, first_pair: int, n_pairs: int, read_len: int, err: float, ratio: Tuple[int, int, int],
                       slack: int, device: torch.device):
    """Generates pairs on the device.  Returns (text uint8 [n, stride], text_len int64 [n], reads uint8 [n, L])."""
    stride = int(bench_lib().sg_synth_text_stride(read_len, slack))
    text = torch.empty((n_pairs, stride), dtype=torch.uint8, device=device)
    tlen = torch.empty(n_pairs, dtype=torch.int64, device=device)
    reads = torch.empty((n_pairs, read_len), dtype=torch.uint8, device=device)
    bench_check(bench_lib().sg_dev_synth_pairs(seed, first_pair, n_pairs, read_len, float(err), ratio[0], ratio[1], ratio[2], slack,
                                   _p(text), stride, _p(tlen), _p(reads), _stream()))
    return text, tlen, reads
