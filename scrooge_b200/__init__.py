"""scrooge_b200 -- B200-native (sm_100a) implementation of Scrooge's windowed GenASM aligner.

Only what the hot path needs: csrc/ (CUDA kernels + the C ABI of include/scrooge_b200.h + the C++ drop-in
genasm_gpu::align_all), and this Python mirror of the reference's library interface.
"""
from ._lib import ScroogeError, bench_lib, build, lib  # noqa: F401
from .api import Aligner, Alignment, CandidateLocation, Genome, Read, Result, align_all  # noqa: F401

__all__ = ["Aligner", "Alignment", "CandidateLocation", "Genome", "Read", "Result", "align_all", "ScroogeError",
           "build", "lib"]
