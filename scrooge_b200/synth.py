"""Synthetic workloads of BASELINE.json's configs (generator: csrc/sg_synth.h; shapes: SURVEY.md section 8d)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ._lib import bench_check, bench_lib, check, lib

PACBIO = (6, 50, 54)     # sub:ins:del, reference DATASETS.md:51 (--difference-ratio)
ILLUMINA = (90, 5, 5)    # substitution dominated (our choice; the reference does not specify one)
UNIFORM = (1, 1, 1)
BASE_SEED = 0x5C2006E


@dataclass
class Workload:
    name: str
    read_len: int
    err: float
    ratio: Tuple[int, int, int]
    W: int
    seed: int
    slack: int = 64
    O: Optional[int] = None   # window overlap; None = min(W//2+1, W-1) (reference scripts/profile.py:78)

    @property
    def overlap(self) -> int:
        return min(self.W // 2 + 1, self.W - 1) if self.O is None else self.O


WORKLOADS = {
    # BASELINE.json configs[0..2]
    "library_example_100bp": Workload("library_example_100bp", 100, 0.05, UNIFORM, 64, BASE_SEED + 1),
    "short_150bp": Workload("short_150bp", 150, 0.05, ILLUMINA, 64, BASE_SEED + 2),
    "long_10kbp": Workload("long_10kbp", 10000, 0.10, PACBIO, 64, BASE_SEED + 3),
    # the reference's short-read window setting (README.md:208: W=32, O=17) on the same pairs as short_150bp
    "short_150bp_w32": Workload("short_150bp_w32", 150, 0.05, ILLUMINA, 32, BASE_SEED + 2),
    # BASELINE.json configs[4] sweep points (1-100 kbp at 5/10/15 %)
    "sweep_1kbp_5": Workload("sweep_1kbp_5", 1000, 0.05, PACBIO, 64, BASE_SEED + 5),
    "sweep_1kbp_15": Workload("sweep_1kbp_15", 1000, 0.15, PACBIO, 64, BASE_SEED + 5),
    "sweep_10kbp_5": Workload("sweep_10kbp_5", 10000, 0.05, PACBIO, 64, BASE_SEED + 5),
    "sweep_10kbp_15": Workload("sweep_10kbp_15", 10000, 0.15, PACBIO, 64, BASE_SEED + 5),
    "sweep_100kbp_10": Workload("sweep_100kbp_10", 100000, 0.10, PACBIO, 64, BASE_SEED + 5),
}


def text_stride(read_len: int, slack: int = 64) -> int:
    return int(bench_lib().sg_synth_text_stride(read_len, slack))


def pairs_host(wl: Workload, first_pair: int, n_pairs: int):
    """Returns (text uint8 [n, stride], text_len uint64 [n], reads uint8 [n, L]) generated on the host."""
    stride = text_stride(wl.read_len, wl.slack)
    text = np.empty((n_pairs, stride), dtype=np.uint8)
    tlen = np.empty(n_pairs, dtype=np.uint64)
    reads = np.empty((n_pairs, wl.read_len), dtype=np.uint8)
    bench_check(bench_lib().sg_synth_pairs_host(wl.seed, first_pair, n_pairs, wl.read_len, float(wl.err), wl.ratio[0], wl.ratio[1],
                                    wl.ratio[2], wl.slack, text.ctypes.data, stride, tlen.ctypes.data, reads.ctypes.data))
    return text, tlen, reads


def pairs_as_blobs(text: np.ndarray, tlen: np.ndarray, reads: np.ndarray):
    """Concatenated ASCII blobs + offsets (the C ABI's input form) from the strided generator output."""
    n, L = reads.shape
    toff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(tlen, out=toff[1:])
    qoff = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    mask = np.arange(text.shape[1], dtype=np.uint64)[None, :] < tlen[:, None]
    tblob = text[mask]
    return np.ascontiguousarray(tblob), toff, np.ascontiguousarray(reads.reshape(-1)), qoff


def pairs_as_strings(text: np.ndarray, tlen: np.ndarray, reads: np.ndarray) -> Tuple[List[str], List[str]]:
    texts = [text[i, : int(tlen[i])].tobytes().decode() for i in range(len(tlen))]
    queries = [reads[i].tobytes().decode() for i in range(reads.shape[0])]
    return texts, queries


def mapping_host(seed: int, genome_len: int, n_reads: int, read_len: int, err: float = 0.10, ratio: Tuple[int, int, int] = PACBIO):
    """A small read-mapping case generated on the host (BASELINE.json configs[3] in miniature): returns (genome uint8
    [genome_len], reads uint8 [n, L], true start uint64 [n])."""
    genome = np.empty(genome_len, dtype=np.uint8)
    bench_check(bench_lib().sg_synth_genome(seed, 0, genome_len, genome.ctypes.data, None, None))
    reads = np.empty((n_reads, read_len), dtype=np.uint8)
    pos = np.empty(n_reads, dtype=np.uint64)
    bench_check(bench_lib().sg_synth_reads(seed + 1, 0, n_reads, read_len, float(err), ratio[0], ratio[1], ratio[2], genome.ctypes.data,
                                           genome_len, reads.ctypes.data, pos.ctypes.data, 0, None))
    return genome, reads, pos
