#!/usr/bin/env python
"""Accuracy of the windowed heuristic against the exact answer, per window configuration (W, O) -- the axes of the
reference's accuracy study (scripts/profile.py:588-645 cpu_accuracy_sweep_wo / cpu_accuracy_sweep_o), run on the GPU with
the window configuration chosen at run time (the reference rebuilds its binary per point).

For every (W, O): the edit distance the aligner returns against the exact semi-global edit distance (whole read against
the best prefix of the text, start anchored -- computed here by a plain dynamic programme in numpy), and the affine-gap
score of the returned CIGAR as the reference scores it (src/cpu_baseline.cpp:694-725: +mat per '=', -sub per 'X',
-gapo once per maximal run of I/D entries, -gape per gap base; scoring 2,4,4,2 as scripts/profile.py:592-597).

    python tools/accuracy_sweep.py [--pairs 200] [--len 2000] [--err 0.10,0.15] > profiles/...jsonl

One JSON object per line on stdout.
"""
import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import scrooge_b200  # noqa: E402
from scrooge_b200 import synth  # noqa: E402

SCORING = dict(mat=2, sub=4, gapo=4, gape=2)   # scripts/profile.py:592-597
RUN = re.compile(r"(\d+)([=XID])")


def exact_distance(text: np.ndarray, read: np.ndarray) -> int:
    """min over text prefixes of the edit distance between the whole read and that prefix."""
    n = text.shape[0]
    ar = np.arange(n + 1, dtype=np.int64)
    prev = ar.copy()                      # zero read characters against j text characters: j deletions
    for i in range(read.shape[0]):
        tmp = np.empty(n + 1, dtype=np.int64)
        tmp[0] = i + 1
        np.minimum(prev[1:] + 1, prev[:-1] + (text != read[i]), out=tmp[1:])
        prev = ar + np.minimum.accumulate(tmp - ar)   # deletions of text characters: cur[j] = min_k tmp[k] + (j - k)
    return int(prev.min())


def cigar_score(cigar: str) -> int:
    """src/cpu_baseline.cpp:694-725"""
    score, was_gap = 0, False
    for cnt, op in RUN.findall(cigar):
        c = int(cnt)
        if op == "=":
            score += c * SCORING["mat"]
            was_gap = False
        elif op == "X":
            score -= c * SCORING["sub"]
            was_gap = False
        else:
            if not was_gap:
                score -= SCORING["gapo"]
            score -= c * SCORING["gape"]
            was_gap = True
    return score


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200)
    ap.add_argument("--len", type=int, default=2000)
    ap.add_argument("--err", default="0.10,0.15")
    args = ap.parse_args()
    points = [(W, min(W // 2 + 1, W - 1)) for W in (16, 32, 48, 64, 96, 128, 192, 256)] + [(64, O) for O in (1, 8, 16, 24, 40, 48, 56, 63)] + \
             [(64, 0)] + [(32, O) for O in (0, 8, 24)] + [(128, O) for O in (1, 32, 80, 100, 120)]
    for err in (float(x) for x in args.err.split(",")):
        wl = synth.Workload(f"accuracy_{args.len}bp_{int(err * 100)}pct", args.len, err, synth.PACBIO, 64, synth.BASE_SEED + 6)
        text, tlen, reads = synth.pairs_host(wl, 0, args.pairs)
        T, Q = synth.pairs_as_strings(text, tlen, reads)
        exact = np.array([exact_distance(text[k, : int(tlen[k])], reads[k]) for k in range(args.pairs)], dtype=np.int64)
        for W, O in points:
            al = scrooge_b200.Aligner(W=W, O=O, n_gpus=1)
            res = al.align_pairs(T, Q)
            ed = np.asarray(res.edit_distances, dtype=np.int64)
            assert (ed >= exact).all(), "a heuristic distance below the exact one"
            scores = np.array([cigar_score(c) for c in res.cigars()], dtype=np.int64)
            print(json.dumps({"workload": wl.name, "pairs": args.pairs, "read_len": args.len, "error_rate": err, "W": W, "O": O,
                              "exact_mean_distance": float(exact.mean()), "mean_distance": float(ed.mean()),
                              "optimal_fraction": float((ed == exact).mean()),
                              "mean_excess_edits": float((ed - exact).mean()),
                              "mean_relative_excess": float(((ed - exact) / np.maximum(exact, 1)).mean()),
                              "max_excess_edits": int((ed - exact).max()),
                              "mean_affine_score": float(scores.mean()), "scoring": SCORING}), flush=True)
            al.close()


if __name__ == "__main__":
    main()
