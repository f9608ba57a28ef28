#!/usr/bin/env python
"""Renders the JSON lines of tools/bench_configs.py as the markdown table committed under profiles/.
usage: configs_md.py out.md in1.jsonl [in2.jsonl ...]"""
import json
import sys

rows = [json.loads(l) for f in sys.argv[2:] for l in open(f) if l.strip()]
with open(sys.argv[1], "w") as o:
    o.write("# Round 1 -- BASELINE.json configs[1], [3], [4] on one B200 (device-timed; `tools/bench_configs.py`, delta kernel)\n\n"
            "Every point was checked bit-exact (edit distance, CIGAR, consumed reference prefix) against the oracle on a sample of its own inputs.\n"
            "`frac` = INT32 ops of the delta formulation (windows x W columns x 20 ops at W=64, 10 at W=32) / kernel time / measured LOP3+SHF peak;\n"
            "`ref x` = the reference formulation's count (14 ops per R[d][i] entry at W=64, 7 at W=32; SURVEY 8d) over the same time, as a multiple of that peak.\n\n"
            "| workload | L | err | mode | alignments | M aln/s (kernel) | M aln/s (ingest+align+compact) | GCUPS (kernel) | windows/aln | DC entries/aln | frac | ref x | parity |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for d in rows:
        n = d.get("pairs", d.get("alignments"))
        o.write(f"| {d['workload']} | {d['read_len']} | {d['error_rate']} | {d.get('mode', 'full_cigar')} | {n} | "
                f"{d['alignments_per_s_kernel'] / 1e6:.3f} | {d['alignments_per_s_step'] / 1e6:.3f} | {d['gcups_kernel']:.0f} | "
                f"{d.get('windows_per_alignment', 0):.1f} | {d['dc_entries_per_alignment']:.0f} | {d['int32_frac']:.3f} | "
                f"{d['reference_formulation_ratio']:.2f} | {'bit-exact' if d['parity']['bit_exact'] else 'MISMATCH'} ({d['parity']['checked']}) |\n")
