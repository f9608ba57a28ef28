#!/usr/bin/env python
"""Turns ncu exports into the small text summaries committed under profiles/.
  launches: ncu_summarize.py launches <launches.csv>
  kernel:   ncu_summarize.py kernel <raw.csv>      (from `ncu -i x.ncu-rep --page raw --csv`)"""
import collections
import csv
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(data)} launches, {tot:.3f} ms total (cold-cache, serialised: compare shares, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:70s} n={v[0]:3d} total_ms={v[1]:10.3f} avg_ms={v[1]/v[0]:9.3f} share={v[1]/tot:6.3f}")


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]


def kernel(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    for d in rows[2:]:
        print("# kernel:", d[h.index("Kernel Name")] if "Kernel Name" in h else "?")
        for k in KEYS + [x for x in h if "issue_stalled" in x and "per_issue_active" in x]:
            if k in h:
                i = h.index(k)
                print(f"{k:82s} {u[i]:16s} {d[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
