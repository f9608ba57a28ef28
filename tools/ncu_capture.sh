#!/bin/bash
# One `ncu --set full` capture of the alignment kernel on the benchmark workload (B200_PROFILING.md recipe) + the launch
# list of a short bench run.  usage (on the GPU box): tools/ncu_capture.sh <tag> [kernel-regex] [pairs]
# Writes gpurun_out/<tag>.ncu-rep, <tag>_raw.csv, <tag>_src.csv, <tag>_launches.csv.  Numbers printed by bench.py under
# ncu are never bench values.
set -u
TAG=${1:-prof}
KRE=${2:-genasm_delta_kernel}
PAIRS=${3:-303104}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$KRE -c 1 -f -o gpurun_out/$TAG \
    python bench.py --pairs $PAIRS --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out | grep $TAG
