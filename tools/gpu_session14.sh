#!/bin/bash
TAG=${1:-s14}
mkdir -p gpurun_out
SG_MIN_BATCH_UNITS=64 timeout 900 compute-sanitizer --tool memcheck python tools/memcheck_windows.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/${TAG}_memcheck.log
