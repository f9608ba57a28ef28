#!/bin/bash
# session 21: general kernel with op planes in global memory (L2-resident) for W - O > 32: parity in both modes, sweep, memcheck
TAG=${1:-s21}
mkdir -p gpurun_out
for mode in auto global smem; do
  env $( [ $mode != auto ] && echo SG_GENERIC_PLANES=$mode ) timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or generic" > gpurun_out/${TAG}_pytest_$mode.log 2>&1; echo "pytest $mode rc=$?"
  tail -1 gpurun_out/${TAG}_pytest_$mode.log
done
timeout 600 python tools/bench_configs.py windows --pairs 1000000 > gpurun_out/${TAG}_windows.jsonl 2> gpurun_out/${TAG}_windows.err; echo "windows rc=$?"; tail -2 gpurun_out/${TAG}_windows.err
SG_GENERIC_PLANES=global SG_MIN_BATCH_UNITS=64 timeout 600 compute-sanitizer --tool memcheck python tools/memcheck_windows.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -2 gpurun_out/${TAG}_memcheck.log
