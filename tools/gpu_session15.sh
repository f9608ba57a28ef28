#!/bin/bash
# session 15: general kernel v2 (word-blocked DC loop, stream traceback + RLE after the walk): parity, window sweep
TAG=${1:-s15}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "window or generic" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
SG_GENERIC=1 timeout 300 python tools/kernel_time.py long_10kbp 200000,1000000 2>&1 | tail -2 | tee gpurun_out/${TAG}_generic_6433.log
timeout 600 python tools/bench_configs.py windows --pairs 200000 > gpurun_out/${TAG}_windows.jsonl 2> gpurun_out/${TAG}_windows.err; echo "windows rc=$?"
SG_MIN_BATCH_UNITS=64 timeout 600 compute-sanitizer --tool memcheck python tools/memcheck_windows.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/${TAG}_memcheck.log
