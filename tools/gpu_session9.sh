#!/bin/bash
# session 9: run-time window configurations (generic kernel) -- GPU parity tests, the window sweep, the default bench
TAG=${1:-s9}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
SG_DEBUG=1 timeout 600 python tools/bench_configs.py windows --pairs 200000 > gpurun_out/${TAG}_windows.jsonl 2> gpurun_out/${TAG}_windows.err; echo "windows rc=$?"
tail -3 gpurun_out/${TAG}_windows.err
SG_GENERIC=1 timeout 300 python tools/kernel_time.py long_10kbp 200000 2>&1 | tail -1 | tee gpurun_out/${TAG}_generic_6433.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-600
