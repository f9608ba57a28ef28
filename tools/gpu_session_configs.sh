#!/bin/bash
TAG=${1:-c}
mkdir -p gpurun_out
timeout 500 python tools/bench_configs.py short > gpurun_out/${TAG}_short.jsonl 2> gpurun_out/${TAG}_short.err; echo "short rc=$?"
timeout 700 python tools/bench_configs.py mapping > gpurun_out/${TAG}_mapping.jsonl 2> gpurun_out/${TAG}_mapping.err; echo "mapping rc=$?"
timeout 900 python tools/bench_configs.py sweep > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"
wc -l gpurun_out/${TAG}_*.jsonl; tail -2 gpurun_out/${TAG}_*.err
