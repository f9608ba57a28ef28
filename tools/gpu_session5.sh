#!/bin/bash
TAG=${1:-s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
SG_DEBUG=1 timeout 600 python tools/e2e_workloads.py short_150bp 10000000 2>&1 | grep "call\|aligns" | tail -4
SG_DEBUG=1 timeout 600 python tools/e2e_workloads.py long_10kbp 524288 2>&1 | grep "call\|aligns" | tail -3
timeout 300 ./build/sg_tests --unit_tests 2>&1 | tail -3
