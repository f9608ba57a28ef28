#!/bin/bash
TAG=${1:-s}
mkdir -p gpurun_out
for v in v0 v1 v2; do
  SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_$v.so timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
done
timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
