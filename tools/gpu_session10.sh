#!/bin/bash
# session 10: A/B of the window-setup variants (SG_DELTA_GATHER, SG_DELTA_OFFMUL), parity of the default build
TAG=${1:-s10}
mkdir -p gpurun_out
for v in base g o; do
  SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_$v.so timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
done
timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
timeout 200 python tools/kernel_time.py short_150bp 10000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_base.so timeout 200 python tools/kernel_time.py short_150bp 10000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
