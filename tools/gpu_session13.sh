#!/bin/bash
# session 13: sensitivity of the tuned kernel to its occupancy (persistent CTAs per SM, 4 warps each)
TAG=${1:-s13}
mkdir -p gpurun_out
for c in 3 4 5 6; do
  echo "SG_DELTA_CTAS=$c" | tee -a gpurun_out/${TAG}_occupancy.log
  SG_DELTA_CTAS=$c timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_occupancy.log
done
