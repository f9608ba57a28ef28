#!/bin/bash
# session 11: A/B of SG_DELTA_GATHER=2, accuracy sweep over the window configurations
TAG=${1:-s11}
mkdir -p gpurun_out
for v in base g2; do
  SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_$v.so timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
done
timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_g2.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_mixed or long_reads" 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest_g2.log
timeout 900 python tools/accuracy_sweep.py --pairs 200 --len 2000 > gpurun_out/${TAG}_accuracy.jsonl 2> gpurun_out/${TAG}_accuracy.err; echo "accuracy rc=$?"
tail -2 gpurun_out/${TAG}_accuracy.err; wc -l gpurun_out/${TAG}_accuracy.jsonl
