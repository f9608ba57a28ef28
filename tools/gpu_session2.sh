#!/bin/bash
# kernel A/B variants + adaptive-ingest end-to-end measurements
TAG=${1:-s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for v in v0 v1 v2 v3; do
  SG_LIB=$PWD/scrooge_b200/lib/variants/libscrooge_b200_$v.so timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
done
timeout 200 python tools/kernel_time.py long_10kbp 1000000 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_variants.log
timeout 900 python tools/e2e_sweep.py 524288 > gpurun_out/${TAG}_e2e_sweep.log 2>&1
cat gpurun_out/${TAG}_e2e_sweep.log | tail -14
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
