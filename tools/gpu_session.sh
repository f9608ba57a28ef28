#!/bin/bash
# One GPU-box session: parity tests, the bench lines, ncu capture of the alignment kernel, host-time breakdown of the
# end-to-end path.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --impl reference > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench_ref.json
timeout 600 bash tools/ncu_capture.sh ${TAG}_delta genasm_delta_kernel
SG_DEBUG=1 timeout 300 python tools/e2e_debug.py > gpurun_out/${TAG}_e2e_debug.log 2>&1
tail -30 gpurun_out/${TAG}_e2e_debug.log
