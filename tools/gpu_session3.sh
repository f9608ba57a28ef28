#!/bin/bash
# adaptive-ingest end-to-end measurements after the fix
TAG=${1:-s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python tools/e2e_sweep.py 524288 > gpurun_out/${TAG}_e2e_sweep.log 2>&1
cat gpurun_out/${TAG}_e2e_sweep.log | tail -14
SG_DEBUG=1 timeout 300 python tools/e2e_debug.py 2>&1 | grep "call\|py" | tail -4
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --e2e-pairs 1000000 --no-cpu-baseline > gpurun_out/${TAG}_bench_e2e1m.json 2>> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${TAG}_bench_e2e1m.json'));print(d['e2e'])"
