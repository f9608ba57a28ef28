#!/bin/bash
TAG=${1:-s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
SG_DEBUG=1 timeout 600 python tools/e2e_workloads.py short_150bp 10000000 2>&1 | grep "call\|aligns" | tail -2
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['parity'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt; head -7 gpurun_out/${TAG}_launches.txt
