#!/bin/bash
# session 27: full verification of HEAD + refreshed window sweep
TAG=${1:-s27}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench.json
timeout 900 python tools/bench_configs.py windows --pairs 1000000 > gpurun_out/${TAG}_windows.jsonl 2> gpurun_out/${TAG}_windows.err; echo "windows rc=$?"
wc -l gpurun_out/${TAG}_windows.jsonl
