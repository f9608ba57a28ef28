#!/bin/bash
# multi-GPU session: the multi-device context test + bench.py under torchrun.  usage: gpu_session_multi.sh <tag> <N>
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt; nproc >> gpurun_out/${TAG}_gpus.txt; free -g >> gpurun_out/${TAG}_gpus.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
