#!/bin/bash
# One GPU-box session, parameterised:  tools/gpu_session.sh <tag> <step> [<step> ...]
# Every step writes gpurun_out/<tag>_<step>.* ; steps (run in the order given):
#   info          GPUs, CPUs, NUMA, PCIe topology
#   pytest        the GPU parity tests (pytest -m gpu)
#   smoke         __graft_entry__.smoke()
#   bench         python bench.py                       (N=1; under torchrun when GPUS>1)
#   benchref      python bench.py --impl reference
#   launches      ncu launch list of one bench step     (gpu__time_duration.sum, --clock-control none)
#   ncu:<kernel>  ncu --set full capture of <kernel> during a short bench (tools/ncu_capture.sh)
#   probe         build/sg_e2e_probe over GPUS GPUs, both layouts (host ceilings + ingest policies + breakdown)
#   configs       tools/bench_configs.py all
#   cmd:<...>     any shell command (quote it), output to gpurun_out/<tag>_cmd<k>.log
# Environment: GPUS (default 1), STEPS / WARMUP for bench (defaults 5 / 3), PROBE_ARGS for the probe.
TAG=${1:-s}; shift
GPUS=${GPUS:-1}
mkdir -p gpurun_out
k=0
bench() {  # $1 = extra args, $2 = output stem
  if [ "$GPUS" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $GPUS \
      --steps ${STEPS:-5} --warmup ${WARMUP:-3} $1 > gpurun_out/$2.json 2> gpurun_out/$2.err
  else
    python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} $1 > gpurun_out/$2.json 2> gpurun_out/$2.err
  fi
  echo "$2 rc=$?"; cut -c1-400 gpurun_out/$2.json
}
for step in "$@"; do
  case "$step" in
    info)
      { nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.max.sm,power.limit --format=csv; nvidia-smi topo -m; lscpu | head -30;
        cat /sys/devices/system/node/node*/cpulist; free -g; } > gpurun_out/${TAG}_info.txt 2>&1 ;;
    pytest)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
      tail -4 gpurun_out/${TAG}_pytest.log ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log ;;
    bench) timeout 1200 bash -c "$(declare -f bench); GPUS=$GPUS STEPS=${STEPS:-5} WARMUP=${WARMUP:-3} bench '${BENCH_ARGS:-}' ${TAG}_bench" ;;
    benchref) timeout 600 bash -c "$(declare -f bench); GPUS=$GPUS STEPS=${STEPS:-5} WARMUP=${WARMUP:-3} bench '--impl reference' ${TAG}_benchref" ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${TAG}_launches.log 2>&1; echo "launches rc=$?" ;;
    ncu:*) timeout 900 bash tools/ncu_capture.sh ${TAG}_${step#ncu:} ${step#ncu:} ;;
    probe)
      for layout in procs threads; do
        timeout 900 ./build/sg_e2e_probe --gpus $GPUS --layout $layout ${PROBE_ARGS:-} > gpurun_out/${TAG}_probe_$layout.jsonl 2> gpurun_out/${TAG}_probe_$layout.err
        echo "probe $layout rc=$?"; cut -c1-600 gpurun_out/${TAG}_probe_$layout.jsonl
      done ;;
    configs) for what in short mapping sweep; do timeout 2400 python tools/bench_configs.py $what >> gpurun_out/${TAG}_configs.jsonl 2>> gpurun_out/${TAG}_configs.err; echo "configs $what rc=$?"; done ;;
    cmd:*) k=$((k+1)); timeout 2400 bash -c "${step#cmd:}" > gpurun_out/${TAG}_cmd$k.log 2>&1; echo "cmd$k rc=$?"; tail -5 gpurun_out/${TAG}_cmd$k.log ;;
    *) echo "unknown step $step" ;;
  esac
done
