#!/bin/bash
# session 12: ncu captures of the final tuned kernel and of the general kernel, launch list of a bench run
TAG=${1:-s12}
mkdir -p gpurun_out
bash tools/ncu_capture.sh ${TAG}_delta genasm_delta_kernel 303104 > gpurun_out/${TAG}_cap1.log 2>&1
SG_GENERIC=1 ncu --set full --clock-control none --import-source on -k regex:genasm_generic_kernel -c 1 -f -o gpurun_out/${TAG}_generic \
    python tools/kernel_time.py long_10kbp 200000 > gpurun_out/${TAG}_generic_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_generic.ncu-rep --page raw --csv > gpurun_out/${TAG}_generic_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_generic.ncu-rep --page source --csv > gpurun_out/${TAG}_generic_src.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}
