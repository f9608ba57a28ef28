#!/bin/bash
# A/B builds of the alignment kernel: tools/build_variants.sh name "-DSG_DELTA_FMA=0 ..." [name flags ...]
# -> scrooge_b200/lib/variants/libscrooge_b200_<name>.so (use with SG_LIB=... tools/kernel_time.py)
set -e
cd "$(dirname "$0")/.."
make -s all
mkdir -p scrooge_b200/lib/variants build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ \
      -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas -Iinclude $flags -c scrooge_b200/csrc/sg_device_api.cu -o build/variants/dev_$name.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -Xcompiler -fopenmp \
      -o scrooge_b200/lib/variants/libscrooge_b200_$name.so build/variants/dev_$name.o build/sg_host_api.o build/genasm_gpu.o \
      build/sg_host_pack.o build/sg_host_render.o build/sg_io.o -lcudart -lgomp
  echo "built variant $name ($flags)"
done
