# Builds the product library (sm_100a only) in-tree:
#   scrooge_b200/lib/libscrooge_b200.so   C ABI of include/scrooge_b200.h + C++ drop-in genasm_gpu::align_all
#   scrooge_b200/lib/libscrooge_b200_rdc.a  the reference header's one exported kernel (src/genasm_gpu.hpp:9) as relocatable
#                                           device code, for callers built like the reference (-rdc=true) that launch it
#   scrooge_b200/lib/libscrooge_b200_bench.so  synthetic generators, peak probes, batch checker (bench.py / tests / apps only)
#   build/library_example, build/sg_tests  (C++ programs mirroring the reference's library_example / tests binaries)
# The oracle (test infrastructure) is built by oracle/Makefile.
NVCC ?= /usr/local/cuda/bin/nvcc
CCBIN ?= /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(CCBIN) -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas \
           -Xptxas -v -Iinclude
SRC := scrooge_b200/csrc
LIB := scrooge_b200/lib/libscrooge_b200.so
OBJS := build/sg_device_api.o build/sg_host_api.o build/genasm_gpu.o build/sg_host_pack.o build/sg_host_render.o build/sg_io.o
HDRS := $(wildcard $(SRC)/*.cuh $(SRC)/*.h include/*.h include/*.hpp)

RDC := scrooge_b200/lib/libscrooge_b200_rdc.a
# measurement / synthetic-data / checking helpers (include/scrooge_b200_bench.h): NOT in the product library
BENCHLIB := scrooge_b200/lib/libscrooge_b200_bench.so

all: $(LIB) $(RDC) $(BENCHLIB) build/library_example build/sg_tests build/sg_e2e_probe build/sg_variant_ab

$(BENCHLIB): build/sg_bench_api.o
	@mkdir -p scrooge_b200/lib
	$(NVCC) $(ARCH) -shared -ccbin $(CCBIN) -Xcompiler -fopenmp -o $@ build/sg_bench_api.o -lcudart -lgomp

$(RDC): $(SRC)/sg_dropin_rdc.cu
	@mkdir -p build scrooge_b200/lib
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -rdc=true -ccbin $(CCBIN) -Xcompiler -fPIC -c $< -o build/sg_dropin_rdc.o
	rm -f $@ && ar rcs $@ build/sg_dropin_rdc.o

build/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

# host-only code with ISA-specific function versions: plain g++
build/sg_host_pack.o: $(SRC)/sg_host_pack.cpp
	@mkdir -p build
	$(CCBIN) -O3 -std=c++17 -fPIC -fopenmp -Wall -c $< -o $@

build/sg_host_render.o: $(SRC)/sg_host_render.cpp
	@mkdir -p build
	$(CCBIN) -O3 -std=c++17 -fPIC -Wall -c $< -o $@

build/sg_io.o: $(SRC)/sg_io.cpp include/scrooge_io.hpp include/scrooge_types.hpp
	@mkdir -p build
	$(CCBIN) -O2 -std=c++17 -fPIC -Wall -c $< -o $@

build/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p scrooge_b200/lib
	$(NVCC) $(ARCH) -shared -ccbin $(CCBIN) -Xcompiler -fopenmp -o $@ $(OBJS) -lcudart -lgomp

build/library_example: examples/library_example.cpp $(LIB) $(HDRS)
	$(CCBIN) -O2 -std=c++17 -Iinclude -o $@ $< -Lscrooge_b200/lib -lscrooge_b200 -Wl,-rpath,'$$ORIGIN/../scrooge_b200/lib'

build/sg_tests: apps/sg_tests.cpp $(LIB) $(BENCHLIB) $(HDRS)
	$(CCBIN) -O2 -std=c++17 -Wall -Iinclude -o $@ $< -Lscrooge_b200/lib -lscrooge_b200 -lscrooge_b200_bench -Wl,-rpath,'$$ORIGIN/../scrooge_b200/lib'

# where the end-to-end time goes: host ceilings (copy-only, pack-only) and the ingest policies, 1..N GPUs (apps/sg_e2e_probe.cpp)
CUDA_HOME ?= /usr/local/cuda
build/sg_e2e_probe: apps/sg_e2e_probe.cpp $(LIB) $(BENCHLIB) $(HDRS)
	$(CCBIN) -O2 -std=c++17 -Wall -fopenmp -Iinclude -I$(CUDA_HOME)/include -o $@ $< -Lscrooge_b200/lib -lscrooge_b200 -lscrooge_b200_bench \
	    -L$(CUDA_HOME)/lib64 -lcudart -lpthread -Wl,-rpath,'$$ORIGIN/../scrooge_b200/lib'

# A/B of the kernel's run-emission variants on device-resident synthetic workloads (apps/sg_variant_ab.cpp)
build/sg_variant_ab: apps/sg_variant_ab.cpp $(LIB) $(BENCHLIB) $(HDRS)
	$(CCBIN) -O2 -std=c++17 -Wall -Iinclude -I$(CUDA_HOME)/include -o $@ $< -Lscrooge_b200/lib -lscrooge_b200 -lscrooge_b200_bench \
	    -L$(CUDA_HOME)/lib64 -lcudart -Wl,-rpath,'$$ORIGIN/../scrooge_b200/lib'

clean:
	rm -rf build scrooge_b200/lib

.PHONY: all clean
