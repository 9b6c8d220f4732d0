# Build of libaleo_b200.so (product, sm_100a only) and of the development emulator.
#   make -j        product library  -> aleo_b200/libaleo_b200.so   (in-tree: travels to the GPU box)
#   make -j emu    kernel-logic emulator (g++, no GPU)  -> build/libaleo_b200_emu.so
#   make oracle    C oracle / CPU baseline -> oracle/_build/liboracle.so
NVCC      ?= nvcc
CXX       ?= g++
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -Xptxas -v
CSRC      := aleo_b200/csrc
TUS       := capi ntt_lib msm_lib util_lib poly_lib
HDRS      := $(wildcard $(CSRC)/*.cuh) $(CSRC)/internal.h include/aleo_b200.h
OBJ       := $(patsubst %,build/%.o,$(TUS))
EMUOBJ    := $(patsubst %,build/emu_%.o,$(TUS))

all: aleo_b200/libaleo_b200.so

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ > build/$*.ptxas.log 2>&1 || (cat build/$*.ptxas.log; false)

aleo_b200/libaleo_b200.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

build/emu_%.o: $(CSRC)/%.cu $(HDRS) tests/emu/emu_runtime.hpp
	@mkdir -p build
	$(CXX) -std=c++20 -O2 -DALEO_EMU -x c++ -fPIC -pthread -Wno-psabi -c $< -o $@

emu: build/libaleo_b200_emu.so
build/libaleo_b200_emu.so: $(EMUOBJ)
	$(CXX) -shared -pthread -o $@ $(EMUOBJ)

oracle: oracle/_build/liboracle.so
oracle/_build/liboracle.so: oracle/oracle.c
	@mkdir -p oracle/_build
	$(CC) -O3 -march=x86-64-v3 -fPIC -shared -pthread -o $@ $<

clean:
	rm -rf build aleo_b200/libaleo_b200.so oracle/_build

.PHONY: all emu oracle clean
