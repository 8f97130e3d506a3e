# Builds the product library (CUDA, sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  := /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -std=c++17 -O3 -lineinfo --extended-lambda $(ARCH) -ccbin $(CXX) -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-free-nonheap-object -Iinclude
CSRC := bwtb3m_b200/csrc
OBJS := $(CSRC)/engine.o $(CSRC)/sufsort.o $(CSRC)/stages.o $(CSRC)/blocks.o $(CSRC)/rlencode.o $(CSRC)/formats.o $(CSRC)/hostapi.o $(CSRC)/multi.o
LIB  := bwtb3m_b200/libb3m.so
BINS := $(patsubst cli/%.cpp,bin/%,$(wildcard cli/*.cpp))

all: $(LIB) $(BINS) oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/b3m.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(CSRC)/%.o: $(CSRC)/%.cpp $(wildcard $(CSRC)/*.h) include/b3m.h
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -ccbin $(CXX) -shared -o $@ $(OBJS)

bin/%: cli/%.cpp $(LIB) include/b3m.h $(wildcard cli/*.h)
	@mkdir -p bin
	$(CXX) -std=c++17 -O2 -Wall -Iinclude -o $@ $< -Lbwtb3m_b200 -lb3m -lz -Wl,-rpath,'$$ORIGIN/../bwtb3m_b200'

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(CSRC)/*.o $(LIB) $(BINS)
	$(MAKE) -s -C oracle clean
.PHONY: all oracle clean
