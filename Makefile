# Build of the drop-in library without Python (the same nvcc line as livevideo_b200/_build.py):
#   make            -> livevideo_b200/liblivevideo.so   (sm_100a; nvcc cross-compiles without a GPU)
#   make harness    -> tools/harness/shim_demo, lv_harness (C++ consumers of include/livevideo.h / lv_convert_shim.hpp)
#   make oracle     -> oracle/liblv_oracle.so            (test infrastructure: the CPU checker, never linked by the product)
NVCC  ?= $(shell command -v nvcc 2>/dev/null || echo /usr/local/cuda/bin/nvcc)
CSRC  := livevideo_b200/csrc
SRCS  := $(addprefix $(CSRC)/,lv_kernels.cu lv_zerocopy.cu lv_tma.cu lv_aux.cu lv_formats.cu lv_shard.cu lv_exchange.cu lv_capi.cu lv_synth.cu)
DEPS  := $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.h) include/livevideo.h
LIB   := livevideo_b200/liblivevideo.so

all: $(LIB)

$(LIB): $(DEPS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DLV_BUILDING_LIBRARY -shared -Xcompiler -fPIC,-fvisibility=hidden \
	    -Xptxas -O3 -cudart shared -ldl -o $@ $(SRCS)

oracle:
	gcc -O3 -fPIC -std=c11 -pthread -shared -o oracle/liblv_oracle.so oracle/lv_oracle.c -lm

harness: $(LIB) oracle
	$(MAKE) -C tools/harness

clean:
	rm -f $(LIB)
	$(MAKE) -C tools/harness clean

.PHONY: all oracle harness clean
