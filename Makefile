# Builds the C-ABI CUDA library (sm_100a only) and nothing else.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
CSRC := torch_semantic_segmentation_b200/csrc
SRCS := $(wildcard $(CSRC)/*.cu)
OBJS := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
LIB := torch_semantic_segmentation_b200/libtss_b200.so

all: $(LIB)

HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/tss_b200.h

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart static

# Instrumented copy for tools/trace_kernels.py (TSS_MARK timestamps inside the kernels); never loaded by the product.
TRACE_LIB := torch_semantic_segmentation_b200/libtss_b200_trace.so
TRACE_OBJS := $(patsubst $(CSRC)/%.cu,build_trace/%.o,$(SRCS))
build_trace/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build_trace
	$(NVCC) $(NVCCFLAGS) -DTSS_TRACE -c $< -o $@
$(TRACE_LIB): $(TRACE_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(TRACE_OBJS) -cudart static
trace: $(TRACE_LIB)

clean:
	rm -rf build build_trace $(LIB) $(TRACE_LIB)
.PHONY: all clean trace
