# Builds libnerfb200.so (sm_100a only) in-tree.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
PROFILE   ?= 0
NVCCFLAGS := -DNB2_TC_PROFILE=$(PROFILE) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Iinclude
CSRC      := nerf_b200/csrc
SRCS      := $(CSRC)/nb2_api.cu $(CSRC)/nb2_ops.cu $(CSRC)/nb2_pack.cu $(CSRC)/nb2_mlp_simt.cu $(CSRC)/nb2_mlp_tc.cu $(CSRC)/nb2_mlp_tc4.cu $(CSRC)/nb2_gemm.cu $(CSRC)/nb2_train.cu $(CSRC)/nb2_refnerf.cu $(CSRC)/nb2_microbench.cu
OBJS      := $(SRCS:.cu=.o)
LIB       := nerf_b200/libnerfb200.so

all: $(LIB)

# the HBM-bound stages mirror PyTorch's unfused elementwise arithmetic: no FMA contraction there
$(CSRC)/nb2_ops.o: EXTRA += -fmad=false

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/nb2_common.cuh $(CSRC)/nb2_rowio.cuh $(CSRC)/nb2_tc_ptx.cuh $(CSRC)/nb2_tc_device.cuh include/nerf_b200.h
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

# same library with the tensor kernel's per-role cycle counters compiled in (tools/gpu_probe.py roles; NB2_LIB=libnerfb200_prof.so)
PROF_OBJS := $(SRCS:$(CSRC)/%.cu=build/prof/%.o)
build/prof/nb2_ops.o: EXTRA += -fmad=false
build/prof/%.o: $(CSRC)/%.cu $(CSRC)/nb2_common.cuh $(CSRC)/nb2_rowio.cuh $(CSRC)/nb2_tc_ptx.cuh $(CSRC)/nb2_tc_device.cuh include/nerf_b200.h
	@mkdir -p build/prof
	$(NVCC) $(NVCCFLAGS) -DNB2_TC_PROFILE=1 $(EXTRA) -c $< -o $@
prof: nerf_b200/libnerfb200_prof.so
nerf_b200/libnerfb200_prof.so: $(PROF_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(PROF_OBJS) -lcudart

clean:
	rm -rf $(OBJS) $(LIB) build/prof nerf_b200/libnerfb200_prof.so
.PHONY: all clean prof
