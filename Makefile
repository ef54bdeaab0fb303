# Builds the C-ABI shared library (sm_100a only) and the thin torch extension over it; the oracle has no native part
# (torch CPU restatement).
NVCC ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden
SRC_DIR := nerf_qa_b200/csrc
SRCS := $(SRC_DIR)/ctx.cu $(SRC_DIR)/conv_tc.cu $(SRC_DIR)/conv_tc_halo.cu $(SRC_DIR)/conv_tc_pair.cu $(SRC_DIR)/conv_f32.cu $(SRC_DIR)/conv1_1_tc.cu $(SRC_DIR)/stage1_fused.cu $(SRC_DIR)/pool_stats.cu $(SRC_DIR)/pool_stats_tma.cu $(SRC_DIR)/backward.cu $(SRC_DIR)/adists.cu $(SRC_DIR)/microbench.cu $(SRC_DIR)/adists_tc.cu $(SRC_DIR)/adists_strips.cu $(SRC_DIR)/resize.cu
OBJS := $(SRCS:.cu=.o)
LIB := nerf_qa_b200/libnerfqa_b200.so

TORCH_EXT := nerf_qa_b200/_nqa_torch_ops.so
PY ?= python

all: $(LIB) $(TORCH_EXT)

# the thin PyTorch C++ extension over the C-ABI (torch.ops.nerfqa_b200.*): host code only, links the library above
$(TORCH_EXT): nerf_qa_b200/csrc_torch/nqa_torch_ops.cpp include/nerfqa_b200.h $(LIB)
	$(PY) scripts/build_torch_ext.py $@

$(SRC_DIR)/%.o: $(SRC_DIR)/%.cu $(SRC_DIR)/common.cuh $(SRC_DIR)/tc_common.cuh $(SRC_DIR)/kernels.h $(SRC_DIR)/adists.h $(SRC_DIR)/adists_tc.h include/nerfqa_b200.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $@ $(OBJS)

clean:
	rm -f $(OBJS) $(LIB) $(TORCH_EXT)

.PHONY: all clean
