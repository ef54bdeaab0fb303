# Builds the C-ABI shared library (sm_100a only) and the oracle has no native part (torch CPU restatement).
NVCC ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden
SRC_DIR := nerf_qa_b200/csrc
SRCS := $(SRC_DIR)/ctx.cu $(SRC_DIR)/conv_tc.cu $(SRC_DIR)/conv_tc_halo.cu $(SRC_DIR)/conv_tc_pair.cu $(SRC_DIR)/conv_f32.cu $(SRC_DIR)/conv1_1_tc.cu $(SRC_DIR)/stage1_fused.cu $(SRC_DIR)/pool_stats.cu $(SRC_DIR)/pool_stats_tma.cu $(SRC_DIR)/backward.cu $(SRC_DIR)/adists.cu $(SRC_DIR)/microbench.cu $(SRC_DIR)/adists_tc.cu $(SRC_DIR)/resize.cu
OBJS := $(SRCS:.cu=.o)
LIB := nerf_qa_b200/libnerfqa_b200.so

all: $(LIB)

$(SRC_DIR)/%.o: $(SRC_DIR)/%.cu $(SRC_DIR)/common.cuh $(SRC_DIR)/tc_common.cuh $(SRC_DIR)/kernels.h $(SRC_DIR)/adists.h $(SRC_DIR)/adists_tc.h include/nerfqa_b200.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $@ $(OBJS)

clean:
	rm -f $(OBJS) $(LIB)

.PHONY: all clean
