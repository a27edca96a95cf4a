# Builds libsuperman_b200.so (C host + sm_100a CUDA kernels) and the `perman` CLI, in-tree.
#   make -j8            everything
#   make oracle         the CPU oracle (test infrastructure) and, when /root/reference exists,
#                       the reference shim under oracle/_ref/
NVCC      ?= nvcc
CC        := gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Isuperman_b200/csrc
CFLAGS    := -O2 -std=c11 -fPIC -Wall -Wextra -Iinclude -Isuperman_b200/host -pthread
CUDA_HOME ?= /usr/local/cuda

BUILD   := build
PKG     := superman_b200
LIB     := $(PKG)/libsuperman_b200.so
CLI     := $(PKG)/perman
CONNECT := $(PKG)/libConnect.so

GROUPS  := 0 1 2 3 4 5 6 7
CU_SRCS := sp_device sp_dense sp_sparse sp_approx
CU_OBJS := $(CU_SRCS:%=$(BUILD)/%.o) $(BUILD)/sp_level_inst_b3s0.o $(BUILD)/sp_level_inst_b3s1.o $(BUILD)/sp_level_inst_b4s0.o $(BUILD)/sp_level_inst_b4s1.o $(GROUPS:%=$(BUILD)/sp_dense_inst_g%.o) $(GROUPS:%=$(BUILD)/sp_sparse_inst_g%.o)
C_SRCS  := sp_sched sp_api sp_matrix sp_reduce sp_connector sp_level
C_OBJS  := $(C_SRCS:%=$(BUILD)/%.o)

all: $(LIB) $(CLI) $(CONNECT) profiles/resource_usage.txt

# registers / stack (spills) / shared memory of every shipped kernel, regenerated from the linked library
profiles/resource_usage.txt: $(LIB) tools/resource_report.py
	python3 tools/resource_report.py $(LIB) $@

$(BUILD):
	mkdir -p $(BUILD)

$(BUILD)/%.o: $(PKG)/csrc/%.cu $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# the estimators are compared bit for bit with the C oracle: no FMA contraction in this unit
$(BUILD)/sp_approx.o: $(PKG)/csrc/sp_approx.cu $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h) | $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

$(BUILD)/sp_level_inst_b3s0.o $(BUILD)/sp_level_inst_b3s1.o $(BUILD)/sp_level_inst_b4s0.o $(BUILD)/sp_level_inst_b4s1.o: \
$(BUILD)/sp_level_inst_b%.o: $(PKG)/csrc/sp_level_inst.cu $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DSPB_LV_B=$(word 1,$(subst s, ,$*)) -DSPB_LV_SKIP=$(word 2,$(subst s, ,$*)) -c $< -o $@

$(BUILD)/sp_dense_inst_g%.o: $(PKG)/csrc/sp_dense_inst.cu $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DSPB_GROUP=$* -c $< -o $@

$(BUILD)/sp_sparse_inst_g%.o: $(PKG)/csrc/sp_sparse_inst.cu $(wildcard $(PKG)/csrc/*.cuh $(PKG)/csrc/*.h include/*.h) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DSPB_GROUP=$* -c $< -o $@

$(BUILD)/%.o: $(PKG)/host/%.c $(wildcard $(PKG)/host/*.h include/*.h) | $(BUILD)
	$(CC) $(CFLAGS) -c $< -o $@

$(LIB): $(CU_OBJS) $(C_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lpthread -lm

$(CLI): $(PKG)/host/perman_main.c $(LIB) include/superman_b200.h
	$(CC) -O2 -std=c11 -Wall -Wextra -Iinclude -o $@ $< -L$(PKG) -lsuperman_b200 -lm -Wl,-rpath,'$$ORIGIN'

# the file the reference's bindings load by name (superPython.py, supermaTlab.m): reference symbol names,
# forwarding to $(LIB)
$(CONNECT): $(PKG)/host/libconnect.c $(LIB) include/superman_b200.h
	$(CC) -O2 -std=c11 -fPIC -shared -Wall -Wextra -Iinclude -Wl,-Bsymbolic -o $@ $< -L$(PKG) -lsuperman_b200 -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(BUILD) $(LIB) $(CLI) $(CONNECT)

.PHONY: all oracle clean
