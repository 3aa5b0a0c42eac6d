# Build of the B200-native CFD-Proxy hot path (sm_100a only) and of the oracle rig.
#   make lib      -> cfd_proxy_b200/libcfdp_b200.so   (product: C ABI + CUDA kernels)
#   make oracle   -> oracle/liboracle.so              (CPU restatement, test infrastructure)
#   make ref      -> oracle/_ref/*                    (unmodified reference + shims; needs /root/reference)
NVCC    ?= /usr/local/cuda/bin/nvcc
CC      := gcc
CXX     := g++
ARCH    := -gencode arch=compute_100a,code=sm_100a
PKG     := cfd_proxy_b200
SRC     := $(PKG)/csrc
BUILD   := build
INC     := -Iinclude -I$(SRC)
CFLAGS  := -O3 -fPIC -fopenmp -std=gnu11 -Wall -Wno-unused-function $(INC)
CXXFLAGS:= -O3 -fPIC -fopenmp -std=c++17 -Wall -Wno-unused-function $(INC) -I/usr/local/cuda/include
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unused-function $(INC) -Xptxas -v

C_SRCS   := $(wildcard $(SRC)/*.c)
CXX_SRCS := $(wildcard $(SRC)/*.cpp)
CU_SRCS  := $(wildcard $(SRC)/*.cu)
OBJS := $(patsubst $(SRC)/%.c,$(BUILD)/%.o,$(C_SRCS)) \
        $(patsubst $(SRC)/%.cpp,$(BUILD)/%.o,$(CXX_SRCS)) \
        $(patsubst $(SRC)/%.cu,$(BUILD)/%.cu.o,$(CU_SRCS))
HDRS := $(wildcard include/*.h) $(wildcard $(SRC)/*.h) $(wildcard $(SRC)/*.cuh)

.PHONY: all lib oracle ref clean
all: lib oracle

lib: $(PKG)/libcfdp_b200.so

$(BUILD)/%.o: $(SRC)/%.c $(HDRS) | $(BUILD)
	$(CC) $(CFLAGS) -c $< -o $@
$(BUILD)/%.o: $(SRC)/%.cpp $(HDRS) | $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(BUILD)/%.cu.o: $(SRC)/%.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)
$(BUILD):
	mkdir -p $(BUILD)

$(PKG)/libcfdp_b200.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -Xcompiler -fopenmp -lgomp -ldl -lpthread

oracle:
	$(MAKE) -C oracle liboracle.so
ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(BUILD) $(PKG)/libcfdp_b200.so oracle/liboracle.so oracle/_ref/*
