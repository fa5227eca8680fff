# Builds the C-ABI shared library (sm_100a only) and the oracle's C helper.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -cudart static
SRC := gddim_b200/csrc
OBJ := build/obj
LIB := gddim_b200/libgddim_b200.so
CU := $(SRC)/conv_gemm.cu $(SRC)/attn.cu $(SRC)/gn_qkv.cu $(SRC)/norm.cu $(SRC)/small.cu $(SRC)/update.cu
# `make ABLATE=1` compiles the timing-only epilogue ablations (GDDIM_GEMM_DBG, results invalid) and the clock64 timeline;
# never in the default library.
ifeq ($(ABLATE),1)
NVFLAGS += -DGDDIM_ABLATE
endif
CPP := $(SRC)/tables.cpp $(SRC)/unet.cpp $(SRC)/api.cpp
OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU)) $(patsubst $(SRC)/%.cpp,$(OBJ)/%.o,$(CPP))
HDRS := $(wildcard $(SRC)/*.h $(SRC)/*.cuh include/*.h)

all: $(LIB) oracle

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/%.o: $(SRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS)

oracle:
	@mkdir -p oracle/_build
	gcc -O2 -shared -fPIC -o oracle/_build/libcld_ode.so oracle/cld_ode.c -lm

clean:
	rm -rf build $(LIB) oracle/_build

.PHONY: all oracle clean
