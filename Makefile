# Builds libfrkb200.so (sm_100a only) in-tree, plus the oracle's C helper.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
# basis.cu carries the bit-exact distance/basis arithmetic: never contract mul+add there
SRC       := frk_b200/csrc
OBJ       := $(SRC)/build
LIB       := frk_b200/libfrkb200.so
CUS       := ctx comm basis bau sparse bucket dense model areal coupled host_api

all: $(LIB)

$(OBJ)/%.o: $(SRC)/%.cu $(SRC)/common.cuh $(SRC)/coupled.cuh include/frkb200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(if $(filter basis,$*),-fmad=false,) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(LIB): $(addprefix $(OBJ)/,$(addsuffix .o,$(CUS)))
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

clean:
	rm -rf $(OBJ) $(LIB)
.PHONY: all clean
