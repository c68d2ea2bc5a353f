# Build libb2f_cuda.so (sm_100a only) and the CPU oracle library.  No cmake, no network.
NVCC      ?= nvcc
HOSTCC    ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CSRC      := back2future_b200/csrc
OBJDIR    := build
SRCS      := $(CSRC)/api.cu $(CSRC)/costvol.cu $(CSRC)/warp.cu $(CSRC)/criterions.cu
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := back2future_b200/libb2f_cuda.so
ORACLE    := oracle/c/libb2f_cpu.so

all: $(LIB) $(ORACLE)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/tma.cuh include/b2f.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker --version-script=$(CSRC)/exports.map

$(ORACLE): oracle/c/b2f_cpu.c
	$(HOSTCC) -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -fvisibility=hidden -o $@ $< -lm

clean:
	rm -rf $(OBJDIR) $(LIB) $(ORACLE)

.PHONY: all clean
