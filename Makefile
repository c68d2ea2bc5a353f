# Build libb2f_cuda.so (sm_100a only) and the CPU oracle library.  No cmake, no network.
NVCC      ?= nvcc
HOSTCC    ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v
CSRC      := back2future_b200/csrc
OBJDIR    := build
SRCS      := $(CSRC)/api.cu $(CSRC)/costvol.cu $(CSRC)/warp.cu $(CSRC)/criterions.cu $(CSRC)/conv.cu $(CSRC)/train.cu $(CSRC)/conv_tc.cu $(CSRC)/costvol_tc.cu $(CSRC)/wgrad_tc.cu
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := back2future_b200/libb2f_cuda.so
COMMLIB   := back2future_b200/libb2f_comm.so
ORACLE    := oracle/c/libb2f_cpu.so
CHECK64   := oracle/c/libb2f_check64.so
# the reference's own sampler (test-only parity pin), compiled from where it lies; only when the reference is mounted
REFROOT   ?= /root/reference
REFLIB    := oracle/_ref/libstn_ref.so

all: $(LIB) $(COMMLIB) $(ORACLE) $(CHECK64)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh $(CSRC)/tma.cuh include/b2f.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -Xlinker --version-script=$(CSRC)/exports.map

# the gradient all-reduce behind its own C ABI (include/b2f_comm.h); NCCL is dlopen'ed, not linked
$(COMMLIB): $(CSRC)/comm.cu include/b2f_comm.h
	$(NVCC) -O2 -std=c++17 $(ARCH) -Xcompiler -fPIC,-fvisibility=hidden -shared -o $@ $< -ldl

$(ORACLE): oracle/c/b2f_cpu.c
	$(HOSTCC) -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -fvisibility=hidden -o $@ $< -lm

# float64 closed-form checker used for the full-size parity gates (tests/test_bench_parity.py, bench.py `parity`)
$(CHECK64): oracle/c/b2f_check64.c
	$(HOSTCC) -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp -fPIC -shared -fvisibility=hidden -o $@ $< -lm

# Reference sampler: unmodified extras/stnbhwd/{utils.c,BilinearSamplerBHWD.cu} against the stand-in Torch7
# headers in oracle/ref_shim (the reference's own CMake build wants luarocks + TH/THC/luaT and -arch=sm_30).
$(REFLIB): oracle/ref_shim/stn_ref.cu $(wildcard oracle/ref_shim/*.h)
	@mkdir -p oracle/_ref
	$(NVCC) -O2 -std=c++17 $(ARCH) -Xcompiler -fPIC,-fvisibility=hidden -diag-suppress 177 -shared \
	    -I oracle/ref_shim -I $(REFROOT)/extras/stnbhwd -o $@ $<

ref: $(REFLIB)

clean:
	rm -rf $(OBJDIR) $(LIB) $(COMMLIB) $(ORACLE) $(CHECK64)

.PHONY: all clean ref
