# Build libamgb200.so (the product: sm_100a CUDA kernels + C ABI + host helpers),
# tests/libamgb200_testhooks.so (CPU emulations of the schedule / layouts for the CPU test-suite; not shipped in the product) and
# oracle/liboracle.so (the CPU checker; test infrastructure only).
NVCC      ?= nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the reference's CPU objects contain no FMA (gcc, baseline x86-64); kernels that
# promise bit-identical rows additionally use __dmul_rn/__dadd_rn explicitly.
NVFLAGS   := -O3 -std=c++17 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fopenmp -Iinclude
CSRC      := amg_b200/csrc
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(filter-out $(CSRC)/debug_host.cpp,$(wildcard $(CSRC)/*.cpp))
OBJS      := $(CU_SRCS:.cu=.o) $(CPP_SRCS:.cpp=.o)
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh) include/amg_b200.h

all: amg_b200/libamgb200.so tests/libamgb200_testhooks.so oracle/liboracle.so oracle/libmatgen.so

amg_b200/libamgb200.so: $(OBJS)
	$(NVCC) -shared $(ARCH) -Xcompiler -fopenmp -o $@ $(OBJS) -lcudart

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(@:.o=.ptxas.log) || (cat $(@:.o=.ptxas.log); false)

$(CSRC)/%.o: $(CSRC)/%.cpp $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

tests/libamgb200_testhooks.so: $(CSRC)/debug_host.o $(CSRC)/analysis.o
	$(NVCC) -shared -Xcompiler -fopenmp -o $@ $^

oracle/liboracle.so: oracle/amg_oracle.c include/amg_b200.h
	$(CC) -O2 -ffp-contract=off -fPIC -shared -o $@ oracle/amg_oracle.c -lm

oracle/libmatgen.so: oracle/matgen.c
	$(CC) -O2 -ffp-contract=off -fPIC -shared -o $@ oracle/matgen.c -lm

clean:
	rm -f $(CSRC)/*.o $(CSRC)/*.ptxas.log amg_b200/libamgb200.so tests/libamgb200_testhooks.so oracle/liboracle.so oracle/libmatgen.so

.PHONY: all clean
