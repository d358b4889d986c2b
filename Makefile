# Top-level build: product library (sm_100a only), CPU oracle, and — when /root/reference is present —
# the compiled reference used as the parity checker / CPU baseline.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -warn-spills
CSRC      := libxaac_b200/csrc
LIB       := libxaac_b200/libxaac_b200.so
CU        := $(wildcard $(CSRC)/*.cu)
HDR       := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/xaac_b200.h
OBJ       := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU))

.PHONY: all lib oracle ref clean
all: lib oracle

lib: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(FLAGS_$*) -c $< -o $@

# float code written as plain expressions in the reference's evaluation order: no FMA contraction (the reference build has none)
FLAGS_esbr_hbe_kernel := -fmad=false

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

oracle:
	$(MAKE) -s -C oracle oracle

ref:
	$(MAKE) -s -C oracle ref

clean:
	rm -rf build $(LIB)
	$(MAKE) -s -C oracle clean
