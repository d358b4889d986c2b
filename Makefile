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

.PHONY: all lib oracle ref dropin clean
all: lib oracle

lib: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(FLAGS_$*) -c $< -o $@

# float code written as plain expressions in the reference's evaluation order: no FMA contraction (the reference build has none)
FLAGS_esbr_hbe_kernel := -fmad=false
FLAGS_esbr_ps_kernel  := -fmad=false

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ)

oracle:
	$(MAKE) -s -C oracle oracle

ref:
	$(MAKE) -s -C oracle ref

# The drop-in demonstration: the reference's OWN testbench + decoder library (unmodified, compiled by oracle/Makefile) linked
# with the stage overrides of libxaac_b200/dropin/ixheaacd_b200_glue.c (ld --wrap) against libxaac_b200.so.  The binary decodes
# real files through the reference's parser with the DSP stages running on the GPU (tests/test_dropin_gpu.py).
REF        ?= /root/reference
DROPIN     := libxaac_b200/dropin
DROPIN_OUT := $(DROPIN)/_build
DROPIN_WRAPS := -Wl,--wrap=ixheaacd_imdct_process -Wl,--wrap=ixheaacd_sbr_dec -Wl,--wrap=ixheaacd_fd_frm_dec \
                -Wl,--wrap=ixheaacd_channel_pair_process -Wl,--wrap=ixheaacd_dec_sbrdata -Wl,--wrap=ixheaacd_decode_ps_data
DROPIN_FLAGS := -std=gnu99 -D_X86_ -DX86_64 -D_X86_64_ -DLOUDNESS_LEVELING_SUPPORT -O2 -fwrapv -w \
                -UARM_PROFILE_HW -UARM_PROFILE_BOARD -DDRC_ENABLE -DMULTICHANNEL_ENABLE -DECLIPSE -DWIN32
dropin: $(DROPIN_OUT)/xaacdec_b200
$(DROPIN_OUT)/xaacdec_b200: $(LIB) $(DROPIN)/ixheaacd_b200_glue.c $(DROPIN)/ixheaacd_b200_pack.h $(DROPIN)/ixheaacd_b200_pack_ps_flt.h $(DROPIN)/ixheaacd_b200_pack_spec.h $(DROPIN)/ixheaacd_b200_pack_sd.h $(DROPIN)/ixheaacd_b200_ref_headers.h include/xaac_b200.h oracle/_ref/libxaacdec.a
	@mkdir -p $(DROPIN_OUT)
	gcc $(DROPIN_FLAGS) -I$(REF)/common -I$(REF)/decoder -I$(REF)/decoder/drc_src -I$(REF)/test/decoder -I$(DROPIN) -Iinclude \
	    -o $@ $(wildcard $(REF)/test/decoder/*.c) $(DROPIN)/ixheaacd_b200_glue.c $(DROPIN_WRAPS) oracle/_ref/libxaacdec.a \
	    -Llibxaac_b200 -lxaac_b200 -Wl,-rpath,'$$ORIGIN/../..' -lm
oracle/_ref/libxaacdec.a:
	$(MAKE) -s -C oracle ref

clean:
	rm -rf build $(LIB) $(DROPIN_OUT)
	$(MAKE) -s -C oracle clean
