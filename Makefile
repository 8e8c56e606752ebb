# Top-level build: libraytrace_b200.so (C ABI, include/rt_cuda.h) for sm_100a.
#
#   make            build the product library in-tree (ray_tracing_b200/)
#   make oracle     build the test-only checkers (oracle/)
#   make harness    build the headless C harness (tools/rt_headless)
# `make` also builds tools/librt_skybox_stb.so when the reference's 3p/ is present.
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
SRC       := ray_tracing_b200/csrc
OBJ       := build/obj
LIB       := ray_tracing_b200/libraytrace_b200.so
HARNESS   := tools/rt_headless
SKYLIB    := tools/librt_skybox_stb.so
STB_DIR   ?= /root/reference/3p
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Iinclude -I$(SRC) -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
             --expt-relaxed-constexpr -Xptxas -v $(NVFLAGS_EXTRA)
# host C: the reference's own flags matter for float parity (no contraction)
CFLAGS    := -std=c11 -O2 -fPIC -ffp-contract=off -Wall -Wextra -Iinclude -I$(SRC)

HOST_OBJS := $(OBJ)/scene_parse.o $(OBJ)/camera_host.o $(OBJ)/scene_pack.o $(OBJ)/screenshot.o $(OBJ)/bvh_sah.o
CUDA_OBJS := $(OBJ)/rt_api.o $(OBJ)/rt_lbvh.o $(OBJ)/rt_render_exact.o $(OBJ)/rt_render_fast.o
DEVICE_HDRS := $(SRC)/rt_device.cuh $(SRC)/rt_params.h $(SRC)/rt_host.h $(SRC)/rt_lbvh.h $(SRC)/rt_lbvh_rule.h include/rt_cuda.h

.PHONY: all oracle harness clean
all: $(LIB) $(HARNESS) $(if $(wildcard $(STB_DIR)/stb/stb_image.h),$(SKYLIB),)

$(OBJ)/%.o: $(SRC)/%.c include/rt_cuda.h $(SRC)/rt_host.h
	@mkdir -p $(OBJ)
	$(CC) $(CFLAGS) -c -o $@ $<

$(OBJ)/rt_api.o: $(SRC)/rt_api.cu $(DEVICE_HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJ)/rt_api.ptxas.log || (cat $(OBJ)/rt_api.ptxas.log; false)
$(OBJ)/rt_lbvh.o: $(SRC)/rt_lbvh.cu $(DEVICE_HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(OBJ)/rt_lbvh.ptxas.log || (cat $(OBJ)/rt_lbvh.ptxas.log; false)
# the render kernels twice: bit-exact (no FMA contraction) and fast
$(OBJ)/rt_render_exact.o: $(SRC)/rt_render.cu $(DEVICE_HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DRT_NS=rt_exact -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -c -o $@ $< \
	    2> $(OBJ)/rt_render_exact.ptxas.log || (cat $(OBJ)/rt_render_exact.ptxas.log; false)
$(OBJ)/rt_render_fast.o: $(SRC)/rt_render.cu $(DEVICE_HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DRT_NS=rt_fast -DRT_FAST_MATH=1 -fmad=true -c -o $@ $< \
	    2> $(OBJ)/rt_render_fast.ptxas.log || (cat $(OBJ)/rt_render_fast.ptxas.log; false)

$(LIB): $(HOST_OBJS) $(CUDA_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart_static -lm -lpthread -ldl -lrt

oracle:
	$(MAKE) -C oracle all

# headless C stand-in for the reference's main() on top of the C ABI.  stb_image
# is used in place from the reference's vendored 3p/ when present (never copied).
harness: $(HARNESS)
$(HARNESS): tools/rt_headless.c include/rt_cuda.h $(LIB)
	$(CC) -std=c11 -O2 -Iinclude $(if $(wildcard $(STB_DIR)/stb/stb_image.h),-DRT_HAVE_STB -I$(STB_DIR),) \
	    -o $@ tools/rt_headless.c -Lray_tracing_b200 -lraytrace_b200 -Wl,-rpath,'$$ORIGIN/../ray_tracing_b200' -lm

# the reference's JPEG decoder (stb_image, compiled from the reference's 3p/ in place)
# for Python hosts: bench.py and the tests feed the texels the reference would
$(SKYLIB): tools/rt_skybox_stb.c
	$(CC) -std=c11 -O2 -fPIC -shared -w -I$(STB_DIR) -o $@ $< -lm

clean:
	rm -rf build $(LIB)
