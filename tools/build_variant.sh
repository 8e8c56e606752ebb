#!/bin/sh
# Development tool: build libraytrace_b200 with extra nvcc flags into
# build/variants/lib<name>.so (objects in build/var_<name>/) for A/B runs on the
# GPU box (tools/lbvh_ab.py --lib, tools/ab_kernels.py --lib).
#   tools/build_variant.sh i16r1 "-DRT_WALK_ITERS=16 -DRT_WALK_ROUNDS=1"
set -e
name="$1"; shift
cd "$(dirname "$0")/.."
mkdir -p build/variants
make -s OBJ="build/var_$name" LIB="build/variants/lib$name.so" NVFLAGS_EXTRA="$*" "build/variants/lib$name.so"
echo "build/variants/lib$name.so"
