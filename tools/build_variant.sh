#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...]   (env SRC=dir with an alternative rollout_tc.cu / headers; default csrc)
# Builds build/variants/NAME.so: rollout_tc.cu recompiled with the extra flags, other objects from the regular build.
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SRC=${SRC:-tdmpc_b200/csrc}
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v \
  -I tdmpc_b200/csrc -I include "$@" -c $SRC/rollout_tc.cu -o build/variants/$NAME.o 2> build/variants/$NAME.ptxas.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$NAME.so build/obj/api.o build/obj/encode.o build/obj/rollout_fp32.o \
  build/variants/$NAME.o build/obj/select.o build/obj/replay.o build/obj/update_fwd.o build/obj/train.o -lcudart_static -lpthread -ldl -lrt
grep -h "registers\|spill" build/variants/$NAME.ptxas.log | head -2
