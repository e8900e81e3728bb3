#!/bin/bash
# Build a kernel variant of libdsmgp.so for A/B runs: tools/variant.sh <name> "<-D flags>"  ->  build_variants/libdsmgp_<name>.so
# (only the engine kernels' translation unit k_v2.cu is recompiled with the flags; everything else is the current build)
set -e
NAME=$1; FLAGS=$2
SRC=deepstructuredmixtures_b200/csrc
OUT=build_variants; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
make -C $SRC -j8 > /dev/null
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 177 $FLAGS -c $SRC/k_v2.cu -o $OUT/k_v2_$NAME.o
OBJS=$(ls $SRC/*.o | grep -v k_v2.o)
nvcc $ARCH -shared -o $OUT/libdsmgp_$NAME.so $OBJS $OUT/k_v2_$NAME.o -lcudart -ldl
echo built $OUT/libdsmgp_$NAME.so
