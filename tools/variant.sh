#!/bin/bash
# Build a kernel variant of libdsmgp.so for A/B runs: tools/variant.sh <name> "<-D flags>"  ->  build_variants/libdsmgp_<name>.so
set -e
NAME=$1; FLAGS=$2
SRC=deepstructuredmixtures_b200/csrc
OUT=build_variants; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 177 $FLAGS -c $SRC/k_v2.cu -o $OUT/k_v2_$NAME.o
nvcc $ARCH -shared -o $OUT/libdsmgp_$NAME.so $SRC/api.o $SRC/api_predict.o $SRC/api_ops.o $SRC/comm.o $SRC/k_gram.o $SRC/k_potrf.o $SRC/k_lauum.o $SRC/k_misc.o $OUT/k_v2_$NAME.o -lcudart -ldl
echo built $OUT/libdsmgp_$NAME.so
