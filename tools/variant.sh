#!/bin/bash
# tools/variant.sh NAME file.cu [-DFLAG ...]: builds dogs_b200/build/lib_NAME.so with one translation unit
# recompiled under extra flags (A/B runs on the GPU box: DOGS_B200_LIB=dogs_b200/build/lib_NAME.so).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p dogs_b200/build/variants
obj=dogs_b200/build/variants/${src%.cu}_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
  -c dogs_b200/csrc/$src -o $obj 2> dogs_b200/build/variants/${src%.cu}_$name.ptxas.log
objs=$(ls dogs_b200/build/*.o | grep -v "/${src%.cu}\.o$")
/usr/local/cuda/bin/nvcc -shared -o dogs_b200/build/lib_$name.so $objs $obj -lcudart
grep "Used\|spill" dogs_b200/build/variants/${src%.cu}_$name.ptxas.log | head -6
echo dogs_b200/build/lib_$name.so
