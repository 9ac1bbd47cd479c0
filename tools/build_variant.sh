#!/bin/bash
# Build krets_b200/variants/libkrets_b200_<name>.so: the in-tree objects with runtime.cu recompiled under extra -D flags
# (kernel experiments are compared on the GPU in one gpurun call: KRETS_B200_LIB=krets_b200/variants/lib..._<name>.so).
# usage: tools/build_variant.sh <name> [-DFLAG=...]...
set -e
name=$1; shift
cd "$(dirname "$0")/../krets_b200/csrc"
mkdir -p ../variants build/var_$name
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin g++ -Xcompiler -fPIC "$@" -c runtime.cu -o build/var_$name/runtime.o
objs=$(ls build/*.o build/host/*.o | grep -v "build/runtime.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libkrets_b200_$name.so $objs build/var_$name/runtime.o -cudart static -ldl
echo built ../variants/libkrets_b200_$name.so
