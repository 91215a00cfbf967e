#!/bin/bash
# developer tool: build libvbmp_b200_<name>.so with extra -D flags on one source (variant timing experiments)
# usage: tools/build_variant.sh <name> <source.cu> [-DFLAG ...]
set -e
cd "$(dirname "$0")/../pyvbmp_b200/csrc"
name=$1; src=$2; shift 2
make -s > /dev/null
mkdir -p build/var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $src -o build/var/${src%.cu}_$name.o
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o"; echo build/var/${src%.cu}_$name.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libvbmp_b200_$name.so $objs
echo built libvbmp_b200_$name.so
