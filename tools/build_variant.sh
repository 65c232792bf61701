#!/bin/bash
# Builds a variant of the library with extra -D flags for tuning experiments:
#   tools/build_variant.sh NAME file.cu [file2.cu ...] -- -DFOO=1 ...
# Only the named sources are recompiled; the result is cloudatlas.jl_b200/lib/variant_NAME.so, used through
# CLOUDATLAS_B200_LIB=cloudatlas.jl_b200/lib/variant_NAME.so.
set -e
cd "$(dirname "$0")/../cloudatlas.jl_b200/csrc"
name=$1; shift
srcs=()
while [ "$1" != "--" ]; do srcs+=("$1"); shift; done
shift
mkdir -p ../build/variant_$name
objs=()
for o in ../build/*.o; do
  keep=1
  for s in "${srcs[@]}"; do [ "$(basename $o)" == "$s.o" ] && keep=0; done
  [ $keep == 1 ] && objs+=("$o")
done
for s in "${srcs[@]}"; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $s -o ../build/variant_$name/$s.o
  objs+=("../build/variant_$name/$s.o")
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variant_$name.so "${objs[@]}" -cudart static -lpthread
echo built ../lib/variant_$name.so
