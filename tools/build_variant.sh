#!/bin/bash
# usage: tools/build_variant.sh NAME -DFOO=1 ... : variants/libmodgpu_NAME.so = the library with the select kernels
# (hash_select.cu, hash_count2.cu) rebuilt with extra flags; select it with MODGPU_LIB=variants/libmodgpu_NAME.so
set -e
cd "$(dirname "$0")/../modimizer_b200/csrc"
N=$1; shift
mkdir -p ../../variants/$N
for f in hash_select hash_count2; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c $f.cu -o ../../variants/$N/$f.o 2> ../../variants/$N/$f.log &
done
wait
OBJ=$(ls build/*.o | grep -v "hash_select.o\|hash_count2.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libmodgpu_$N.so $OBJ ../../variants/$N/hash_select.o ../../variants/$N/hash_count2.o -lz -ldl
grep -h -A2 "hash_count2_kernelILi1ELi31ELi1ELb0ELb0\|hash_count2_kernelILi0ELi1ELi1ELb0ELb0" ../../variants/$N/hash_count2.log | grep "Used"
