#!/bin/bash
# Build a variant of the library with extra nvcc flags (kernel tuning experiments):
#   bash profiles/build_variant.sh w24 -DCLEDB_PRUNED_WPC=24   ->  solar-coronal-inversion_b200/lib/variants/libcledb_b200_w24.so
# Use it with CLEDB_B200_LIB=<path> (cledb_b200/native.py).
set -e
tag=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
src=$here/solar-coronal-inversion_b200/csrc
out=$here/solar-coronal-inversion_b200/lib/variants
mkdir -p $out/obj_$tag
for f in api db match elementwise solution reduced integrate comm; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c $src/$f.cu -o $out/obj_$tag/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libcledb_b200_$tag.so $out/obj_$tag/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $out/obj_$tag
echo built $out/libcledb_b200_$tag.so
