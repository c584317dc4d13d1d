#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DFB_PF=2 ..."   -> moonbit_flate_b200/variants/libflate_b200_NAME.so
# (kernel experiments: same sources, other compile-time switches; select with FB200_LIB=<path>)
set -e
cd "$(dirname "$0")/../moonbit_flate_b200/csrc"
name=$1; shift
out=../variants; mkdir -p $out/obj_$name
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in api parse encode inflate inflate3; do
  $NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c -o $out/obj_$name/$f.o $f.cu &
done
wait
$NVCC $ARCH -shared -o $out/libflate_b200_$name.so $out/obj_$name/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $out/obj_$name
echo built $out/libflate_b200_$name.so
