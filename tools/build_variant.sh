#!/bin/bash
# Build an experimental variant of the library next to the shipped one: extra nvcc flags -> vaevar_b200/libvaevar_<tag>.so
#     tools/build_variant.sh gw32 -DVV_EPI_GW=32        then:  VV_LIB=vaevar_b200/libvaevar_gw32.so python tools/variant_bench.py
set -eu
tag=$1; shift
cd "$(dirname "$0")/../vaevar_b200"
mkdir -p build_$tag
for f in gemm kernels obs_lbfgs engine lbfgs seams net1_kernels net1 mlp_fused; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c csrc/$f.cu -o build_$tag/$f.o &
done
wait
nvcc -shared -o libvaevar_$tag.so build_$tag/*.o -gencode arch=compute_100a,code=sm_100a
ls -la libvaevar_$tag.so
