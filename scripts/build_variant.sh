#!/bin/bash
# Builds an experimental variant of the library for A/B runs on the GPU box (selected with MF_LIB=<path>):
#   scripts/build_variant.sh NAME "-DMACRO=1 ..."   ->  cuda-recommender_b200/libmfb200_NAME.so
set -e
cd "$(dirname "$0")/../cuda-recommender_b200"
NAME=$1; DEFS=$2
rm -rf build_var/$NAME libmfb200_$NAME.so
mkdir -p build_var/$NAME
pids=()
for f in csrc/*.cu; do
  o=build_var/$NAME/$(basename ${f%.cu}).o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ $DEFS -c $f -o $o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libmfb200_$NAME.so build_var/$NAME/*.o -lcudart -ldl -ccbin /usr/bin/g++
echo built libmfb200_$NAME.so
