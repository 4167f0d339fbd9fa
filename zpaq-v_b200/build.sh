#!/bin/sh
# Builds zpaq-v_b200/libzpaqgpu.so for sm_100a (in-tree, so the .so travels to the GPU box).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden"
mkdir -p build
for f in kernels_generic kernels_chain kernels_encpipe kernels_aux api jidac; do
  if [ ! -f build/$f.o ] || [ csrc/$f.cu -nt build/$f.o ] || [ csrc/common.cuh -nt build/$f.o ] || [ csrc/kernels.h -nt build/$f.o ] || [ csrc/model.h -nt build/$f.o ] || [ csrc/ctx.h -nt build/$f.o ] || [ ../include/zpaqgpu.h -nt build/$f.o ]; then
    $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o &
  fi
done
if [ ! -f build/model.o ] || [ csrc/model.cpp -nt build/model.o ] || [ csrc/model.h -nt build/model.o ]; then
  g++ -std=c++17 -O2 -fPIC -ffp-contract=off -fvisibility=hidden -c csrc/model.cpp -o build/model.o &
fi
wait
$NVCC -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o libzpaqgpu.so build/kernels_generic.o build/kernels_chain.o build/kernels_encpipe.o build/kernels_aux.o build/api.o build/jidac.o build/model.o
echo "built $(pwd)/libzpaqgpu.so"
