#!/bin/sh
# Builds zpaq-v_b200/libzpaqgpu.so for sm_100a (in-tree, so the .so travels to the GPU box).
# Every compile job is waited for by PID and a failed one stops the build; an object is deleted before
# it is rebuilt, so a failed compile can never leave a stale object for the link step.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden"
CU="kernels_generic kernels_genwarp kernels_chain kernels_dectree kernels_encpipe kernels_aux api stream multi jidac"
mkdir -p build
pids=""
for f in $CU; do
  [ -f csrc/$f.cu ] || continue
  stale=0
  [ -f build/$f.o ] || stale=1
  for dep in csrc/$f.cu csrc/common.cuh csrc/kernels.h csrc/model.h csrc/ctx.h csrc/hostlogic.h csrc/sha1_lane.h ../include/zpaqgpu.h; do
    [ $stale -eq 0 ] && [ $dep -nt build/$f.o ] && stale=1
  done
  if [ $stale -eq 1 ]; then
    rm -f build/$f.o
    $NVCC $FLAGS -c csrc/$f.cu -o build/$f.o &
    pids="$pids $!"
  fi
done
if [ ! -f build/model.o ] || [ csrc/model.cpp -nt build/model.o ] || [ csrc/model.h -nt build/model.o ]; then
  rm -f build/model.o
  g++ -std=c++17 -O2 -fPIC -ffp-contract=off -fvisibility=hidden -c csrc/model.cpp -o build/model.o &
  pids="$pids $!"
fi
for p in $pids; do
  wait $p || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
OBJS="build/model.o"
for f in $CU; do
  [ -f csrc/$f.cu ] && OBJS="$OBJS build/$f.o"
done
$NVCC -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o libzpaqgpu.so $OBJS -ldl
echo "built $(pwd)/libzpaqgpu.so"
