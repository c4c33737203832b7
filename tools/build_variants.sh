#!/bin/bash
# cross-compiles tuning variants of libspn_b200.so (different -D geometry of csrc/ogs_plan.cu) into build/variants/;
# usage: tools/build_variants.sh tag1 "-DSPN_WIN_S=6 ..." tag2 "..." ...   (run on the CPU box; the .so files travel with gpurun)
set -e
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude"
C=streamparanumal_b200/csrc
mkdir -p build/variants build/obj
for f in runtime stream_abi ogs_abi ogs_halo; do
  if [ ! -f build/obj/$f.o ] || [ $C/$f.cu -nt build/obj/$f.o ]; then $NVCC $FLAGS -c -o build/obj/$f.o $C/$f.cu & fi
done
wait
while [ $# -gt 1 ]; do
  tag=$1; defs=$2; shift 2
  ( $NVCC $FLAGS $defs -c -o build/obj/ogs_plan_$tag.o $C/ogs_plan.cu && \
    $NVCC $FLAGS -shared -o build/variants/libspn_b200_$tag.so build/obj/runtime.o build/obj/stream_abi.o build/obj/ogs_abi.o build/obj/ogs_halo.o build/obj/ogs_plan_$tag.o -cudart static && echo "built $tag" ) &
done
wait
