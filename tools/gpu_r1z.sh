#!/bin/bash
# round-1 GPU pass z2 (1 GPU): occupancy / chunk sweep of the gather+scatter warp tiles around U=4, 4 CTAs
mkdir -p gpurun_out; : > gpurun_out/quick_z2.log
for c in 0 8 32; do
  echo "== default lib, chunk iters $c" | tee -a gpurun_out/quick_z2.log
  OGS_CHUNK=$c timeout 300 python tools/quick_ogs.py 64 7 none 2>&1 | grep "tile plan" | tee -a gpurun_out/quick_z2.log
done
for f in build/variants2/libspn_u4_b5.so build/variants2/libspn_u5_b4.so build/variants2/libspn_u3_b6.so; do
  echo "== $f" | tee -a gpurun_out/quick_z2.log
  SPN_B200_LIB=$PWD/$f timeout 300 python tools/quick_ogs.py 64 7 none 2>&1 | grep "tile plan" | tee -a gpurun_out/quick_z2.log
done
