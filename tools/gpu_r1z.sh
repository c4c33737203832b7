#!/bin/bash
# round-1 GPU pass z (1 GPU): U / occupancy sweep of the warp-tile kernels after the instruction diet
mkdir -p gpurun_out; : > gpurun_out/quick_z.log
for f in streamparanumal_b200/lib/libspn_b200.so build/variants/*.so; do
  echo "== $f" | tee -a gpurun_out/quick_z.log
  SPN_B200_LIB=$PWD/$f timeout 300 python tools/quick_ogs.py 64 7 none 2>&1 | grep "tile plan" | tee -a gpurun_out/quick_z.log
done
