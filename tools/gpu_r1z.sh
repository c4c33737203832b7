#!/bin/bash
# round-1 GPU pass z3 (1 GPU): last sweep of the gather+scatter warp tiles — store policy, L1 allocation of the value loads, CTA size
mkdir -p gpurun_out; : > gpurun_out/quick_z3.log
for f in build/v3/*.so; do
  echo "== $f" | tee -a gpurun_out/quick_z3.log
  SPN_B200_LIB=$PWD/$f timeout 200 python tools/quick_ogs.py 64 7 none 2>&1 | grep "tile plan" | tee -a gpurun_out/quick_z3.log
done
