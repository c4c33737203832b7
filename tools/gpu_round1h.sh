#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ogs_sweep_h.log
for v in default u2_b8 u3_b6 u3_b5 u4_b4 u4_b5 u6_b3 u6_b4 u8_b2 u4_b3; do
  echo "== $v" >> gpurun_out/ogs_sweep_h.log
  if [ $v = default ]; then unset SPN_B200_LIB; else export SPN_B200_LIB=build/variants/libspn_$v.so; fi
  timeout 200 python tools/quick_ogs.py 48 7 none 2>&1 | grep "tile plan" >> gpurun_out/ogs_sweep_h.log
done
cat gpurun_out/ogs_sweep_h.log
unset SPN_B200_LIB
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
