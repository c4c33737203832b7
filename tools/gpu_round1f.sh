#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ogs_sweep_f.log
for v in default u8_b3 u6_b3 u4_b4 u4_b6 u12_b2 u6_b4; do
  echo "== $v" >> gpurun_out/ogs_sweep_f.log
  if [ $v = default ]; then unset SPN_B200_LIB; else export SPN_B200_LIB=build/variants/libspn_$v.so; fi
  timeout 200 python tools/quick_ogs.py 48 7 none 2>&1 | grep "tile plan" >> gpurun_out/ogs_sweep_f.log
done
cat gpurun_out/ogs_sweep_f.log
unset SPN_B200_LIB
for s in 10 30 52; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:ogs_tile -s $s -c 1 -o gpurun_out/tile40_s${s}_r1f python tools/quick_ogs.py 40 7 none > gpurun_out/ncu_f_$s.log 2>&1; done
