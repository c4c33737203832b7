#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_d.log
tail -5 gpurun_out/pytest_d.log
rm -f gpurun_out/ogs_sweep_d.log
for v in default u4_b5 u2_b8 u3_b6 u8_b4 u2_b6 u1_b8; do
  echo "== $v" >> gpurun_out/ogs_sweep_d.log
  if [ $v = default ]; then unset SPN_B200_LIB; else export SPN_B200_LIB=build/variants/libspn_$v.so; fi
  timeout 200 python tools/quick_ogs.py 40 7 none 2>&1 | grep "tile plan" >> gpurun_out/ogs_sweep_d.log
done
cat gpurun_out/ogs_sweep_d.log
