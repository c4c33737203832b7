#!/bin/bash
# second GPU pass of round 1: parity of the pipelined ogs kernels, geometry sweep, stream sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_b.log
tail -3 gpurun_out/pytest_b.log
for v in t512_n2048 t256_n2048 t1024_n2048 t256_n1024 t512_n1024; do
  echo "== $v" >> gpurun_out/ogs_sweep_b.log
  SPN_B200_LIB=build/variants/libspn_$v.so timeout 200 python tools/quick_ogs.py 40 7 native >> gpurun_out/ogs_sweep_b.log 2>&1
done
cat gpurun_out/ogs_sweep_b.log
timeout 300 build/tune_stream 26 > gpurun_out/tune26_b.log 2>&1
grep "cta/sm  0" gpurun_out/tune26_b.log
