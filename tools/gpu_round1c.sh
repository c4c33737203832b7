#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
rm -f gpurun_out/ogs_sweep_c.log
for v in default u2 u8; do
  echo "== $v" >> gpurun_out/ogs_sweep_c.log
  if [ $v = default ]; then unset SPN_B200_LIB; else export SPN_B200_LIB=build/variants/libspn_$v.so; fi
  timeout 200 python tools/quick_ogs.py 40 7 native >> gpurun_out/ogs_sweep_c.log 2>&1
done
unset SPN_B200_LIB
cat gpurun_out/ogs_sweep_c.log
timeout 600 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 3000 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
