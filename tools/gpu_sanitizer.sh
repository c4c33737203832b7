#!/bin/bash
# compute-sanitizer over the smoke invocation of every hot-path kernel family (stream kernels, CSR ogs kernels, tile plan,
# ogs_t::GatherScatter): memcheck, then racecheck.  Output: gpurun_out/sanitizer_{memcheck,racecheck}.txt
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 125 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.txt
  tail -4 gpurun_out/sanitizer_$tool.txt
done
