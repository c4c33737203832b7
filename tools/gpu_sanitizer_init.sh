#!/bin/bash
# compute-sanitizer initcheck and synccheck over the smoke invocation
mkdir -p gpurun_out
for tool in initcheck synccheck; do
  timeout 50 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 12 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.txt; tail -4 gpurun_out/sanitizer_$tool.txt
done
