#!/bin/bash
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_kernels_gpu.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -2 > gpurun_out/pytest_after_cpasync.txt; cat gpurun_out/pytest_after_cpasync.txt
timeout 40 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 12 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_initcheck.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_initcheck.txt; tail -4 gpurun_out/sanitizer_initcheck.txt
