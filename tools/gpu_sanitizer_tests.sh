#!/bin/bash
# compute-sanitizer over the single-process GPU parity tests of the kernels (C ABI) and of the host layer
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 125 $CS --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_golden.py -m gpu -q -x > gpurun_out/sanitizer_memcheck_tests.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_memcheck_tests.txt; tail -5 gpurun_out/sanitizer_memcheck_tests.txt
timeout 95 $CS --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/sanitizer_racecheck_tests.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_racecheck_tests.txt; tail -5 gpurun_out/sanitizer_racecheck_tests.txt
