#!/bin/bash
# compute-sanitizer memcheck over the single-process GPU tests of the C++ host layer (ogs_t on a platform_t, graphs, host-buffer path)
mkdir -p gpurun_out
timeout 110 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_host_ogs_gpu.py tests/test_ref_maps.py -m gpu -q -x -k "not multirank" > gpurun_out/sanitizer_memcheck_host.txt 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_memcheck_host.txt; tail -6 gpurun_out/sanitizer_memcheck_host.txt
