#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -6
