#!/bin/bash
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_zz_cg_skeleton_gpu.py -m gpu -q -s > gpurun_out/cg_skeleton.txt 2>&1; echo "rc=$?" >> gpurun_out/cg_skeleton.txt
tail -15 gpurun_out/cg_skeleton.txt
