#!/bin/bash
# 4-GPU check: halo_t on 3 ranks (both exchanges), ogs_t at 4 ranks, one-way stress, then the small-size tail at 4 ranks
timeout 400 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q -k "halo_t or one_way or (multirank_ogs_t and 4-)" > gpurun_out/pytest_r2_x4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2_x4.log
tail -3 gpurun_out/pytest_r2_x4.log
timeout 200 tools/gpu_tail_ngpu.sh 4 > gpurun_out/tail_r2_x4.txt 2>&1
grep -v NCCL gpurun_out/tail_r2_x4.txt | cut -c1-120
