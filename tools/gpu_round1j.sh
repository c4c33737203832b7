#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/quick_ogs.py 48 7 none 2>&1 | tee gpurun_out/ogs_j.log
OGS_NO_GS_ORDER=1 timeout 200 python tools/quick_ogs.py 48 7 none 2>&1 | tee -a gpurun_out/ogs_j.log
