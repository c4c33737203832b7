#!/bin/bash
mkdir -p gpurun_out
timeout 55 ncu --set full --clock-control none -k regex:"ogs_tile_pf" --launch-skip 2 -c 2 -o gpurun_out/prof_ogs_f3 python tools/prof_ogs.py 64 > gpurun_out/ncu_ogs_f3.log 2>&1
tail -1 gpurun_out/ncu_ogs_f3.log
