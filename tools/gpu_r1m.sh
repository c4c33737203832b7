#!/bin/bash
# round-1 GPU pass m (1 GPU): box facts, parity (incl. host-layer ogs_t tests), gs row-order A/B, bench both arms
mkdir -p gpurun_out
(nproc; free -g | head -2; nvidia-smi -L) > gpurun_out/box_m.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_m.log
tail -5 gpurun_out/pytest_m.log
timeout 600 python tools/quick_ogs.py 64 7 none > gpurun_out/quick_m.log 2>&1; tail -4 gpurun_out/quick_m.log
timeout 900 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -c 2500 gpurun_out/bench_m.json; tail -3 gpurun_out/bench_m.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_m.json 2> gpurun_out/bench_ref_m.err; tail -c 400 gpurun_out/bench_ref_m.json
