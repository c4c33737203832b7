#!/bin/bash
# round-1 GPU pass r (1 GPU): host-buffer (pipelined e2e) path — parity, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -5 gpurun_out/pytest_r.log
timeout 900 python bench.py --no-suite --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; tail -c 900 gpurun_out/bench_r.json; tail -3 gpurun_out/bench_r.err
