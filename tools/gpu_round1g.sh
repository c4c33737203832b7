#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_g.log
tail -5 gpurun_out/pytest_g.log
timeout 200 python tools/quick_ogs.py 48 7 none 2>&1 | tee gpurun_out/ogs_g.log
timeout 600 python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 2500 gpurun_out/bench_g.json; tail -5 gpurun_out/bench_g.err
