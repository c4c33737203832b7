#!/bin/bash
# round-1 GPU pass k: parity, bench (both arms), ncu launch list of the bench command, ncu --set full of every kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k.log
tail -3 gpurun_out/pytest_k.log
timeout 600 python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -c 1500 gpurun_out/bench_k.json; tail -3 gpurun_out/bench_k.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_k.json 2> gpurun_out/bench_ref_k.err; tail -c 600 gpurun_out/bench_ref_k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k.csv \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"copy_kernel|axpby_kernel|dot_kernel|cg_update_kernel|ogs_tile|ogs_scatter_cols" -c 16 -o gpurun_out/prof_all_k \
   python tools/prof_all.py 48 1 > gpurun_out/ncu_full_k.log 2>&1
tail -2 gpurun_out/ncu_full_k.log
ls -la gpurun_out
