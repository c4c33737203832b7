#!/bin/bash
# round-1 GPU pass w (1 GPU): full parity suite, bench (both arms), ncu launch list of the bench command, ncu --set full of every kernel at full size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_w.log
tail -3 gpurun_out/pytest_w.log
timeout 900 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; tail -c 600 gpurun_out/bench_w.json; tail -3 gpurun_out/bench_w.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_w.json 2> gpurun_out/bench_ref_w.err; tail -c 300 gpurun_out/bench_ref_w.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w.csv \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_w.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"copy_kernel|axpby_kernel|dot_kernel|cg_update_kernel|ogs_tile|ogs_scatter_cols" -c 16 -o gpurun_out/prof_all_w \
   python tools/prof_all.py 64 1 > gpurun_out/ncu_full_w.log 2>&1
tail -2 gpurun_out/ncu_full_w.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
