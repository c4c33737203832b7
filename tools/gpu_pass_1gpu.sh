#!/bin/bash
# round-1 GPU final pass (1 GPU): full parity suite, smoke, bench (both arms), ncu launch list of the bench command,
# ncu --set full of every hot-path kernel at full size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f.log
tail -3 gpurun_out/pytest_f.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_f.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "ms | e2e", d["e2e"]["value"], "| frac", d["roofline"]["frac"], "| cpu", d["cpu_baseline"]["value"])
print({k:(round(v.get("GB/s",0)),round(v.get("ms",0),4)) for k,v in d["suite"].items() if "GB/s" in v})
PY
tail -2 gpurun_out/bench_f.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_f.json 2> gpurun_out/bench_ref_f.err; tail -c 200 gpurun_out/bench_ref_f.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f.csv \
   python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"copy_kernel|axpby_kernel|dot_kernel|cg_update_kernel|ogs_tile|ogs_scatter_cols" -c 16 -o gpurun_out/prof_all_f \
   python tools/prof_all.py 64 1 > gpurun_out/ncu_full_f.log 2>&1
tail -1 gpurun_out/ncu_full_f.log
