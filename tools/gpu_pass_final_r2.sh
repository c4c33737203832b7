#!/bin/bash
# round-2 closing pass on one GPU: smoke, bench with the setup phases on stderr, then the whole GPU suite
# (multi-rank peer-memory cases with the ranks sharing the GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SPN_SETUP_TIMING=1 timeout 300 python bench.py > gpurun_out/bench_r2_closing.json 2> gpurun_out/bench_r2_closing.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_closing.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "ms | e2e", d["e2e"]["value"], "| frac", d["roofline"]["frac"], "| parity", d["parity_checked"], "| setup", d["config"]["host_setup_s"])
print({k:(round(v.get("GB/s",0)),round(v.get("ms",0),4)) for k,v in d["suite"].items() if "GB/s" in v}); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
grep "ogs setup" gpurun_out/bench_r2_closing.err | head -12
timeout 1100 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_r2_closing.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2_closing.log
tail -30 gpurun_out/pytest_r2_closing.log
