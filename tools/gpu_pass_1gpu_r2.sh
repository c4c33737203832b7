#!/bin/bash
# round-2 final pass on one GPU: GPU parity suite, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2_final.log
tail -3 gpurun_out/pytest_r2_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_final.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "ms | e2e", d["e2e"]["value"], d["e2e"]["copy_ceiling"], "| frac", d["roofline"]["frac"], "| parity", d["parity_checked"], "| setup", d["config"]["host_setup_s"])
print({k:(round(v.get("GB/s",0)),round(v.get("ms",0),4)) for k,v in d["suite"].items() if "GB/s" in v}); print(d["tail"]); print(d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
PY
tail -2 gpurun_out/bench_r2_final.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_r2_final.json 2> gpurun_out/bench_ref_r2_final.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref_r2_final.json').read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"
