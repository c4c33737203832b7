#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "ogs" 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_f3.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "ms | e2e", d["e2e"]["value"], "| frac", d["roofline"]["frac"])
print({k:(round(v.get("GB/s",0)),round(v.get("ms",0),4)) for k,v in d["suite"].items() if "GB/s" in v})
PY
tail -2 gpurun_out/bench_f3.err
