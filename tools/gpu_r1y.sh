#!/bin/bash
# round-1 GPU pass y (1 GPU): leaner warp-tile fold (opaque column mask, select-free FP64 add rounds) — parity, timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_host_ogs_gpu.py -m gpu -x -q > gpurun_out/pytest_y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_y.log
tail -3 gpurun_out/pytest_y.log
timeout 600 python tools/quick_ogs.py 64 7 none 2>&1 | grep "tile plan" | tee gpurun_out/quick_y.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_y.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], "ms | e2e", d["e2e"]["value"], "| frac", d["roofline"]["frac"])
print({k:(round(v.get("GB/s",0)),round(v.get("ms",0),4)) for k,v in d["suite"].items() if "GB/s" in v})
PY
tail -2 gpurun_out/bench_y.err
