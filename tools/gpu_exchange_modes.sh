#!/bin/bash
# round-1 GPU pass s (N GPUs): peer-memory halo exchange vs NCCL — multi-rank parity, bench at N ranks in both modes
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q -k multirank -s > gpurun_out/pytest_s$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s$N.log
grep -i "warning\|passed\|failed\|error\|rc=" gpurun_out/pytest_s$N.log | tail -8
for mode in p2p nccl; do
  SPN_OGS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_s${N}_$mode.json 2> gpurun_out/bench_s${N}_$mode.err
  echo "bench N=$N $mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_s${N}_$mode.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], d["ms_per_step"], "ms", "| e2e", d["e2e"]["value"], "|", d["config"]["exchange"][:60])
except Exception as e:
    print("no json", e)
PY
  tail -3 gpurun_out/bench_s${N}_$mode.err
done
