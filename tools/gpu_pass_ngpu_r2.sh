#!/bin/bash
# round-2 multi-GPU pass: N-rank parity (both exchanges, one-way stress), bench.py at N (weak + strong + tail objects), tail through
# the BS executables
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q -k "multirank" > gpurun_out/pytest_r2_x$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2_x$N.log
tail -3 gpurun_out/pytest_r2_x$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_r2_x$N.json 2> gpurun_out/bench_r2_x$N.err
echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r2_x$N.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], d["ms_per_step"], "ms", "| e2e", d["e2e"]["value"], "| parity", d["parity_checked"], "| setup", d["config"]["host_setup_s"])
    print("strong", d.get("strong")); print("tail", d.get("tail"))
except Exception as e:
    print("no json", e)
PY
tail -2 gpurun_out/bench_r2_x$N.err
timeout 600 tools/gpu_tail_ngpu.sh $N > gpurun_out/tail_r2_x$N.txt 2>&1
grep -v NCCL gpurun_out/tail_r2_x$N.txt | cut -c1-150
