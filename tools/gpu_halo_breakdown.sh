#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q > gpurun_out/pytest_t$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_t$N.log
tail -3 gpurun_out/pytest_t$N.log
for mode in p2p nccl; do
SPN_OGS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/halo_breakdown.py 64 2>&1 | grep "^mask\|^chunk" | tee gpurun_out/halo_breakdown_${N}_$mode.txt
done
