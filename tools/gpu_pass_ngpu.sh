#!/bin/bash
# round-1 GPU pass n (N GPUs): multi-rank parity of ogs_t over NCCL, bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_host_ogs_gpu.py tests/test_bs_executables.py -m gpu -x -q > gpurun_out/pytest_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n$N.log
tail -5 gpurun_out/pytest_n$N.log
for n in $(seq 1 8); do
  if [ $n -le $N ] && { [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; }; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
      bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    echo "bench N=$n rc=$?"; tail -c 1800 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
  fi
done
