#!/bin/bash
# round-1 GPU pass x (8 GPUs): 8-rank parity, BS8 weak scaling at 8 ranks (peer-memory and NCCL exchange), strong scaling and
# the small-size tail through the BS executables (spnrun = mpirun stand-in)
N=${1:-8}
mkdir -p gpurun_out
B=streamparanumal_b200/bin
nvidia-smi topo -m > gpurun_out/topo_x$N.txt 2>&1
timeout 900 python -m pytest tests/test_host_ogs_gpu.py -m gpu -x -q -k "multirank and p2p and $N-" > gpurun_out/pytest_x$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_x$N.log
tail -3 gpurun_out/pytest_x$N.log
for mode in p2p nccl; do
  SPN_OGS_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_x${N}_$mode.json 2> gpurun_out/bench_x${N}_$mode.err
  echo "bench N=$N $mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_x${N}_$mode.json").read().strip().splitlines()[-1])
    print(d["value"], d["unit"], d["ms_per_step"], "ms", "| e2e", d["e2e"]["value"], "|", d["config"]["exchange"][:40], "| setup", d["config"]["host_setup_s"])
except Exception as e:
    print("no json", e)
PY
  tail -2 gpurun_out/bench_x${N}_$mode.err
done
{
echo "# $N ranks over spnrun; nproc=$(nproc)"
echo "## strong scaling: 64^3 elements in total (-nx 32 per rank at 8 ranks), degree 7"
$B/spnrun -np $N $B/BS8 -nx 32 -ny 32 -nz 32 -p 7
$B/spnrun -np $N $B/BS6 -nx 32 -ny 32 -nz 32 -p 7
$B/spnrun -np $N $B/BS7 -nx 32 -ny 32 -nz 32 -p 7
echo "## small-size tail, BS8, N ranks"
for nx in 2 4 8 16; do $B/spnrun -np $N $B/BS8 -nx $nx -ny $nx -nz $nx -p 7; done
echo "## BS5 N=2^26 per rank and small sizes, N ranks (independent replicas, rank 0 prints its own line)"
$B/spnrun -np $N $B/BS5 -b 1610612736 | head -2
$B/spnrun -np $N $B/BS5 -b 24576 | head -1
$B/spnrun -np $N $B/BS0 -n 1000 | head -1
} > gpurun_out/bs_executables_x$N.txt 2>&1
cat gpurun_out/bs_executables_x$N.txt | cut -c1-160
