#!/bin/bash
# round-1 GPU pass q (1 GPU): the BS executables — tests, then every BASELINE config through the reference CLI
mkdir -p gpurun_out
B=streamparanumal_b200/bin
timeout 600 python -m pytest tests/test_bs_executables.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -4 gpurun_out/pytest_q.log
{
echo "# nproc=$(nproc) $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
echo "## C5 BS0 launch latency"; $B/BS0 -n 1000
echo "## C1 BS5 N=2^26"; $B/BS5 -m CUDA -b 1610612736
echo "## C2 BS1-BS4 size sweep (powers of 4, 1 KB .. 4 GB moved)"
for n in 1 2 3 4 5; do for b in 1024 4096 16384 65536 262144 1048576 4194304 16777216 67108864 268435456 1073741824 4294967296; do
  if [ $n -eq 5 ] && [ $b -gt 1610612736 ]; then continue; fi
  $B/BS$n -b $b; done; done
echo "## C3 BS6 / BS7, C4 BS8 at P=1: -nx 64 -ny 64 -nz 64 -p 7"
for n in 6 7 8; do $B/BS$n -m CUDA -nx 64 -ny 64 -nz 64 -p 7; done
echo "## C5 BS8 small sizes"
for nx in 2 4 8 16 32; do $B/BS8 -nx $nx -ny $nx -nz $nx -p 7; done
} > gpurun_out/bs_executables_q.txt 2>&1
grep -c "^BS" gpurun_out/bs_executables_q.txt; grep -A1 "^## C1\|^## C3" gpurun_out/bs_executables_q.txt; grep "^BS[678]" gpurun_out/bs_executables_q.txt | head -4; grep "4294967296\|268435456 " gpurun_out/bs_executables_q.txt | head; grep "BS1 = \[64 \|BS1 = \[65536 " gpurun_out/bs_executables_q.txt
