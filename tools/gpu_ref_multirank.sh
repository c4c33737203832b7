#!/bin/bash
# The reference's own drivers and ogs_t, unchanged, at P ranks (one GPU each) through include/shim: parity self-check and the
# BS6-8 lines of its pairwise (host-staged) exchange beside the native layer's peer-memory exchange on the same mesh.
# usage (on a GPU box): tools/gpu_ref_multirank.sh P
P=${1:-2}
mkdir -p gpurun_out
out=gpurun_out/ref_multirank_${P}gpu.txt
{
echo "# $P ranks; $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)"
timeout 600 python -m pytest tests/test_ref_drivers.py -m gpu -q 2>&1 | tail -3
SPN=streamparanumal_b200/bin/spnrun
for nx in 4 16 32; do
  echo "## reference BS8 (its ogs_t, pairwise exchange over the MPI stand-in), $P ranks, nx=$nx per rank, degree 7"
  timeout 300 $SPN -np $P build/ref_drivers/BS8 -nx $nx -ny $nx -nz $nx -p 7 2>&1 | grep "BS8 ="
  echo "## native layer BS8 (tile plan, peer-memory exchange), same mesh"
  timeout 300 $SPN -np $P streamparanumal_b200/bin/BS8 -nx $nx -ny $nx -nz $nx -p 7 2>&1 | grep "BS8 ="
done
for n in 6 7; do
  echo "## reference BS$n, nx=16"
  timeout 300 $SPN -np $P build/ref_drivers/BS$n -nx 16 -ny 16 -nz 16 -p 7 2>&1 | grep "BS$n ="
done
} > $out 2>&1
cat $out
