#!/bin/bash
# small-size tail and strong scaling of BS8 / BS6 / BS7 / BS5 over N ranks through the BS executables (spnrun = mpirun
# stand-in), with the captured-call path on and off (SPN_OGS_GRAPH)
N=${1:-2}
B=streamparanumal_b200/bin
echo "# $N ranks over spnrun; nproc=$(nproc)"
for g in 1 0; do
  echo "## SPN_OGS_GRAPH=$g: BS8 small-size tail (elements per rank nx^3, degree 7)"
  for nx in 2 4 8 16; do SPN_OGS_GRAPH=$g $B/spnrun -np $N $B/BS8 -nx $nx -ny $nx -nz $nx -p 7; done
  echo "## SPN_OGS_GRAPH=$g: strong scaling, 64^3 elements in total"
  case $N in 2) d="32 64 64";; 4) d="32 32 64";; 8) d="32 32 32";; *) d="64 64 64";; esac
  set -- $d
  SPN_OGS_GRAPH=$g $B/spnrun -np $N $B/BS8 -nx $1 -ny $2 -nz $3 -p 7
done
echo "## BS6 / BS7 strong scaling"
$B/spnrun -np $N $B/BS6 -nx $1 -ny $2 -nz $3 -p 7
$B/spnrun -np $N $B/BS7 -nx $1 -ny $2 -nz $3 -p 7
echo "## BS5 N=2^26 per rank, small size; BS0"
$B/spnrun -np $N $B/BS5 -b 1610612736 | grep "^BS" | head -1
$B/spnrun -np $N $B/BS5 -b 24576 | grep "^BS" | head -1
$B/spnrun -np $N $B/BS0 -n 1000 | grep "^BS" | head -1
