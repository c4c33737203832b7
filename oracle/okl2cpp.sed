# Minimal OKL -> C++ rewrite giving the *Serial* semantics of streamParanumal's OKL
# files (BS/BS*/okl/*.okl, libs/ogs/okl/ogsKernels.okl).  Used only by oracle/Makefile
# to compile the reference kernels from where they lie into oracle/_ref/.
#  - @outer/@inner/@tile loops become plain nested loops in source order, which is
#    the execution order of OCCA's Serial backend;
#  - @shared arrays become locals of the enclosing outer iteration;
#  - @exclusive scalars stay plain scalars: every @exclusive in these files is
#    assigned the same value in each inner iteration, so one copy is equivalent.
s/@kernel void/extern "C" void/
s/@restrict//g
s/;[[:space:]]*@tile(.*)[[:space:]]*)[[:space:]]*{/){/
s/;[[:space:]]*@outer([0-9]*))/)/
s/;[[:space:]]*@inner([0-9]*))/)/
s/@shared //
s/@exclusive //
