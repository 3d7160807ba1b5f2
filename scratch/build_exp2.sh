#!/bin/bash
# experiment variants of libtnb.so for A/B timing of the Jacobi round (round 2): -D flags given as arguments name the variant
set -e
cd "$(dirname "$0")/.."
L=tncontract_b200/lib; mkdir -p scratch/exp
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
for v in "$@"; do
  nvcc $F -D$v -c tncontract_b200/csrc/svd.cu -o scratch/exp/svd_$v.o &
done
wait
for v in "$@"; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/exp/libtnb_$v.so $L/elementwise.o $L/gemm.o $L/mps_mpo.o $L/permute.o $L/prof.o $L/qr.o $L/tensordot.o scratch/exp/svd_$v.o
done
rm -f scratch/exp/*.o
