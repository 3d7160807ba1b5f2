#!/bin/bash
# build experiment variants of libtnb.so with one phase of the Jacobi round kernel skipped (timing only; results are wrong)
set -e
cd "$(dirname "$0")/.."
L=tncontract_b200/lib; mkdir -p scratch/exp
for v in GRAM EIGEN APPLY; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -DTNB_EXP_SKIP_$v -c tncontract_b200/csrc/svd.cu -o scratch/exp/svd_$v.o &
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -DTNB_EXP_SKIP_GRAM -DTNB_EXP_SKIP_EIGEN -DTNB_EXP_SKIP_APPLY -c tncontract_b200/csrc/svd.cu -o scratch/exp/svd_ALL.o &
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -DTNB_EXP_EMPTY -c tncontract_b200/csrc/svd.cu -o scratch/exp/svd_EMPTY.o &
wait
for v in GRAM EIGEN APPLY ALL EMPTY; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/exp/libtnb_SKIP_$v.so $L/elementwise.o $L/gemm.o $L/mps_mpo.o $L/permute.o $L/prof.o $L/qr.o $L/tensordot.o scratch/exp/svd_$v.o
done
rm -f scratch/exp/*.o
