#!/bin/bash
# experiment variants of libtnb.so (svd.cu only): arguments are name:"-DFLAG[=v] ..." pairs -> scratch/exp/libtnb_<name>.so,
# selected at run time with TNB_LIB_PATH
set -e
cd "$(dirname "$0")/.."
L=tncontract_b200/lib; mkdir -p scratch/exp
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc $F $flags -c tncontract_b200/csrc/svd.cu -o scratch/exp/svd_$name.o &
done
wait
for v in "$@"; do
  name=${v%%:*}
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/exp/libtnb_$name.so $L/elementwise.o $L/gemm.o $L/mps_mpo.o $L/permute.o $L/prof.o $L/qr.o $L/tensordot.o scratch/exp/svd_$name.o
done
rm -f scratch/exp/*.o
