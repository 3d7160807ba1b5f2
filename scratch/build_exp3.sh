#!/bin/bash
# experiment variants of libtnb.so: arguments are name:"-DFLAG[=v] ..." pairs -> scratch/exp/libtnb_<name>.so, selected at
# run time with TNB_LIB_PATH.  The variant flags apply to ONE source file, SRC=svd (default) | gemm | qr | ...
set -e
cd "$(dirname "$0")/.."
SRC=${SRC:-svd}
L=tncontract_b200/lib; mkdir -p scratch/exp
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include"
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc $F $flags -c tncontract_b200/csrc/$SRC.cu -o scratch/exp/${SRC}_$name.o &
done
wait
for v in "$@"; do
  name=${v%%:*}
  objs=""
  for o in elementwise gemm mps_mpo permute prof qr tensordot svd; do
    if [ $o = $SRC ]; then objs="$objs scratch/exp/${SRC}_$name.o"; else objs="$objs $L/$o.o"; fi
  done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/exp/libtnb_$name.so $objs
done
rm -f scratch/exp/*.o
