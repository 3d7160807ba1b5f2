#!/bin/bash
# libtnb variants from the NEXT kernel sources (tncontract_b200/csrc_next/: not yet validated on the GPU, so not what
# build.py builds): arguments are name:"flags" pairs -> scratch/exp/libtnb_<name>.so; flags apply to the file SRC
# (default svd).  Files that have no csrc_next version come from lib/*.o.
set -e
cd "$(dirname "$0")/.."
SRC=${SRC:-svd}
L=tncontract_b200/lib; N=tncontract_b200/csrc_next; mkdir -p scratch/exp/obj
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include -I tncontract_b200/csrc"
for f in $N/*.cu; do
  b=$(basename $f .cu)
  if [ ! -f scratch/exp/obj/$b.o ] || [ $f -nt scratch/exp/obj/$b.o ]; then nvcc $F -c $f -o scratch/exp/obj/$b.o & fi
done
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  [ -n "$flags" ] && nvcc $F $flags -c $N/$SRC.cu -o scratch/exp/obj/${SRC}_$name.o &
done
wait
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  objs=""
  for o in elementwise gemm mps_mpo permute prof qr tensordot svd; do
    if [ $o = $SRC ] && [ -n "$flags" ]; then objs="$objs scratch/exp/obj/${SRC}_$name.o"
    elif [ -f scratch/exp/obj/$o.o ]; then objs="$objs scratch/exp/obj/$o.o"
    else objs="$objs $L/$o.o"; fi
  done
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o scratch/exp/libtnb_$name.so $objs
done
