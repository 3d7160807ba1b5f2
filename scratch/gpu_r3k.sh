#!/bin/bash
# warp-role placement variants of the rotation phase
mkdir -p gpurun_out
timeout 300 python scratch/jac_time.py 2>&1 | head -1
for v in roles1 roles2 roles3; do TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 300 python scratch/jac_time.py 2>&1 | head -1; done
for v in st0 st1; do TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 120 python scratch/jac_stamps.py 2>&1 | sed -n 1,8p; done
