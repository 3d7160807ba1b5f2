#!/bin/bash
mkdir -p gpurun_out
TNB_LIB_PATH=scratch/exp/libtnb_st0.so timeout 120 python scratch/jac_stamps.py 2>&1 | tail -12
