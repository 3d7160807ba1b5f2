#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
TNB_LIB_PATH=$PWD/scratch/exp/libtnb_STAMPS.so timeout 120 python scratch/jac_stamps.py
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py svd 3
} > gpurun_out/exp11.log 2>&1
tail -32 gpurun_out/exp11.log
