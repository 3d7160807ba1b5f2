#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py svd 3
} > gpurun_out/exp13.log 2>&1
tail -12 gpurun_out/exp13.log
