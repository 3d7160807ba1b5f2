#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_cabi_gpu.py -x -q -m gpu 2>&1 | tail -3
for v in EMPTY ALL; do
  TNB_LIB_PATH=$PWD/scratch/exp/libtnb_SKIP_$v.so TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
done
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py svd 3
TNB_QR_FAST_PANEL=3 timeout 100 python scratch/panel_stamps.py
} > gpurun_out/exp3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r01e python scratch/one_op.py svd > gpurun_out/ncu_exp3_j.log 2>&1
tail -60 gpurun_out/exp3.log
