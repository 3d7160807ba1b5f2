#!/bin/bash
# GPU experiment 1 (round 1, session 3): phase timing of the Jacobi round, per-op timings, ncu full captures
mkdir -p gpurun_out
{
for v in GRAM EIGEN APPLY ALL; do
  TNB_LIB_PATH=$PWD/scratch/exp/libtnb_SKIP_$v.so TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
done
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py all 3
timeout 300 python scratch/site_ops.py qrprof
} > gpurun_out/exp1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 2 -f -o gpurun_out/prof_jacobi_r01d python scratch/one_op.py svd > gpurun_out/ncu_exp1_j.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 400 -c 24 -f -o gpurun_out/prof_gemm_r01b python scratch/one_op.py qr > gpurun_out/ncu_exp1_g.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qr_panel -s 10 -c 2 -f -o gpurun_out/prof_panel_r01c python scratch/one_op.py qr > gpurun_out/ncu_exp1_p.log 2>&1
tail -50 gpurun_out/exp1.log
