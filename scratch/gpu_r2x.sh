#!/bin/bash
# source-level ncu page of one Jacobi round (warp-stall samples per SASS instruction / source line)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r02d python scratch/one_op.py svd > gpurun_out/ncu_j3.log 2>&1
ncu -i gpurun_out/prof_jacobi_r02d.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_jacobi_r02d.source.csv 2> gpurun_out/ncu_j3_src.log
ncu -i gpurun_out/prof_jacobi_r02d.ncu-rep --page source --csv > gpurun_out/prof_jacobi_r02d.source_default.csv 2>> gpurun_out/ncu_j3_src.log
ls -la gpurun_out/prof_jacobi_r02d*; head -c 1500 gpurun_out/prof_jacobi_r02d.source.csv; tail -3 gpurun_out/ncu_j3_src.log
rm -f gpurun_out/prof_jacobi_r02d.ncu-rep
