#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python scratch/site_ops.py qrprof
} > gpurun_out/exp9.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_j.json 2> gpurun_out/bench_r1_j.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r01g python scratch/one_op.py svd > gpurun_out/ncu_exp9_j.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 300 -c 4 -f -o gpurun_out/prof_gemm_r01c python scratch/one_op.py qr > gpurun_out/ncu_exp9_g.log 2>&1
tail -12 gpurun_out/exp9.log; cat gpurun_out/bench_r1_j.json | cut -c1-300
