#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py svd 3
} > gpurun_out/exp8.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --sites 20 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r01f python scratch/one_op.py svd > gpurun_out/ncu_exp8_j.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 330 -c 6 -f -o gpurun_out/prof_gemm_r01c python scratch/one_op.py qr > gpurun_out/ncu_exp8_g.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chol_inv -s 40 -c 1 -f -o gpurun_out/prof_cholinv_r01 python scratch/one_op.py qr > gpurun_out/ncu_exp8_c.log 2>&1
tail -12 gpurun_out/exp8.log; cat gpurun_out/bench_r1_i.json | cut -c1-400
