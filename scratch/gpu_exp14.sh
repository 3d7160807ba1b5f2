#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py
timeout 300 python scratch/site_ops.py svd 3
} > gpurun_out/exp14.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err
tail -10 gpurun_out/exp14.log; cut -c1-300 gpurun_out/bench_r1_n.json
