#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python scratch/site_ops.py gemm 3
timeout 300 python scratch/site_ops.py qr 3
} > gpurun_out/exp10.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_k.json 2> gpurun_out/bench_r1_k.err
tail -14 gpurun_out/exp10.log; cat gpurun_out/bench_r1_k.json | cut -c1-300
