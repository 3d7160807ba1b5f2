#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 300 python scratch/site_ops.py qr 3
timeout 300 python scratch/site_ops.py qrprof
} > gpurun_out/exp7.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err
tail -30 gpurun_out/exp7.log; cat gpurun_out/bench_r1_h.json; tail -5 gpurun_out/bench_r1_h.err
