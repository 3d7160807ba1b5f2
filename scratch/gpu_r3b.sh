#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python bench.py --no-batched --no-cpu-baseline > gpurun_out/bench_r3b_$i.json 2>> gpurun_out/r3b_err.log; cut -c70-200 gpurun_out/bench_r3b_$i.json
done
