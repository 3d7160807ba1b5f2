#!/bin/bash
# batched path (cfg 4): throughput against the batch size, with the device profile of one group
mkdir -p gpurun_out; rm -f gpurun_out/r3e_batch.log
for b in 148 296 592; do
  timeout 600 python bench_batch.py --networks-per-gpu 592 --batch $b >> gpurun_out/r3e_batch.log 2>> gpurun_out/r3e_err.log
done
cut -c90-330 gpurun_out/r3e_batch.log; tail -3 gpurun_out/r3e_err.log
