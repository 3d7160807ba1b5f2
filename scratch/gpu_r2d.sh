#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2d_batch.log
for b in 37 64 74 148; do
  timeout 600 python bench_batch.py --networks-per-gpu 296 --batch $b >> gpurun_out/r2d_batch.log 2>&1
  TNB_LIB_PATH=$PWD/scratch/exp/libtnb_TNB_EXP_NO_BATCH_HALVING.so timeout 600 python bench_batch.py --networks-per-gpu 296 --batch $b >> gpurun_out/r2d_batch.log 2>&1
done
cut -c90-330 gpurun_out/r2d_batch.log
