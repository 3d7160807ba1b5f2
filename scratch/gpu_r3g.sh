#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r3g_batch.log
TNB_LIB_PATH=scratch/exp/libtnb_nohalve.so timeout 600 python bench_batch.py --networks-per-gpu 296 --batch 148 >> gpurun_out/r3g_batch.log 2>> gpurun_out/r3g_err.log
timeout 600 python bench_batch.py --networks-per-gpu 296 --batch 148 >> gpurun_out/r3g_batch.log 2>> gpurun_out/r3g_err.log
cut -c90-200 gpurun_out/r3g_batch.log; tail -3 gpurun_out/r3g_err.log
