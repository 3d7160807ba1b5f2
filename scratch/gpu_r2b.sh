#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2b_tests.log
timeout 300 python scratch/jac_time.py > gpurun_out/r2b_jac.log 2>&1
TNB_LIB_PATH=$PWD/scratch/exp/libtnb_TNB_EXP_STAMPS.so timeout 300 python scratch/jac_stamps.py 2>&1 | tail -12 >> gpurun_out/r2b_jac.log
cat gpurun_out/r2b_tests.log gpurun_out/r2b_jac.log
