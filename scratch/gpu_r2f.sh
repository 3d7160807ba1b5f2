#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_frows_parity.py -x -q -m gpu --tb=long 2>&1 | grep -v "^E  \|^    " | tail -60) > gpurun_out/r2f_tests.log
timeout 300 python scratch/jac_time.py > gpurun_out/r2f_jac.log 2>&1
cat gpurun_out/r2f_tests.log gpurun_out/r2f_jac.log
