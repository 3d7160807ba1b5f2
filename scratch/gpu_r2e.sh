#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2e_tests.log
timeout 300 python scratch/cfg5_full.py 6 3 32 > gpurun_out/r2e_cfg5.log 2>&1
timeout 900 python scratch/cfg5_full.py 8 4 256 >> gpurun_out/r2e_cfg5.log 2>&1
cat gpurun_out/r2e_tests.log; cut -c1-1500 gpurun_out/r2e_cfg5.log
