#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_batched_gpu.py -x -q -m gpu 2>&1 | tail -30) > gpurun_out/r2c_tests.log
cat gpurun_out/r2c_tests.log
