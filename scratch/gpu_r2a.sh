#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2a_tests.log
for v in "" scratch/exp/libtnb_TNB_EXP_NO_FP32ROT.so scratch/exp/libtnb_TNB_EXP_NO_PDL.so; do
  TNB_LIB_PATH=$v timeout 300 python scratch/jac_time.py >> gpurun_out/r2a_jac.log 2>&1
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
cat gpurun_out/r2a_tests.log gpurun_out/r2a_jac.log; cut -c1-1500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
