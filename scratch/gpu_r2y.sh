#!/bin/bash
# re + im table of W for the apply phase: Jacobi timing, tests, bench
mkdir -p gpurun_out
timeout 300 python scratch/jac_time.py > gpurun_out/r2y_jac.log 2>&1; cat gpurun_out/r2y_jac.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_pytest.log; tail -4 gpurun_out/r2y_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r2y.json 2> gpurun_out/r2y_bench_err.log; cut -c1-260 gpurun_out/bench_r2y.json; tail -3 gpurun_out/r2y_bench_err.log
