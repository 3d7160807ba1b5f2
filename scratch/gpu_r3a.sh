#!/bin/bash
# split-K for tensordot products between one and two waves: tests, GEMM shapes, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3a_pytest.log; tail -4 gpurun_out/r3a_pytest.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r3a_gemm.log 2>&1; cat gpurun_out/r3a_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r3a.json 2> gpurun_out/r3a_bench_err.log; cut -c1-260 gpurun_out/bench_r3a.json; tail -3 gpurun_out/r3a_bench_err.log
