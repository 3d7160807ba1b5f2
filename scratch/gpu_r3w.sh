#!/bin/bash
# device-side skip of the second-pass update of a group that is already orthonormal: tests, QR timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3w_pytest.log; tail -4 gpurun_out/r3w_pytest.log
TNB_LIB_PATH=scratch/exp/libtnb_noskip.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
TNB_LIB_PATH=scratch/exp/libtnb_skip13.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r3w.json 2> gpurun_out/r3w_bench_err.log; cut -c1-260 gpurun_out/bench_r3w.json; tail -3 gpurun_out/r3w_bench_err.log
