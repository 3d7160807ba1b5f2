#!/bin/bash
# own Gram-partial buffer (one cluster barrier fewer) as the default: tests, shapes, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3d_pytest.log; tail -4 gpurun_out/r3d_pytest.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r3d_gemm.log 2>&1; cat gpurun_out/r3d_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r3d.json 2> gpurun_out/r3d_bench_err.log; cut -c1-260 gpurun_out/bench_r3a.json; tail -3 gpurun_out/r3d_bench_err.log
